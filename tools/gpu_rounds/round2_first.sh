#!/bin/bash
# FIRST GPU call of round 2 (written at the end of round 1, not yet run): everything that was changed after GPU time ran
# out in round 1 goes through the product on a B200, then the whole GPU suite, smoke and the
# default bench line.  Usage: gpurun --timeout 1500 -- tools/gpu_rounds/round2_first.sh
cd /root/repo
mkdir -p gpurun_out
echo "== front-end features verified on the CPU only so far (no -x: list every failure)"
timeout 600 python -m pytest tests/test_gpu_zz_frontend.py -q -m gpu 2>&1 | tail -15
echo "== whole GPU suite"
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
echo "== smoke"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench, both arms"
( time timeout 400 python bench.py > gpurun_out/r02_bench_first.json 2> gpurun_out/r02_bench_first.err ) 2>&1 | grep real
tail -c 300 gpurun_out/r02_bench_first.err
( time timeout 400 python bench.py --impl reference > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err ) 2>&1 | grep real
python - <<'P'
import json
for f in ('gpurun_out/r02_bench_first.json', 'gpurun_out/r02_bench_reference.json'):
    try:
        a = json.load(open(f))
    except Exception as e:
        print(f, 'unreadable:', e); continue
    for k in ('impl', 'value', 'gcells_per_s', 'ms_per_step', 'e2e', 'cpu_baseline', 'gpu_launches', 'clocks', 'roofline'):
        print(k, a.get(k))
P
