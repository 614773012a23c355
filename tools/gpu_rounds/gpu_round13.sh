#!/bin/bash
# round-1 closing run: the GPU tests added since the last full-suite pass, smoke, the default bench line,
# one --set full capture of the Y-drop kernel, the ncu launch list of the bench command
cd /root/repo
mkdir -p /tmp/syn gpurun_out
echo "== new gpu tests + smoke"
timeout 200 python -m pytest tests/test_gpu_cli.py -x -q -m gpu -k "hwseeded or general or self" 2>&1 | tail -3
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench (ours, default flags)"
( time timeout 300 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err ) 2>&1 | grep real
tail -c 300 gpurun_out/bench_final.err
python - <<'P'
import json
a=json.load(open('gpurun_out/bench_final.json'))
for k in ('value','gcells_per_s','ms_per_step','stage_ms_per_step','wall_ms_per_step','e2e','cpu_baseline','gpu_launches','clocks','roofline'): print(k, a.get(k))
P
echo "== ncu full: k_ydrop_mw (5 Mbp pair, 4th DP launch)"
tools/gen_synth 5000000 20260925 /tmp/syn/t5.fa /tmp/syn/q5.fa
LZB_SPECULATION=1 timeout 150 ncu --set full --clock-control none --import-source on -k regex:^k_ydrop_mw -s 3 -c 1 -f -o gpurun_out/r01_full_k_ydrop_mw \
   lastz_b200/csrc/lastz_b200 /tmp/syn/t5.fa /tmp/syn/q5.fa --stats > /dev/null 2> gpurun_out/ncu_k_ydrop_mw.log
tail -2 gpurun_out/ncu_k_ydrop_mw.log | cut -c1-200
echo "== ncu launch list of the bench command (1 step, no warm-up)"
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_bench_final.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
wc -l gpurun_out/launches_bench_final.csv
