#!/bin/bash
# closing run of round 2: the default bench line, the ncu launch list of one bench step, then as much of the GPU suite as fits
cd /root/repo
mkdir -p gpurun_out
( time timeout 900 python bench.py > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err ) 2>&1 | grep real
tail -3 gpurun_out/bench_r2_final.err | cut -c1-300
python - <<'P'
import json
a=[json.loads(l) for l in open('gpurun_out/bench_r2_final.json') if l.startswith('{')][-1]
for k in ('value','seed_hits_per_s','gcells_per_s','ms_per_step','stage_ms_per_step','e2e','gpu_launches','clocks','speedup_vs_cpu_baseline','config4_one_gpu'): print(k, a.get(k))
print('cpu_baseline', {k: a['cpu_baseline'][k] for k in ('value','seed_hits_per_s','gcells_per_s','cores','kind')})
print('roofline', {k: a['roofline'][k] for k in ('achieved','peak','frac','traffic','avg_launch_ms')})
for r in a['roofline_kernels']: print('  ', {k: (round(v,4) if isinstance(v,float) else v) for k,v in r.items() if k in ('kernel','ms_per_step','share_of_seed_stage','frac_of_hbm_peak','int32_frac','gapped_stage_ms_per_step')})
P
echo "== ncu launch list of the bench command (1 step, no warm-up, config3 only, resident pass only)"
( time timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches_bench_final.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-config4-base --resident-only > gpurun_out/r02_bench_under_ncu.log 2>&1 ) 2>&1 | grep real
wc -l gpurun_out/r02_launches_bench_final.csv
echo "== pytest -m gpu (gapped, seed, parity at size, adapter first)"
( time timeout ${1:-500} python -m pytest tests/test_gpu_gapped.py tests/test_gpu_seed.py tests/test_gpu_parity_at_size.py tests/test_adapter.py tests/test_maxwordcount.py -x -q -m gpu 2>&1 | tail -5 ) 2>&1 | tail -8
