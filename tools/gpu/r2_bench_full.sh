#!/bin/bash
# the default bench line (what the driver runs), both arms
cd /root/repo
mkdir -p gpurun_out
( time timeout 1200 python bench.py > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err ) 2>&1 | grep real
tail -3 gpurun_out/bench_r2_n1.err | cut -c1-300
python - <<'P'
import json
a=json.load(open('gpurun_out/bench_r2_n1.json'))
for k in ('value','seed_hits_per_s','gcells_per_s','ms_per_step','stage_ms_per_step','e2e','gpu_launches','clocks','cpu_baseline','speedup_vs_cpu_baseline','config4_one_gpu'): print(k, a.get(k))
print('roofline', {k: a['roofline'][k] for k in ('achieved','peak','frac','traffic','avg_launch_ms')})
for r in a['roofline_kernels']: print('  ', {k: (round(v,4) if isinstance(v,float) else v) for k,v in r.items() if k in ('kernel','ms_per_step','share_of_seed_stage','frac_of_hbm_peak','int32_frac','gapped_stage_ms_per_step')})
P
( time timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_r2_ref.json 2> gpurun_out/bench_r2_ref.err ) 2>&1 | grep real
cut -c1-700 gpurun_out/bench_r2_ref.json
