#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
( time timeout 600 python bench.py > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err ) 2>&1 | grep real
tail -2 gpurun_out/bench_r2_final.err | cut -c1-300
python - <<'P'
import json
a=[json.loads(l) for l in open('gpurun_out/bench_r2_final.json') if l.startswith('{')][-1]
for k in ('value','seed_hits_per_s','gcells_per_s','ms_per_step','stage_ms_per_step','e2e','gpu_launches','clocks','speedup_vs_cpu_baseline','config4_one_gpu'): print(k, a.get(k))
print('roofline', {k: a['roofline'][k] for k in ('achieved','peak','frac','traffic','avg_launch_ms')})
P
timeout 400 tools/gpu/r2_parity_full.sh
