#!/bin/bash
# gapped phase with the two strands' schedulers side by side: how many anchors in flight (both strands together)
cd /root/repo
mkdir -p gpurun_out
for W in 280 200; do
  ( time timeout 300 python bench.py --no-cpu-baseline --no-config4-base --resident-only --steps 2 --warmup 3 --speculation $W > gpurun_out/bench_spec_$W.json 2> gpurun_out/bench_spec_$W.err ) 2>&1 | grep real
  python - $W <<'P'
import json,sys
a=[json.loads(l) for l in open(f'gpurun_out/bench_spec_{sys.argv[1]}.json') if l.startswith('{')][-1]
print(sys.argv[1], 'ms_per_step', round(a['ms_per_step'],1), a['stage_ms_per_step']['seed'], a['stage_ms_per_step']['gapped'], 'gcells', round(a['gcells_per_s'],1))
P
done
