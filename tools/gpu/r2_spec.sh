#!/bin/bash
# gapped phase with the two strands' schedulers side by side: lanes, resident cap, against one after the other
cd /root/repo
mkdir -p gpurun_out
run() {  # name, env..., -- args
  name=$1; shift
  ( time env "$@" LZB_GAP_PROFILE=1 timeout 200 python bench.py --no-cpu-baseline --no-config4-base --resident-only --steps 2 --warmup 3 $ARGS > gpurun_out/bench_spec_$name.json 2> gpurun_out/bench_spec_$name.err ) 2>&1 | grep real
  python - $name <<'P'
import json,sys
a=[json.loads(l) for l in open(f'gpurun_out/bench_spec_{sys.argv[1]}.json') if l.startswith('{')][-1]
print(sys.argv[1], 'ms_per_step', round(a['ms_per_step'],1), round(a['stage_ms_per_step']['seed'],1), round(a['stage_ms_per_step']['gapped'],1), 'gcells', round(a['gcells_per_s'],1))
P
  grep "gx profile" gpurun_out/bench_spec_$name.err | tail -2 | cut -c1-330
}
ARGS="--overlap-extend-ctas 2" run stagger_e2 LZB_WARP_SMEM_KB=23
ARGS="--overlap-extend-ctas 1" run stagger_e1 LZB_WARP_SMEM_KB=23
ARGS="--overlap-extend-ctas 3" run stagger_e3 LZB_WARP_SMEM_KB=23
