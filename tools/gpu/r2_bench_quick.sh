#!/bin/bash
# the new bench line on config3, strands overlapped and not, with the gapped scheduler's profile lines
cd /root/repo
mkdir -p gpurun_out
for mode in "--no-overlap"; do
  echo "== bench $mode"
  ( time LZB_GAP_PROFILE=1 timeout 600 python bench.py --steps 2 --warmup 2 --no-cpu-baseline --no-config4-base $mode > gpurun_out/bench_q$mode.json 2> gpurun_out/bench_q$mode.err ) 2>&1 | grep real
  grep "gx profile" gpurun_out/bench_q$mode.err | grep -v "lanes=2 " | tail -4 | cut -c1-420
  grep -v "gx profile" gpurun_out/bench_q$mode.err | tail -5
  python - "$mode" <<'P'
import json,sys
a=json.load(open('gpurun_out/bench_q%s.json' % sys.argv[1]))
for k in ('value','seed_hits_per_s','gcells_per_s','ms_per_step','stage_ms_per_step','e2e','gpu_launches','clocks'): print(k, a.get(k))
print(a['wall_ms_per_step'])
print(a['counts_per_step'])
P
done
