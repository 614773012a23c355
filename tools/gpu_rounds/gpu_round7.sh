#!/bin/bash
# k_extend2 (warp-cooperative replay) against the first kernel: parity suite, segments identical at 50 Mbp,
# kernel times, host-side trace of the seed call, speculation-lane sweep for the gapped stage
cd /root/repo
mkdir -p /tmp/syn gpurun_out
echo "== seed parity suite (k_extend2 default)"
timeout 300 python -m pytest tests/test_gpu_seed.py -x -q -m gpu 2>&1 | tail -3
tools/gen_synth 50000000 20260925 /tmp/syn/t50.fa /tmp/syn/q50.fa
for v in 1 0; do
  echo "== CLI 50 Mbp nogapped LZB_EXTEND_V1=$v"
  ( time LZB_EXTEND_V1=$v LZB_SEED_TRACE=1 lastz_b200/csrc/lastz_b200 /tmp/syn/t50.fa /tmp/syn/q50.fa --stats --nogapped --format=segments > /tmp/syn/seg50.$v.txt ) 2>&1 | grep -E "real|FAIL|seed kernels|raw_seed|seed trace|query load" | cut -c1-400
done
cmp /tmp/syn/seg50.0.txt /tmp/syn/seg50.1.txt && echo SAME_SEGMENTS_50M
echo "== gapped lanes sweep (50 Mbp, full pipeline)"
for W in 32 64 128; do
  ( time LZB_SPECULATION=$W LZB_GAP_PROFILE=1 timeout 120 lastz_b200/csrc/lastz_b200 /tmp/syn/t50.fa /tmp/syn/q50.fa --stats > /tmp/syn/out50.$W.lav ) 2> gpurun_out/gap50_w$W.log
  echo "-- W=$W"; grep -E "real|FAIL|gx profile|gapped:" gpurun_out/gap50_w$W.log | cut -c1-420
  md5sum /tmp/syn/out50.$W.lav
done
echo "== bench, short, with the seed trace"
( time LZB_SEED_TRACE=1 timeout 200 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err ) 2>&1 | grep real
grep -E "seed trace|query load" gpurun_out/bench_short.err | tail -8 | cut -c1-300
python - <<'P'
import json
a=json.load(open('gpurun_out/bench_short.json'))
for k in ('value','gcells_per_s','ms_per_step','stage_ms_per_step','wall_ms_per_step','e2e','roofline'): print(k, a.get(k))
for r in a.get('roofline_kernels',[]): print(r)
P
