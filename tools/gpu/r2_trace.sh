#!/bin/bash
# scheduler trace of a WARM gapped call (third call of the same process) on the 50 Mbp pair
cd /root/repo
mkdir -p gpurun_out
LZB_GAP_PROFILE=1 LZB_GAP_TRACE=1 timeout 150 python bench.py --steps 1 --warmup 2 --no-cpu-baseline --no-config4-base --no-overlap > gpurun_out/bench_t.json 2> gpurun_out/bench_t.err
grep "gx profile" gpurun_out/bench_t.err | grep -v "lanes=2 " | tail -3 | cut -c1-420
python tools/gx_trace.py < gpurun_out/bench_t.err | tail -80
gzip -9f gpurun_out/bench_t.err
