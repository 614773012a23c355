#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 300 tools/dp_bench 1000000 80 296 1 2>&1 | tee gpurun_out/dp_bench_r2c.txt
timeout 300 tools/dp_bench 1000000 80 296 0 2>&1 | tail -4 | tee -a gpurun_out/dp_bench_r2c.txt
tools/gpu/r2_trace.sh 2>&1 | grep -E "gx lanes|gx profile|call 4|starts|commits |late" | tail -12 | cut -c1-300
