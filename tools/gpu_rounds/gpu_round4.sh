#!/bin/bash
# one-warp kernel: per-row diff against the shared-memory kernel (K=16, K=24) + an ncu capture with source counters
cd /root/repo
mkdir -p /tmp/syn gpurun_out
timeout 500 python tools/dp_diff.py 2>&1 | tail -40
tools/gen_synth 1000000 20260925 /tmp/syn/t1.fa /tmp/syn/q1.fa
LZB_SPECULATION=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ydrop_warp -c 1 -f -o gpurun_out/prof_warp16 \
   lastz_b200/csrc/lastz_b200 /tmp/syn/t1.fa /tmp/syn/q1.fa --stats > /dev/null 2> gpurun_out/ncu_warp16.log
tail -3 gpurun_out/ncu_warp16.log
ls -la gpurun_out/prof_warp16.ncu-rep
