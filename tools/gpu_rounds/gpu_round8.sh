#!/bin/bash
# ncu --set full of k_extend2 (and the first k_extend) on one 130 M-hit chunk of the 50 Mbp pair
cd /root/repo
mkdir -p /tmp/syn gpurun_out
tools/gen_synth 50000000 20260925 /tmp/syn/t50.fa /tmp/syn/q50.fa
timeout 150 ncu --set full --clock-control none --import-source on -k regex:^k_extend2 -s 2 -c 1 -f -o gpurun_out/r01_full_k_extend2_50M \
   lastz_b200/csrc/lastz_b200 /tmp/syn/t50.fa /tmp/syn/q50.fa --nogapped --strand=plus --format=segments > /dev/null 2> gpurun_out/ncu_k_extend2_50M.log
tail -2 gpurun_out/ncu_k_extend2_50M.log | cut -c1-200
LZB_EXTEND_V1=1 timeout 150 ncu --set full --clock-control none --import-source on -k regex:^k_extend -s 2 -c 1 -f -o gpurun_out/r01_full_k_extend1_50M \
   lastz_b200/csrc/lastz_b200 /tmp/syn/t50.fa /tmp/syn/q50.fa --nogapped --strand=plus --format=segments > /dev/null 2> gpurun_out/ncu_k_extend1_50M.log
tail -2 gpurun_out/ncu_k_extend1_50M.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
