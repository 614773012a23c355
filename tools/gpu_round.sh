#!/bin/bash
cd /root/repo
mkdir -p /tmp/syn gpurun_out
python -m pytest tests/test_gpu_gapped.py -x -q -m gpu 2>&1 | tail -3
tools/gen_synth 50000000 20260925 /tmp/syn/t50.fa /tmp/syn/q50.fa
for W in 32 64 128; do
echo "== CLI 50 Mbp speculation=$W"
( time LZB_DP_THREADS=128 lastz_b200/csrc/lastz_b200 /tmp/syn/t50.fa /tmp/syn/q50.fa --stats --speculation=$W > /tmp/syn/lav50.$W.txt ) 2>&1 | grep -E "real|FAIL|gapped:|dp_cells|strict"
done
cmp <(sed 1,4d /tmp/syn/lav50.32.txt) <(sed 1,4d /tmp/syn/lav50.128.txt) && echo SAME_32_128
for T in 256; do
echo "== CLI 50 Mbp speculation=64 threads=$T"
( time LZB_DP_THREADS=$T lastz_b200/csrc/lastz_b200 /tmp/syn/t50.fa /tmp/syn/q50.fa --stats --speculation=64 > /tmp/syn/lav50.t.txt ) 2>&1 | grep -E "real|FAIL|gapped:"
done
