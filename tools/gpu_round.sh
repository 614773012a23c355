#!/bin/bash
cd /root/repo
mkdir -p /tmp/syn gpurun_out
python -m pytest tests/test_gpu_gapped.py tests/test_gpu_cli.py -x -q -m gpu 2>&1 | tail -3
tools/gen_synth 50000000 20260925 /tmp/syn/t50.fa /tmp/syn/q50.fa
for cfg in "256 32" "256 64" "128 64"; do
set -- $cfg
echo "== CLI 50 Mbp threads=$1 speculation=$2"
( time LZB_DP_THREADS=$1 lastz_b200/csrc/lastz_b200 /tmp/syn/t50.fa /tmp/syn/q50.fa --stats --speculation=$2 > /tmp/syn/lav50.$1.$2.txt ) 2>&1 | grep -E "real|FAIL|gapped:|dp_cells"
done
cmp <(sed 1,4d /tmp/syn/lav50.256.32.txt) <(sed 1,4d /tmp/syn/lav50.128.64.txt) && echo SAME_OUTPUT
