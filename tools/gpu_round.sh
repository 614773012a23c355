#!/bin/bash
cd /root/repo
mkdir -p /tmp/syn gpurun_out
python -m pytest tests/test_gpu_gapped.py tests/test_gpu_cli.py -x -q -m gpu 2>&1 | tail -3
tools/gen_synth 50000000 20260925 /tmp/syn/t50.fa /tmp/syn/q50.fa
echo "== CLI 50 Mbp speculation=32"
( time LZB_GAP_TRACE=1 LZB_DP_THREADS=128 lastz_b200/csrc/lastz_b200 /tmp/syn/t50.fa /tmp/syn/q50.fa --stats --speculation=32 > /tmp/syn/lav50.a.txt ) 2> gpurun_out/trace50.log
grep -E "real|FAIL|gapped:|dp_cells|strict" gpurun_out/trace50.log
echo "== CLI 50 Mbp speculation=1 (sequential reference for equality)"
( time LZB_DP_THREADS=128 lastz_b200/csrc/lastz_b200 /tmp/syn/t50.fa /tmp/syn/q50.fa --stats --speculation=1 > /tmp/syn/lav50.b.txt ) 2>&1 | grep -E "real|FAIL|gapped:"
cmp <(sed 1,4d /tmp/syn/lav50.a.txt) <(sed 1,4d /tmp/syn/lav50.b.txt) && echo SAME_AS_SEQUENTIAL
