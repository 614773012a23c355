#!/bin/bash
cd /root/repo
mkdir -p /tmp/syn gpurun_out
tools/gen_synth 50000000 20260925 /tmp/syn/t50.fa /tmp/syn/q50.fa
echo "== CLI 50 Mbp threads=128 speculation=32"
( time LZB_GAP_TRACE=1 LZB_DP_THREADS=128 lastz_b200/csrc/lastz_b200 /tmp/syn/t50.fa /tmp/syn/q50.fa --stats --speculation=32 > /tmp/syn/lav50.txt ) 2> gpurun_out/trace50.log
grep -E "real|FAIL|gapped:|dp_cells" gpurun_out/trace50.log
python -m pytest tests/test_gpu_gapped.py -x -q -m gpu 2>&1 | tail -2
