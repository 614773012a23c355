#!/bin/bash
cd /root/repo
mkdir -p /tmp/syn gpurun_out
echo "== tests mode 0 (register K=4 first)"
python -m pytest tests/test_gpu_gapped.py tests/test_gpu_cli.py -x -q -m gpu 2>&1 | tail -3
echo "== tests mode 1 (register K=8 first)"
LZB_DP_MODE=1 python -m pytest tests/test_gpu_gapped.py -x -q -m gpu 2>&1 | tail -3
tools/gen_synth 50000000 20260925 /tmp/syn/t50.fa /tmp/syn/q50.fa
for m in 0 2; do
echo "== CLI 50 Mbp mode=$m speculation=32"
( time LZB_GAP_TRACE=1 LZB_DP_MODE=$m lastz_b200/csrc/lastz_b200 /tmp/syn/t50.fa /tmp/syn/q50.fa --stats --speculation=32 > /tmp/syn/lav50.$m.txt ) 2> gpurun_out/trace50.$m.log
grep -E "real|FAIL|gapped:|dp_cells" gpurun_out/trace50.$m.log
done
cmp <(sed 1,4d /tmp/syn/lav50.0.txt) <(sed 1,4d /tmp/syn/lav50.2.txt) && echo SAME_OUTPUT
