#!/bin/bash
cd /root/repo
mkdir -p /tmp/syn gpurun_out
echo "== seed tests, split extension (default)"
python -m pytest tests/test_gpu_seed.py -x -q -m gpu 2>&1 | tail -3
echo "== seed tests, fused extension"
LZB_FUSED_EXTEND=1 python -m pytest tests/test_gpu_seed.py -x -q -m gpu 2>&1 | tail -2
tools/gen_synth 50000000 20260925 /tmp/syn/t50.fa /tmp/syn/q50.fa
for f in 0 1; do
echo "== CLI 50 Mbp nogapped fused=$f"
( time LZB_FUSED_EXTEND=$f lastz_b200/csrc/lastz_b200 /tmp/syn/t50.fa /tmp/syn/q50.fa --stats --nogapped --format=segments > /tmp/syn/seg50.$f.txt ) 2>&1 | grep -E "real|FAIL|seed kernels|raw_seed"
done
cmp /tmp/syn/seg50.0.txt /tmp/syn/seg50.1.txt && echo SAME_SEGMENTS
