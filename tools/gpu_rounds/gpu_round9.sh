#!/bin/bash
# scheduler rework (commit-time retirement of covered anchors, 128 lanes), branch-free k_extend2 column step,
# parallel replay, seed scratch kept in the context: parity + timing
cd /root/repo
mkdir -p /tmp/syn gpurun_out
echo "== gpu parity suite"
timeout 420 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
tools/gen_synth 50000000 20260925 /tmp/syn/t50.fa /tmp/syn/q50.fa
echo "== 50 Mbp full pipeline"
for W in 128 64 32; do
  ( time LZB_SPECULATION=$W LZB_GAP_PROFILE=1 LZB_SEED_TRACE=1 timeout 120 lastz_b200/csrc/lastz_b200 /tmp/syn/t50.fa /tmp/syn/q50.fa --stats > /tmp/syn/out50.$W.lav ) 2> gpurun_out/gap50b_w$W.log
  echo "-- W=$W"; grep -E "real|FAIL|gx profile|gapped:|seed kernels|seed trace" gpurun_out/gap50b_w$W.log | grep -v "W=2 " | cut -c1-420
  md5sum /tmp/syn/out50.$W.lav
done
echo "expected md5 ae7f4fb3efd6ac7696c2fec524b777f7"
