#!/bin/bash
# verification of the alt-extend kernel + --self paths (full GPU suite), DP shape experiment, k_extend2 with the skewed pair table
cd /root/repo
mkdir -p /tmp/syn gpurun_out
echo "== gpu parity suite"
timeout 500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
tools/gen_synth 50000000 20260925 /tmp/syn/t50.fa /tmp/syn/q50.fa
echo "== 50 Mbp full pipeline, DP shapes"
for S in 84 48 64; do
  ( time LZB_DP_SHAPE=$S LZB_GAP_PROFILE=1 timeout 120 lastz_b200/csrc/lastz_b200 /tmp/syn/t50.fa /tmp/syn/q50.fa --stats > /tmp/syn/o.lav ) 2> gpurun_out/gap50e_$S.log
  echo "-- shape=$S"; grep -E "real|FAIL|gx profile|gapped:|seed kernels" gpurun_out/gap50e_$S.log | grep -v "W=2 " | cut -c1-420
  md5sum /tmp/syn/o.lav | cut -c1-32
done
echo "expected md5 ae7f4fb3efd6ac7696c2fec524b777f7"
