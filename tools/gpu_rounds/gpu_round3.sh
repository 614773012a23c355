#!/bin/bash
# one-warp Y-drop kernel: parity (GPU test suite) and timing against the shared-memory kernel at 50 Mbp
cd /root/repo
mkdir -p /tmp/syn gpurun_out
echo "== gpu tests (default: one-warp kernel first)"
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -8
tools/gen_synth 50000000 20260925 /tmp/syn/t50.fa /tmp/syn/q50.fa
tools/gen_synth 5000000 20260925 /tmp/syn/t5.fa /tmp/syn/q5.fa
echo "== 5 Mbp, warp kernel vs shared-memory kernel"
for m in 0 2; do
  ( time LZB_DP_MODE=$m LZB_GAP_PROFILE=1 timeout 300 lastz_b200/csrc/lastz_b200 /tmp/syn/t5.fa /tmp/syn/q5.fa --stats > /tmp/syn/out5.$m.lav ) 2>&1 | grep -E "real|FAIL|gx profile|gapped:|backend" | cut -c1-400
done
cmp /tmp/syn/out5.0.lav /tmp/syn/out5.2.lav && echo SAME_5M
echo "== 50 Mbp"
for cfg in "0 32" "2 32" "0 64" "0 16"; do
  set -- $cfg
  echo "-- mode=$1 W=$2"
  ( time LZB_DP_MODE=$1 LZB_SPECULATION=$2 LZB_GAP_PROFILE=1 LZB_GAP_TRACE=1 timeout 600 lastz_b200/csrc/lastz_b200 /tmp/syn/t50.fa /tmp/syn/q50.fa --stats > /tmp/syn/out50.$1.$2.lav ) 2> gpurun_out/trace50_m$1_w$2.log
  grep -E "real|FAIL|gx profile|gapped:|backend" gpurun_out/trace50_m$1_w$2.log | cut -c1-500
  grep -c rerun gpurun_out/trace50_m$1_w$2.log
  md5sum /tmp/syn/out50.$1.$2.lav
done
