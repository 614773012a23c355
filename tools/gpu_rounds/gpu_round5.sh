#!/bin/bash
# register-resident multi-warp Y-drop kernel: per-row diff, gapped tests, 50 Mbp timing
cd /root/repo
mkdir -p /tmp/syn gpurun_out
timeout 500 python tools/dp_diff.py 2>&1 | grep -v "^    row" | tail -30
echo "== gapped + cli tests"
timeout 900 python -m pytest tests/test_gpu_gapped.py tests/test_gpu_cli.py -q -m gpu 2>&1 | tail -6
tools/gen_synth 50000000 20260925 /tmp/syn/t50.fa /tmp/syn/q50.fa
echo "== 50 Mbp"
for cfg in "0 32" "2 32" "1 32"; do
  set -- $cfg
  echo "-- mode=$1 W=$2"
  ( time LZB_DP_MODE=$1 LZB_SPECULATION=$2 LZB_GAP_PROFILE=1 LZB_GAP_TRACE=1 timeout 600 lastz_b200/csrc/lastz_b200 /tmp/syn/t50.fa /tmp/syn/q50.fa --stats > /tmp/syn/out50.$1.$2.lav ) 2> gpurun_out/trace50_m$1_w$2.log
  grep -E "real|FAIL|gx profile|gapped:|backend" gpurun_out/trace50_m$1_w$2.log | grep -v "W=2 " | cut -c1-420
  echo "reruns: $(grep -c rerun gpurun_out/trace50_m$1_w$2.log)"; grep "done" gpurun_out/trace50_m$1_w$2.log | head -3
  md5sum /tmp/syn/out50.$1.$2.lav
done
