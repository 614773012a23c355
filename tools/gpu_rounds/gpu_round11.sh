#!/bin/bash
# do event records behind the DP kernels (or the upload stream's queue) serialise lanes beyond 32 streams?
cd /root/repo
mkdir -p /tmp/syn gpurun_out
tools/gen_synth 50000000 20260925 /tmp/syn/t50.fa /tmp/syn/q50.fa
for cfg in "64 0 0" "64 1 1" "64 0 1" "128 0 1" "48 0 1" "32 0 1"; do
  set -- $cfg
  ( time LZB_SPECULATION=$1 LZB_LANE_EVENTS=$2 LZB_UPLOAD_PRIO=$3 LZB_GAP_PROFILE=1 timeout 120 lastz_b200/csrc/lastz_b200 /tmp/syn/t50.fa /tmp/syn/q50.fa --stats > /tmp/syn/o.lav ) 2> gpurun_out/gap50d.log
  echo "-- W=$1 lane_events=$2 upload_prio=$3"; grep -E "real|FAIL|gx profile" gpurun_out/gap50d.log | grep -v "W=2 " | cut -c1-330
  md5sum /tmp/syn/o.lav | cut -c1-32
done
echo "expected md5 ae7f4fb3efd6ac7696c2fec524b777f7"
