#!/bin/bash
# 50 Mbp timing of the gapped scheduler (profile line + trace)
cd /root/repo
mkdir -p /tmp/syn gpurun_out
tools/gen_synth 50000000 20260925 /tmp/syn/t50.fa /tmp/syn/q50.fa
for W in 256 96; do
  ( time LZB_GAP_PROFILE=1 LZB_GAP_TRACE=1 lastz_b200/csrc/lastz_b200 /tmp/syn/t50.fa /tmp/syn/q50.fa --stats --speculation=$W > /tmp/syn/our50.$W.lav ) 2> gpurun_out/trace50_w$W.log
  grep -E "real|FAIL|gx profile|gapped:|backend" gpurun_out/trace50_w$W.log | cut -c1-600
  sed 1,4d /tmp/syn/our50.$W.lav | md5sum; grep -c "^a {" /tmp/syn/our50.$W.lav
done
gzip -9f gpurun_out/trace50_w*.log
