#!/bin/bash
# round 2, first look at the new gapped scheduler on the B200: parity tests, 5 Mbp against the unmodified reference,
# 50 Mbp timing with the scheduler's own profile line
cd /root/repo
mkdir -p /tmp/syn gpurun_out
echo "== gapped + cli gpu tests"
timeout 900 python -m pytest tests/test_gpu_gapped.py tests/test_gpu_cli.py -x -q -m gpu 2>&1 | tail -5
tools/gen_synth 5000000 20260925 /tmp/syn/t5.fa /tmp/syn/q5.fa
tools/gen_synth 50000000 20260925 /tmp/syn/t50.fa /tmp/syn/q50.fa
( time oracle/_ref/lastz /tmp/syn/t5.fa /tmp/syn/q5.fa > /tmp/syn/ref5.lav ) 2> /tmp/syn/ref5.time &
echo "== 5 Mbp"
( time LZB_GAP_PROFILE=1 lastz_b200/csrc/lastz_b200 /tmp/syn/t5.fa /tmp/syn/q5.fa --stats > /tmp/syn/our5.lav ) 2>&1 | grep -E "real|FAIL|gx profile|gapped:|backend" | cut -c1-600
echo "== 50 Mbp"
for W in 256 128; do
  ( time LZB_GAP_PROFILE=1 LZB_GAP_TRACE=1 lastz_b200/csrc/lastz_b200 /tmp/syn/t50.fa /tmp/syn/q50.fa --stats --speculation=$W > /tmp/syn/our50.$W.lav ) 2> gpurun_out/trace50_w$W.log
  grep -E "real|FAIL|gx profile|gapped:|backend|seed kernels" gpurun_out/trace50_w$W.log | cut -c1-600
  md5sum /tmp/syn/our50.$W.lav; grep -c "^a {" /tmp/syn/our50.$W.lav
done
wait
cat /tmp/syn/ref5.time | grep real
cmp <(sed 1,4d /tmp/syn/ref5.lav) <(sed 1,4d /tmp/syn/our5.lav) && echo REF5M_SAME
gzip -9 gpurun_out/trace50_w*.log
