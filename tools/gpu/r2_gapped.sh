#!/bin/bash
# round 2: the gapped scheduler on the B200 -- parity tests, 50 Mbp timing with the scheduler's profile line, the same
# output from two different lane counts, and 4 x 1 Mbp query shards of the 50 Mbp pair against the unmodified reference
cd /root/repo
mkdir -p /tmp/syn gpurun_out
echo "== gapped gpu tests"
timeout 900 python -m pytest tests/test_gpu_gapped.py -x -q -m gpu 2>&1 | tail -3
tools/gen_synth 50000000 20260925 /tmp/syn/t50.fa /tmp/syn/q50.fa
for k in 0 1 2 3; do
  a=$((k * 12000000 + 1)); b=$((a + 999999))
  ( oracle/_ref/lastz /tmp/syn/t50.fa "/tmp/syn/q50.fa[$a..$b]" > /tmp/syn/ref50.$k.lav 2>/dev/null ) &
done
echo "== 50 Mbp"
for W in 256 32; do
  ( time LZB_GAP_PROFILE=1 LZB_GAP_TRACE=1 lastz_b200/csrc/lastz_b200 /tmp/syn/t50.fa /tmp/syn/q50.fa --stats --speculation=$W > /tmp/syn/our50.$W.lav ) 2> gpurun_out/trace50_w$W.log
  grep -E "real|FAIL|gx profile|gapped:|backend|seed kernels" gpurun_out/trace50_w$W.log | cut -c1-600
  sed 1,4d /tmp/syn/our50.$W.lav | md5sum; grep -c "^a {" /tmp/syn/our50.$W.lav
done
wait
echo "== shards vs reference"
for k in 0 1 2 3; do
  a=$((k * 12000000 + 1)); b=$((a + 999999))
  lastz_b200/csrc/lastz_b200 /tmp/syn/t50.fa "/tmp/syn/q50.fa[$a..$b]" > /tmp/syn/our50s.$k.lav
  cmp <(sed 1,4d /tmp/syn/ref50.$k.lav) <(sed 1,4d /tmp/syn/our50s.$k.lav) && echo "shard $k SAME ($(grep -c '^a {' /tmp/syn/our50s.$k.lav) alignments)"
done
gzip -9f gpurun_out/trace50_w*.log
