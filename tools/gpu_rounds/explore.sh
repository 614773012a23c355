#!/bin/bash
# scratch: time the CLI on synthetic pairs of growing size
cd /root/repo
mkdir -p /tmp/syn gpurun_out
for L in 5000000 50000000; do
  tools/gen_synth $L 20260925 /tmp/syn/t$L.fa /tmp/syn/q$L.fa
  for W in 16 64; do
  echo "== L=$L gapped speculation=$W"
  ( time lastz_b200/csrc/lastz_b200 /tmp/syn/t$L.fa /tmp/syn/q$L.fa --stats --speculation=$W > /tmp/syn/lav$L.$W.txt ) 2>&1 | grep -E "backend|real|FAIL|seed kernels|gapped:"
  grep -c "^a {" /tmp/syn/lav$L.$W.txt
  done
  cmp /tmp/syn/lav$L.16.txt /tmp/syn/lav$L.64.txt && echo SAME_16_64
done
( time oracle/_ref/lastz /tmp/syn/t5000000.fa /tmp/syn/q5000000.fa > /tmp/syn/ref5.lav ) 2>&1 | grep real
cmp <(sed 1,4d /tmp/syn/ref5.lav) <(sed 1,4d /tmp/syn/lav5000000.16.txt) && echo REF5M_SAME
