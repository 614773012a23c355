#!/bin/bash
# config 3 at full size against the unmodified reference: the 50 Mbp query cut into 50 subranges of 1 Mbp
# (q.fa[a..b], the unit the reference treats independently), every subrange through oracle/_ref/lastz on the host
# cores and through the product on the GPU at the same time, LAV compared byte for byte (d-stanza aside).
cd /root/repo
mkdir -p /tmp/syn gpurun_out
tools/gen_synth 50000000 20260925 /tmp/syn/t50.fa /tmp/syn/q50.fa
NP=$(nproc); [ $NP -gt 16 ] && NP=16
t0=$(date +%s.%N)
( seq 0 49 | xargs -P $NP -I{} bash -c 'a=$(( {} * 1000000 + 1 )); b=$(( a + 999999 )); oracle/_ref/lastz /tmp/syn/t50.fa "/tmp/syn/q50.fa[$a..$b]" > /tmp/syn/ref.{}.lav 2>/dev/null'; echo "reference_wall_s $(awk "BEGIN{print $(date +%s.%N) - $t0}") on $NP processes" > /tmp/syn/ref.time ) &
t1=$(date +%s.%N)
for k in $(seq 0 49); do
  a=$(( k * 1000000 + 1 )); b=$(( a + 999999 ))
  lastz_b200/csrc/lastz_b200 /tmp/syn/t50.fa "/tmp/syn/q50.fa[$a..$b]" > /tmp/syn/our.$k.lav 2>/tmp/syn/our.$k.err || echo "product failed on shard $k: $(tail -1 /tmp/syn/our.$k.err)"
done
echo "product_wall_s $(awk "BEGIN{print $(date +%s.%N) - $t1}") (50 process starts, 50 index builds)" > /tmp/syn/our.time
wait
same=0; diff=0; aligns=0
for k in $(seq 0 49); do
  if cmp -s <(sed 1,4d /tmp/syn/ref.$k.lav) <(sed 1,4d /tmp/syn/our.$k.lav); then same=$((same+1)); else diff=$((diff+1)); echo "shard $k DIFFERS"; fi
  aligns=$(( aligns + $(grep -c '^a {' /tmp/syn/ref.$k.lav) ))
done
{ echo "config 3 (synthetic 50 Mbp x 50 Mbp, default options, both strands) as 50 query subranges of 1 Mbp:"
  echo "identical LAV: $same of 50 shards, different: $diff; $aligns alignments in the reference output"
  cat /tmp/syn/ref.time /tmp/syn/our.time
  cat /tmp/syn/ref.*.lav | md5sum | sed 's/-/all reference LAV files concatenated (with d-stanzas)/'; } | tee gpurun_out/r02_parity_50Mbp_all_shards.txt
