#!/bin/bash
# the recoverable and twin hit processors on a real GPU, as much as fits in the last GPU minute of the round:
# product command line against the oracle command line (same front end, CPU library) on the fixture pair and a 300 kbp pair
cd /root/repo
mkdir -p /tmp/syn gpurun_out
G=tests/golden
tools/gen_synth 300000 20260925 /tmp/syn/t.fa /tmp/syn/q.fa
n=0; same=0
check() {
  n=$((n+1))
  if cmp -s <(timeout 10 lastz_b200/csrc/lastz_b200 "$@" 2>&1 | grep -v '^  "') <(timeout 10 oracle/lastz_oracle "$@" 2>&1 | grep -v '^  "'); then same=$((same+1)); echo "SAME $*"; else echo "DIFFERENT $*"; fi
}
check $G/pseudocat.fa $G/pseudopig.fa --recoverseeds
check $G/pseudocat.fa $G/pseudopig.fa --twins=-5..30
check /tmp/syn/t.fa /tmp/syn/q.fa --recoverseeds --nogapped --format=general-
check /tmp/syn/t.fa /tmp/syn/q.fa --twins=0..50 --nogapped --format=general-
check /tmp/syn/t.fa /tmp/syn/q.fa --twins=10..100 --nogfextend --nogapped --format=general-
LZB_HIT_CAP=20000 check /tmp/syn/t.fa /tmp/syn/q.fa --twins=0..50 --nogapped --format=general-
echo "$same of $n identical" | tee gpurun_out/r02_hitproc_quick.txt
