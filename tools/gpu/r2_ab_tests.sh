#!/bin/bash
# A/B of two builds of the seed stage on the 50 Mbp pair (lastz_b200/csrc/_ab holds the other build), then the GPU suite
cd /root/repo
mkdir -p /tmp/syn gpurun_out
tools/gen_synth 50000000 20260925 /tmp/syn/t50.fa /tmp/syn/q50.fa
for rep in 1 2; do
  for d in lastz_b200/csrc lastz_b200/csrc/_ab; do
    [ -x $d/lastz_b200 ] || continue
    echo "== $d"
    $d/lastz_b200 /tmp/syn/t50.fa /tmp/syn/q50.fa --nogapped --stats 2>&1 >/tmp/syn/out.$rep.$(basename $d).lav | grep -E "seed kernels|backend" | cut -c1-400
    md5sum < /tmp/syn/out.$rep.$(basename $d).lav
  done
done
echo "== pytest -m gpu"
( time timeout 1300 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 ) 2>&1 | tail -12
echo "== smoke"
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
