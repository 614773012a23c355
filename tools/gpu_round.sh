#!/bin/bash
# one GPU session: parity tests, CLI timings, ncu captures (reports land in gpurun_out/)
cd /root/repo
mkdir -p /tmp/syn gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
tools/gen_synth 5000000 20260925 /tmp/syn/t5.fa /tmp/syn/q5.fa
tools/gen_synth 1000000 20260925 /tmp/syn/t1.fa /tmp/syn/q1.fa
tools/gen_synth 50000000 20260925 /tmp/syn/t50.fa /tmp/syn/q50.fa
for L in 5 50; do
echo "== CLI $L Mbp"
( time lastz_b200/csrc/lastz_b200 /tmp/syn/t$L.fa /tmp/syn/q$L.fa --stats --speculation=32 > /tmp/syn/lav$L.txt ) 2>&1 | grep -E "backend|real|FAIL|seed kernels|gapped:"
done
echo "== ncu k_extend"
ncu --set full --clock-control none --import-source on -k regex:k_extend -c 1 -f -o gpurun_out/prof_extend lastz_b200/csrc/lastz_b200 /tmp/syn/t5.fa /tmp/syn/q5.fa --nogapped --format=segments --strand=plus > /dev/null 2> gpurun_out/ncu_extend.log; tail -2 gpurun_out/ncu_extend.log
echo "== ncu k_ydrop"
ncu --set full --clock-control none --import-source on -k regex:k_ydrop -c 1 -f -o gpurun_out/prof_ydrop lastz_b200/csrc/lastz_b200 /tmp/syn/t1.fa /tmp/syn/q1.fa --strand=plus --allocate:traceback=16M > /dev/null 2> gpurun_out/ncu_ydrop.log; tail -2 gpurun_out/ncu_ydrop.log
