#!/bin/bash
# one GPU session: parity tests, a short bench, ncu captures (reports land in gpurun_out/)
cd /root/repo
mkdir -p /tmp/syn gpurun_out
python -m pytest tests/test_gpu_gapped.py -x -q -m gpu 2>&1 | tail -3
tools/gen_synth 5000000 20260925 /tmp/syn/t5.fa /tmp/syn/q5.fa
tools/gen_synth 1000000 20260925 /tmp/syn/t1.fa /tmp/syn/q1.fa
echo "== CLI 5 Mbp"
( time lastz_b200/csrc/lastz_b200 /tmp/syn/t5.fa /tmp/syn/q5.fa --stats > /tmp/syn/lav5.txt ) 2>&1 | grep -E "backend|real|FAIL|seed kernels|gapped:"
echo "== bench (short)"
python bench.py --steps 1 --warmup 1 --size 5000000 --cpu-procs 4 2> gpurun_out/bench_err.log | tee gpurun_out/bench_short.json
tail -3 gpurun_out/bench_err.log
echo "== ncu k_extend"
ncu --set full --clock-control none --import-source on -k regex:k_extend -c 1 -f -o gpurun_out/prof_extend lastz_b200/csrc/lastz_b200 /tmp/syn/t5.fa /tmp/syn/q5.fa --nogapped --format=segments --strand=plus > /dev/null 2> gpurun_out/ncu_extend.log; tail -2 gpurun_out/ncu_extend.log
echo "== ncu k_ydrop"
ncu --set full --clock-control none --import-source on -k regex:k_ydrop -c 1 -f -o gpurun_out/prof_ydrop lastz_b200/csrc/lastz_b200 /tmp/syn/t1.fa /tmp/syn/q1.fa --strand=plus --allocate:traceback=16M > /dev/null 2> gpurun_out/ncu_ydrop.log; tail -2 gpurun_out/ncu_ydrop.log
ls -la gpurun_out/
