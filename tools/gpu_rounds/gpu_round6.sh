#!/bin/bash
# round-1 checkpoint on a fresh box: GPU parity suite, the default bench line (both arms), the ncu
# launch list of the bench command and one --set full capture per hot kernel.
cd /root/repo
mkdir -p /tmp/syn gpurun_out
echo "== gpu tests"
( time timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 ) 2>&1 | grep -vE "^(user|sys)"
echo "== bench (ours)"
( time timeout 420 python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err ) 2>&1 | grep real
tail -c 600 gpurun_out/bench_ours.err; cut -c1-1500 gpurun_out/bench_ours.json
echo "== bench (reference arm)"
( time timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ) 2>&1 | grep real
cut -c1-600 gpurun_out/bench_ref.json
echo "== ncu launch list of the bench command"
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
wc -l gpurun_out/launches_bench.csv
echo "== ncu full captures (5 Mbp CLI)"
tools/gen_synth 5000000 20260925 /tmp/syn/t5.fa /tmp/syn/q5.fa
for k in k_expand k_extend k_ydrop_mw; do
  skip=0; [ $k = k_ydrop_mw ] && skip=3
  LZB_SPECULATION=1 timeout 200 ncu --set full --clock-control none --import-source on -k regex:^$k -s $skip -c 1 -f -o gpurun_out/r01_full_$k \
     lastz_b200/csrc/lastz_b200 /tmp/syn/t5.fa /tmp/syn/q5.fa --stats > /dev/null 2> gpurun_out/ncu_$k.log
  tail -2 gpurun_out/ncu_$k.log | cut -c1-200
done
ls -la gpurun_out/
