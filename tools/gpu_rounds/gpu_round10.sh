#!/bin/bash
# uploads through mapped staging + copy kernel, preallocated segment table: gapped parity, lane sweep, bench, k_extend2 capture
cd /root/repo
mkdir -p /tmp/syn gpurun_out
echo "== gapped + cli parity"
timeout 420 python -m pytest tests/test_gpu_gapped.py tests/test_gpu_cli.py -x -q -m gpu 2>&1 | tail -3
tools/gen_synth 50000000 20260925 /tmp/syn/t50.fa /tmp/syn/q50.fa
echo "== 50 Mbp full pipeline"
for W in 128 64 32; do
  ( time LZB_SPECULATION=$W LZB_GAP_PROFILE=1 timeout 120 lastz_b200/csrc/lastz_b200 /tmp/syn/t50.fa /tmp/syn/q50.fa --stats > /tmp/syn/out50.$W.lav ) 2> gpurun_out/gap50c_w$W.log
  echo "-- W=$W"; grep -E "real|FAIL|gx profile|gapped:" gpurun_out/gap50c_w$W.log | grep -v "W=2 " | cut -c1-420
  md5sum /tmp/syn/out50.$W.lav
done
echo "expected md5 ae7f4fb3efd6ac7696c2fec524b777f7"
echo "== bench (ours, default flags)"
( time timeout 300 python bench.py > gpurun_out/bench_ours2.json 2> gpurun_out/bench_ours2.err ) 2>&1 | grep real
tail -c 400 gpurun_out/bench_ours2.err
python - <<'P'
import json
a=json.load(open('gpurun_out/bench_ours2.json'))
for k in ('value','gcells_per_s','ms_per_step','stage_ms_per_step','wall_ms_per_step','e2e','cpu_baseline','gpu_launches','clocks'): print(k, a.get(k))
P
echo "== ncu full: k_extend2, one 130 M-hit chunk"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:^k_extend2 -s 2 -c 1 -f -o gpurun_out/r01_full_k_extend2b_50M \
   lastz_b200/csrc/lastz_b200 /tmp/syn/t50.fa /tmp/syn/q50.fa --nogapped --strand=plus --format=segments > /dev/null 2> gpurun_out/ncu_k_extend2b_50M.log
tail -1 gpurun_out/ncu_k_extend2b_50M.log | cut -c1-200
