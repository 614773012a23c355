#!/bin/bash
# the default bench line (what the driver runs), then the ncu evidence for it: the launch list of one bench step and
# one --set full capture each of k_extend2 (50 Mbp pair) and of the Y-drop kernels (5 Mbp pair, serial scheduler so
# that one launch = one sweep)
cd /root/repo
mkdir -p /tmp/syn gpurun_out
( time timeout 1200 python bench.py > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err ) 2>&1 | grep real
tail -3 gpurun_out/bench_r2_n1.err | cut -c1-300
python - <<'P'
import json
a=json.load(open('gpurun_out/bench_r2_n1.json'))
for k in ('value','seed_hits_per_s','gcells_per_s','ms_per_step','stage_ms_per_step','e2e','gpu_launches','clocks','cpu_baseline','speedup_vs_cpu_baseline','config4_one_gpu'): print(k, a.get(k))
print('roofline', {k: a['roofline'][k] for k in ('achieved','peak','frac','traffic','avg_launch_ms')})
for r in a['roofline_kernels']: print('  ', {k: (round(v,4) if isinstance(v,float) else v) for k,v in r.items() if k in ('kernel','ms_per_step','share_of_seed_stage','frac_of_hbm_peak','int32_frac','gapped_stage_ms_per_step')})
P
echo "== ncu launch list of the bench command (1 step, no warm-up, config3 only)"
( time timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-config4-base --no-overlap > gpurun_out/r02_bench_under_ncu.log 2>&1 ) 2>&1 | grep real
wc -l gpurun_out/r02_launches_bench.csv
tools/gen_synth 50000000 20260925 /tmp/syn/t50.fa /tmp/syn/q50.fa
tools/gen_synth 5000000 20260925 /tmp/syn/t5.fa /tmp/syn/q5.fa
echo "== ncu full: k_extend2, 50 Mbp pair, plus strand"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:^k_extend2 -c 1 -f -o gpurun_out/r02_full_k_extend2_50M \
   lastz_b200/csrc/lastz_b200 /tmp/syn/t50.fa /tmp/syn/q50.fa --nogapped --strand=plus > /dev/null 2> gpurun_out/r02_ncu_k_extend2.log
tail -2 gpurun_out/r02_ncu_k_extend2.log | cut -c1-200
echo "== ncu full: k_ydrop_warp / k_ydrop_mw (5 Mbp pair, serial scheduler)"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:^k_ydrop_warp -s 3 -c 1 -f -o gpurun_out/r02_full_k_ydrop_warp \
   lastz_b200/csrc/lastz_b200 /tmp/syn/t5.fa /tmp/syn/q5.fa --stats --speculation=1 > /dev/null 2> gpurun_out/r02_ncu_k_ydrop_warp.log
tail -2 gpurun_out/r02_ncu_k_ydrop_warp.log | cut -c1-200
LZB_DP_MODE=0 timeout 200 ncu --set full --clock-control none --import-source on -k regex:^k_ydrop_mw -s 3 -c 1 -f -o gpurun_out/r02_full_k_ydrop_mw \
   lastz_b200/csrc/lastz_b200 /tmp/syn/t5.fa /tmp/syn/q5.fa --stats --speculation=1 > /dev/null 2> gpurun_out/r02_ncu_k_ydrop_mw.log
tail -2 gpurun_out/r02_ncu_k_ydrop_mw.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep | tail -4
