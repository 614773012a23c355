#!/bin/bash
cd /root/repo
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 1 --warmup 1 --size 5000000 2> gpurun_out/bench2_err.log | tee gpurun_out/bench_2gpu_5M.json | cut -c1-900
tail -3 gpurun_out/bench2_err.log | cut -c1-300
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 1 --warmup 0 --size 5000000 --impl reference --cpu-sample 50000 2>/dev/null | cut -c1-400
