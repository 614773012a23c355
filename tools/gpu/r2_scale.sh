#!/bin/bash
# config 4 (250 Mbp x 250 Mbp, --chain) on N GPUs of one box, launched the way the driver launches it
# usage: gpurun --gpus N -- tools/gpu/r2_scale.sh N
cd /root/repo
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi -L | head -8
( time timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 2 --warmup 3 > gpurun_out/scale_r2_n$N.json 2> gpurun_out/scale_r2_n$N.err ) 2>&1 | grep real
tail -4 gpurun_out/scale_r2_n$N.err | cut -c1-300
python - $N <<'P'
import json,sys
n=sys.argv[1]
a=[json.loads(l) for l in open(f'gpurun_out/scale_r2_n{n}.json') if l.startswith('{')][-1]
for k in ('value','seed_hits_per_s','gcells_per_s','ms_per_step','stage_ms_per_step','e2e','gpu_launches','clocks','counts_per_step','scaling','config'): print(k, a.get(k))
P
