#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
export FCD_DIST_TIMEOUT_S=90
timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 2>gpurun_out/bench_n$N.err | tee gpurun_out/bench_n$N.json
grep -v "^$" gpurun_out/bench_n$N.err | grep -vi "warn\|OMP_NUM\|\*\*\*\*" | tail -n 8
