#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}; MODE=${2:-auto}
export FCD_DIST_TIMEOUT_S=60
timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --graph $MODE --max-seconds 120 > gpurun_out/bench_n${N}_$MODE.json 2>gpurun_out/bench_n${N}_$MODE.err
cut -c1-700 gpurun_out/bench_n${N}_$MODE.json
grep -v "^$" gpurun_out/bench_n${N}_$MODE.err | grep -vi "warn\|OMP_NUM\|\*\*\*\*" | tail -n 4
