#!/bin/bash
# Round-2 GPU call C (2 GPUs): sharded == full-batch equivalence on NCCL, bench at N = 2 (eager vs captured collectives), config 4.
mkdir -p gpurun_out
export FCD_DIST_TIMEOUT_S=90
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout -s KILL 600 python -m pytest tests/test_dp_gpu.py -q -p no:cacheprovider -s 2>&1 | grep -v "^$" | tail -n 15 | cut -c1-400
tr() { # name, nproc, args...
  name=$1; n=$2; shift; shift
  timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $n "$@" \
      --max-seconds 240 > gpurun_out/c_bench_$name.json 2> gpurun_out/c_bench_$name.err
  echo "== $name rc=$?"; cut -c1-330 gpurun_out/c_bench_$name.json; grep -v "Warn\|warn\|^$\|first_losses\|run_backward\|OMP_NUM\|\*\*\*" gpurun_out/c_bench_$name.err | tail -n 4 | cut -c1-300
}
tr c2_n2 2 --config 2 --steps 20 --warmup 5
NCCL_GRAPH_REGISTER=0 tr c2_n2_captured 2 --config 2 --steps 20 --warmup 5 --collectives captured --max-seconds 120
tr c4_n2 2 --config 4 --steps 10 --warmup 3
tr c2_n2_eager 2 --config 2 --steps 10 --warmup 3 --graph off
