#!/bin/bash
# Round-2 multi-GPU call: bash scripts/gpu_run_multi.sh <N> <config> [also_dp_test]
# bench <config> at N = 1 (same box, for the efficiency) and at N ranks; with a third argument also the NCCL equivalence test.
N=${1:-2}; CFG=${2:-4}
mkdir -p gpurun_out
export FCD_DIST_TIMEOUT_S=120
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
if [ -n "$3" ]; then
  timeout -s KILL 600 python -m pytest tests/test_dp_gpu.py -q -p no:cacheprovider -s 2>&1 | grep -v "^$" | tail -n 8 | cut -c1-600
fi
show() { python -c "
import json
try:
    d=json.loads(open('$1').read()); print(d['value'], d['unit'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'n', d['n_gpus'], d['config']['launch'][:50], d['final_losses'])
except Exception as e: print('no json', e)"; }
timeout -s KILL 400 python bench.py --config $CFG --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/m_c${CFG}_n1.json 2> gpurun_out/m_c${CFG}_n1.err
echo "== config $CFG N=1 rc=$?"; show gpurun_out/m_c${CFG}_n1.json
timeout -s KILL 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N \
    --config $CFG --steps 10 --warmup 3 --max-seconds 420 > gpurun_out/m_c${CFG}_n$N.json 2> gpurun_out/m_c${CFG}_n$N.err
echo "== config $CFG N=$N rc=$?"; show gpurun_out/m_c${CFG}_n$N.json
grep -v "Warn\|warn\|^$\|first_losses\|run_backward\|OMP_NUM\|\*\*\*\|Consider using" gpurun_out/m_c${CFG}_n$N.err | tail -n 5 | cut -c1-300
