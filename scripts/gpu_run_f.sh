#!/bin/bash
# Round-2 GPU call F (1 GPU): graph tests after the dangling-gradient fix, segmented form of configs 3/4/5, config 3 at batch 32.
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_graph_gpu.py tests/test_networks_gpu.py -q -p no:cacheprovider 2>&1 | tail -n 12 | cut -c1-300
run() { name=$1; shift
  timeout -s KILL 600 python bench.py "$@" --no-cpu-baseline --no-gpu-baseline > gpurun_out/f_$name.json 2> gpurun_out/f_$name.err
  echo "== $name rc=$?"; python -c "
import json
try:
    d=json.loads(open('gpurun_out/f_$name.json').read()); print(d['value'], d['unit'], d['ms_per_step'], 'mem', d['config']['peak_mem_GiB'], d['config']['launch'][:60], d['final_losses'], d['first_losses'])
except Exception as e: print('no json', e)"
  grep -v "Warn\|warn\|^$\|first_losses\|run_backward\|Consider using" gpurun_out/f_$name.err | tail -n 4 | cut -c1-300
}
run c4_seg --config 4 --steps 5 --warmup 3 --graph segmented
run c4_whole --config 4 --steps 5 --warmup 3
run c5_seg --config 5 --steps 3 --warmup 3 --graph segmented
run c3_seg --config 3 --steps 3 --warmup 3 --graph segmented
run c3_b32 --config 3 --batch 32 --steps 3 --warmup 3
run c3_b24 --config 3 --batch 24 --steps 3 --warmup 3
