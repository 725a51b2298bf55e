#!/bin/bash
# Round-2 GPU call E (1 GPU): root cause of the multi-graph (YieldingStep) failure with the Segmentor in the step.
mkdir -p gpurun_out
run() { name=$1; shift
  timeout -s KILL 300 "$@" > gpurun_out/e_$name.json 2> gpurun_out/e_$name.err
  echo "== $name rc=$?"; python -c "
import json,sys
try:
    d=json.loads(open('gpurun_out/e_$name.json').read()); print(d['value'], d['ms_per_step'], d['config']['launch'][:400])
except Exception as e: print('no json', e)"
  grep -v "Warn\|warn\|^$\|first_losses\|run_backward\|Consider using" gpurun_out/e_$name.err | tail -n 6 | cut -c1-400
}
B="python bench.py --config 4 --batch 4 --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline"
FCD_GRAPH_DEBUG=1 run c4_seg_debug $B --graph segmented
FCD_GRAPH_DEBUG=1 PYTORCH_NO_CUDA_MEMORY_CACHING=0 run c3_seg_debug python bench.py --config 3 --batch 2 --steps 2 --warmup 3 --no-cpu-baseline --no-gpu-baseline --graph segmented
FCD_GRAPH_DEBUG=1 run c2_seg_debug python bench.py --config 2 --batch 4 --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline --graph segmented
