#!/bin/bash
# Round-2 GPU call J (1 GPU): step tests (lean == faithful), configs 3/4/5 lean vs faithful.
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_steps_gpu.py tests/test_graph_gpu.py -q -p no:cacheprovider 2>&1 | tail -n 12 | cut -c1-300
run() { name=$1; shift
  timeout -s KILL 600 python bench.py "$@" --no-cpu-baseline --no-gpu-baseline --profile-out gpurun_out/j_table_$name.txt > gpurun_out/j_$name.json 2> gpurun_out/j_$name.err
  echo "== $name rc=$?"; python -c "
import json
try:
    d=json.loads(open('gpurun_out/j_$name.json').read()); print(d['value'], d['unit'], d['ms_per_step'], 'mem', d['config']['peak_mem_GiB'], 'frac', d['roofline']['step_frac_of_peak'], d['final_losses'], d['first_losses'])
except Exception as e: print('no json', e)"
  grep -v "Warn\|warn\|^$\|first_losses\|run_backward\|Consider using" gpurun_out/j_$name.err | tail -n 4 | cut -c1-300
}
run c4_lean --config 4 --steps 10 --warmup 3
run c4_faithful --config 4 --steps 10 --warmup 3 --form faithful
run c5_lean --config 5 --steps 5 --warmup 3
run c3_lean --config 3 --steps 3 --warmup 3
run c3_lean_b8 --config 3 --batch 8 --steps 5 --warmup 3
