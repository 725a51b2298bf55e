#!/bin/bash
# Same-box A/B of the engine switches on the headline configurations (box-to-box spread is ~3 %: only same-box pairs compare).
mkdir -p gpurun_out
timeout 300 python scripts/bench_small_kernels.py 16 2>&1 | tee gpurun_out/ab_small_kernels.log
for cfg in 2 g32; do
  for sw in ${AB_SWITCHES:-rowpack=0,batch_branches=0 rowpack=1,batch_branches=0 rowpack=1,batch_branches=1}; do
    FCD_ENGINE="$sw" timeout -s KILL 300 python bench.py --config $cfg --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline \
        > gpurun_out/ab_${cfg}_${sw//[,=]/_}.json 2> gpurun_out/ab.err
    python -c "
import json
d=json.loads(open('gpurun_out/ab_${cfg}_${sw//[,=]/_}.json').read())
print('config $cfg', '$sw', d['value'], d['unit'], 'ms', d['ms_per_step'], 'clk', d['clocks']['sm_mhz'])"
  done
done
