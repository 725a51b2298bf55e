#!/bin/bash
# Round-end validation on one B200: the whole -m gpu suite (-x, like the driver), smoke(), then every bench configuration once
# (config 2 with both baselines; the others with the cuDNN baseline), fast precision for 2 / g32.  Outputs: gpurun_out/v_*.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader | head -1
timeout -s KILL 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | grep -v "^$" > gpurun_out/v_tests.log
tail -n 3 gpurun_out/v_tests.log | cut -c1-300
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/v_smoke.log 2>&1; tail -n 2 gpurun_out/v_smoke.log | cut -c1-300
show() { python -c "
import json
try:
    d=json.loads(open('$1').read()); gb=d.get('gpu_baseline') or {}
    print(d['value'], d['unit'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'frac', d['roofline']['step_frac_of_peak'], 'dom', d['roofline']['kernel'], d['roofline']['frac'], 'cudnn', (gb.get('tf32') or {}).get('value'), (gb.get('fp32') or {}).get('value'), 'B', gb.get('batch'), 'cpu', (d.get('cpu_baseline') or {}).get('value'), 'loss_check', (d.get('loss_check') or {}).get('ok'))
except Exception as e: print('no json', e)"; }
timeout 300 python scripts/bench_small_kernels.py 16 > gpurun_out/v_small_kernels.log 2>&1; tail -n 14 gpurun_out/v_small_kernels.log
for cfg in 2 g32 4 5 3; do
  extra="--no-cpu-baseline"; [ "$cfg" = "2" ] && extra=""
  timeout -s KILL 900 python bench.py --config $cfg --steps 10 --warmup 3 $extra --profile-out gpurun_out/v_table_$cfg.txt \
      > gpurun_out/v_bench_$cfg.json 2> gpurun_out/v_bench_$cfg.err
  echo "== config $cfg rc=$?"; show gpurun_out/v_bench_$cfg.json
done
for cfg in 2 g32; do
  timeout -s KILL 600 python bench.py --config $cfg --steps 10 --warmup 3 --precision fast --no-cpu-baseline --no-gpu-baseline \
      --profile-out gpurun_out/v_table_${cfg}_fast.txt > gpurun_out/v_bench_${cfg}_fast.json 2> gpurun_out/v_bench_${cfg}_fast.err
  echo "== config $cfg fast rc=$?"; show gpurun_out/v_bench_${cfg}_fast.json
done
