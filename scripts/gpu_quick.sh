#!/bin/bash
# Quick check of a change on one B200: the GPU tests of the touched paths first (fail fast), then the whole -m gpu suite, then
# short bench runs of config 2 / g32 without the slow baselines.  Outputs: gpurun_out/q_*.
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest "$@" -x -q -p no:cacheprovider 2>&1 | grep -v "^$" > gpurun_out/q_first.log
tail -n 25 gpurun_out/q_first.log | cut -c1-400
timeout -s KILL 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | grep -v "^$" > gpurun_out/q_tests.log
tail -n 12 gpurun_out/q_tests.log | cut -c1-400
for cfg in 2 g32; do
  timeout -s KILL 600 python bench.py --config $cfg --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline \
      --profile-out gpurun_out/q_table_$cfg.txt > gpurun_out/q_bench_$cfg.json 2> gpurun_out/q_bench_$cfg.err
  echo "== config $cfg rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/q_bench_$cfg.json').read())
print(d['value'], d['unit'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'], 'dom', d['roofline']['kernel'], d['roofline']['frac'], 'loss_check', (d.get('loss_check') or {}).get('ok'))" 2>&1 | tail -1
  head -22 gpurun_out/q_table_$cfg.txt
done
