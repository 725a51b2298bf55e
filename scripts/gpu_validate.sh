#!/bin/bash
# Round-2 GPU call A: the whole -m gpu suite (all failures listed, no -x), smoke, then every bench configuration once.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader | head -1
timeout -s KILL 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -s 2>&1 | grep -v "^$" > gpurun_out/a_tests.log
tail -n 40 gpurun_out/a_tests.log | cut -c1-400
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/a_smoke.log 2>&1; tail -n 3 gpurun_out/a_smoke.log | cut -c1-300
for cfg in 2 g32 4 5 3; do
  extra="--no-cpu-baseline"; [ "$cfg" = "2" ] && extra=""
  timeout -s KILL 600 python bench.py --config $cfg --steps 10 --warmup 3 --streams 1 $extra --profile-out gpurun_out/a_table_$cfg.txt \
      > gpurun_out/a_bench_$cfg.json 2> gpurun_out/a_bench_$cfg.err
  echo "== config $cfg rc=$?"; cut -c1-600 gpurun_out/a_bench_$cfg.json; grep -v "Warn\|warn\|^$" gpurun_out/a_bench_$cfg.err | tail -n 3 | cut -c1-300
done
