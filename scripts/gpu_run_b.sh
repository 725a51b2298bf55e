#!/bin/bash
# Round-2 GPU call B: full -m gpu suite again, side-stream A/B on the bench, config 3 at larger batches, fast precision, ncu launch list.
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -s 2>&1 | grep -v "^$" > gpurun_out/b_tests.log
tail -n 25 gpurun_out/b_tests.log | cut -c1-300
run() { # name, args...
  name=$1; shift
  timeout -s KILL 600 python bench.py "$@" --no-cpu-baseline --profile-out gpurun_out/b_table_$name.txt > gpurun_out/b_bench_$name.json 2> gpurun_out/b_bench_$name.err
  echo "== $name rc=$?"; cut -c1-330 gpurun_out/b_bench_$name.json; grep -v "Warn\|warn\|^$\|first_losses\|run_backward" gpurun_out/b_bench_$name.err | tail -n 3 | cut -c1-300
}
run c2_s1 --config 2 --steps 20 --warmup 5 --streams 1 --no-gpu-baseline
run c2_s2 --config 2 --steps 20 --warmup 5 --streams 2 --no-gpu-baseline
run g32_s2 --config g32 --steps 10 --warmup 3 --streams 2 --no-gpu-baseline
run c4_s2 --config 4 --steps 10 --warmup 3 --streams 2 --no-gpu-baseline
run c3_b16 --config 3 --batch 16 --steps 5 --warmup 3 --streams 1 --no-gpu-baseline
run c3_b32 --config 3 --batch 32 --steps 3 --warmup 3 --streams 1 --no-gpu-baseline
run c2_fast --config 2 --steps 20 --warmup 5 --streams 1 --precision fast --no-gpu-baseline
run g32_fast --config g32 --steps 10 --warmup 3 --streams 1 --precision fast --no-gpu-baseline
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/b_launches_c2.csv python scripts/ncu_step.py 2 parity 1 > gpurun_out/b_ncu_step.log 2>&1
tail -1 gpurun_out/b_ncu_step.log | cut -c1-200; wc -l gpurun_out/b_launches_c2.csv
