#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python __graft_entry__.py smoke 2>&1 | grep -v Warning | tail -3 | tee gpurun_out/smoke.log
timeout -s KILL 900 python bench.py --steps 10 --warmup 3 --profile-out gpurun_out/table_parity.txt 2>gpurun_out/bench_parity.err | tee gpurun_out/bench_parity.json
timeout -s KILL 900 python bench.py --steps 10 --warmup 3 --graph off --no-cpu-baseline 2>gpurun_out/bench_parity_eager.err | tee gpurun_out/bench_parity_eager.json
timeout -s KILL 900 python bench.py --steps 10 --warmup 3 --precision fast --no-cpu-baseline --profile-out gpurun_out/table_fast.txt 2>gpurun_out/bench_fast.err | tee gpurun_out/bench_fast.json
for f in gpurun_out/bench_parity.err gpurun_out/bench_parity_eager.err gpurun_out/bench_fast.err; do tail -n 5 $f; done
