#!/bin/bash
# conv-engine parity tests + headline bench with the per-call event table
mkdir -p gpurun_out
( timeout -s KILL 600 python -m pytest tests -m gpu -q -x --timeout 300 ${PYTEST_ARGS:-} 2>&1 | grep -vE "^\s*$|Warning|warnings.warn" | cut -c1-300 | tail -${TAILN:-12} ) 2>&1 | tee gpurun_out/pytest_gpu.log
timeout -s KILL 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/table_parity.txt 2>gpurun_out/bench_parity.err | tee gpurun_out/bench_parity.json | cut -c1-400
tail -n 3 gpurun_out/bench_parity.err | cut -c1-300
head -${TABLEN:-22} gpurun_out/table_parity.txt
