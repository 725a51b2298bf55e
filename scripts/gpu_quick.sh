#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "conv" 2>&1 | tail -8 | cut -c1-300
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --profile-out gpurun_out/table_parity.txt 2>/dev/null | cut -c1-330
head -14 gpurun_out/table_parity.txt
