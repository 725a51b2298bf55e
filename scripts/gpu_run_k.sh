#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_steps_gpu.py -q -p no:cacheprovider 2>&1 | tail -n 6 | cut -c1-300
timeout -s KILL 300 python scripts/bench_conv_variants.py > gpurun_out/k_conv_variants.log 2>&1; cat gpurun_out/k_conv_variants.log | grep -v Warn | cut -c1-200
