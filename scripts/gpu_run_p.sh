#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_modules_gpu.py tests/test_steps_gpu.py tests/test_ops_gpu.py tests/test_networks_gpu.py -q -p no:cacheprovider 2>&1 | grep -v "^$" > gpurun_out/p_tests.log
tail -n 5 gpurun_out/p_tests.log | cut -c1-300
grep -n "^E  \|Error" gpurun_out/p_tests.log | head -30 | cut -c1-400
timeout -s KILL 300 python scripts/bench_conv_variants.py 2>&1 | grep -v Warn | grep "x_slots=auto" | grep "tma_out=1" | cut -c1-200
