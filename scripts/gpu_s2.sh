#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 240 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "case14 or case15 or case16 or case17" 2>&1 | tail -15
echo "=== full"
bash scripts/gpu_all.sh
