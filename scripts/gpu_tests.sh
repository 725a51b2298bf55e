#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 1200 python -m pytest tests -m gpu -q --timeout 600 2>&1 | grep -vE "^\s*$" | cut -c1-400 | tail -${1:-60} | tee gpurun_out/pytest_gpu.log
