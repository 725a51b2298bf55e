#!/bin/bash
# first end-to-end GPU check: parity tests + verbose failures
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -60 | tee gpurun_out/pytest_gpu.log
