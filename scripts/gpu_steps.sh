#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python scripts/bench_steps.py --batch 8 --size 256 --steps 3 2>&1 | grep -v Warn | tail -4
head -45 gpurun_out/table_usss_parity_b8_256.txt
