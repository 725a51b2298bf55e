#!/bin/bash
bash scripts/gpu_tests.sh 30
timeout -s KILL 200 python scripts/dbg_alloc.py 2>&1 | grep -E "plain|dist" | cut -c1-1200
bash scripts/gpu_bench.sh
