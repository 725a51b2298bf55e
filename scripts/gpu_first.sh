#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python scripts/dbg_g2.py 2>&1 | tail -40 | tee gpurun_out/dbg_g2.log
