#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python scripts/host_profile.py 2>&1 | grep -v "^ \|^$" | tail -12 | tee gpurun_out/host_profile2.log
