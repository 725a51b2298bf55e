#!/bin/bash
# Runs every gpu_probe section in its own process (a faulting kernel must not take the others down)
# under a hard timeout (a hung kernel must not hold the GPU box).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.sm,clocks.max.sm --format=csv > gpurun_out/probe_smi.txt 2>&1
for sec in "$@"; do
  echo "##### section $sec" | tee -a gpurun_out/probe.log
  timeout -s KILL 240 python scripts/gpu_probe.py $sec 2>&1 | tee -a gpurun_out/probe.log
  echo "##### section $sec exit=${PIPESTATUS[0]}" | tee -a gpurun_out/probe.log
done
