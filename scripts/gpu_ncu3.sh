#!/bin/bash
# ncu --set full of the conv engines at Generator-layer size + the K-major / MN-major halo descriptor probe
mkdir -p gpurun_out
timeout -s KILL 300 python scripts/probe_halo.py 2>&1 | grep -v Warning | tee gpurun_out/probe_halo.log
timeout -s KILL 600 ncu --set full --clock-control none --import-source on \
   -k regex:"conv_tc_kernel|conv_halo_kernel|wgrad_tc_kernel|wgrad_halo_kernel" \
   -s ${NCU_SKIP:-0} -c ${NCU_COUNT:-12} -o gpurun_out/prof_conv -f python scripts/ncu_kernels.py parity > gpurun_out/ncu_conv.log 2>&1
tail -3 gpurun_out/ncu_conv.log
ls -la gpurun_out/prof_conv.ncu-rep
