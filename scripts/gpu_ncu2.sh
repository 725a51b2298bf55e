#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 ncu --set full --clock-control none --import-source on \
   -k regex:"conv_tc_kernel|wgrad_tc_kernel|bn_stats_kernel|bn_act_fwd_kernel|bn_act_bwd_reduce_kernel|bn_act_bwd_apply_kernel|stage_kernel|stage_pack4_kernel|masked_recon_fwd_kernel|masked_recon_bwd_kernel|ssim_fwd_kernel|ssim_bwd_kernel" \
   -s 12 -c 40 -o gpurun_out/prof_kernels -f python scripts/ncu_kernels.py parity > gpurun_out/ncu_kernels.log 2>&1
tail -3 gpurun_out/ncu_kernels.log
ls -la gpurun_out/prof_kernels.ncu-rep
