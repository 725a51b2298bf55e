#!/bin/bash
# ncu --set full + source for the other dominant kernels of the config-2 iteration (one launch each, eager iteration 4)
mkdir -p gpurun_out
for k in wgrad_halo_kernel bn_act_bwd_apply_kernel bn_act_bwd_reduce_kernel bn_act_fwd_kernel conv_tc_kernel; do
  timeout -s KILL 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$k -s 1 -c 1 \
      -o gpurun_out/n_$k -f python scripts/ncu_step.py 2 parity 1 > gpurun_out/n_$k.log 2>&1
  tail -1 gpurun_out/n_$k.log | cut -c1-150
done
ls -la gpurun_out/n_*.ncu-rep
