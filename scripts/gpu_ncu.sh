#!/bin/bash
# ncu evidence for profiles/ (one GPU; never a multi-rank command):
#   1. launch list of ONE eager iteration of the headline workload (gpu__time_duration per launch);
#   2. `--set full --import-source on` of one launch of every dominant kernel inside that iteration;
#   3. the same for the forward convolution with fused BatchNorm statistics in both precisions.
# Reports land in gpurun_out/; summarise with scripts/summarize_launches.py / `ncu -i ... --page raw|source --csv`.
mkdir -p gpurun_out
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches_config_2.csv python scripts/ncu_step.py 2 parity 1 > gpurun_out/ncu_step.log 2>&1
wc -l gpurun_out/launches_config_2.csv
for k in wgrad_halo_kernel bn_act_bwd_apply_kernel bn_act_bwd_reduce_kernel bn_act_fwd_kernel conv_tc_kernel conv_halo_kernel; do
  timeout -s KILL 300 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$k -s 1 -c 1 \
      -o gpurun_out/ncu_$k -f python scripts/ncu_step.py 2 parity 1 > gpurun_out/ncu_$k.log 2>&1
  tail -1 gpurun_out/ncu_$k.log | cut -c1-150
done
for prec in fast parity; do
  timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:conv_halo_kernel -s 2 -c 1 \
      -o gpurun_out/ncu_conv_fwd_$prec -f python scripts/ncu_conv_fwd.py $prec > gpurun_out/ncu_conv_fwd_$prec.log 2>&1
  tail -1 gpurun_out/ncu_conv_fwd_$prec.log | cut -c1-150
done
ls -la gpurun_out/*.ncu-rep
