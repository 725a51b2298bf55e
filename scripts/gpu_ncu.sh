#!/bin/bash
# launch list of one bench step (+ warm-up) and a full capture of the dominant tcgen05 conv kernel
mkdir -p gpurun_out
timeout -s KILL 1500 ncu --metrics gpu__time_duration.sum --clock-control none -s ${NCU_SKIP:-2000} -c ${NCU_COUNT:-1500} --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/ncu_bench.log
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 30 -c 3 \
    -o gpurun_out/prof_conv_tc -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out/
