#!/bin/bash
# Round-end evidence: headline bench (parity, fast), ncu launch list of the bench command, ncu --set full of the conv engines.
mkdir -p gpurun_out
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --profile-out gpurun_out/table_parity.txt 2>gpurun_out/bench_parity.err | tee gpurun_out/bench_parity.json | cut -c1-200
timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --precision fast --no-cpu-baseline --profile-out gpurun_out/table_fast.txt 2>/dev/null | tee gpurun_out/bench_fast.json | cut -c1-200
timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --graph off --no-cpu-baseline 2>/dev/null | tee gpurun_out/bench_parity_eager.json | cut -c1-200
# launch list of ONE eager iteration (bounded: ncu serialises and replays every profiled launch; a whole bench.py run under ncu takes > 15 min)
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file gpurun_out/launches.csv python scripts/ncu_step.py > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-200
wc -l gpurun_out/launches.csv
timeout -s KILL 600 ncu --set full --clock-control none --import-source on \
   -k regex:"conv_tc_kernel|conv_halo_kernel|wgrad_tc_kernel|wgrad_halo_kernel|bn_act_bwd_apply_kernel|bn_act_bwd_reduce_kernel|bn_act_fwd_kernel" \
   -s 0 -c 14 -o gpurun_out/prof_conv -f python scripts/ncu_kernels.py parity > gpurun_out/ncu_conv.log 2>&1
tail -2 gpurun_out/ncu_conv.log
ls -la gpurun_out/prof_conv.ncu-rep
