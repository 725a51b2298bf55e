#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_steps_gpu.py -q -p no:cacheprovider 2>&1 | tail -n 4 | cut -c1-300
for prec in fast parity; do
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:conv_halo_kernel -s 2 -c 1 -o gpurun_out/l_conv_fwd_$prec -f python scripts/ncu_conv_fwd.py $prec > gpurun_out/l_ncu_$prec.log 2>&1
tail -2 gpurun_out/l_ncu_$prec.log | cut -c1-200
done
ls -la gpurun_out/*.ncu-rep
