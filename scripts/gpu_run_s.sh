#!/bin/bash
mkdir -p gpurun_out
run() { tag=$1; shift
  for i in 1 2 3 4; do
    env "$@" timeout -s KILL 300 python -m pytest tests/test_modules_gpu.py -q -p no:cacheprovider 2>&1 | grep -v "^$" > gpurun_out/s_${tag}_$i.log
    echo "$tag run $i: $(tail -n 1 gpurun_out/s_${tag}_$i.log | cut -c1-80) :: $(grep -n '^E  .*AssertionError' gpurun_out/s_${tag}_$i.log | head -2 | cut -c1-160 | tr '\n' ' ')"
  done
}
run default FCD_NOOP=1
run nohalo FCD_OPTIONS=conv_halo=0
run notma FCD_OPTIONS=conv_tma_out=0
run nowgradhalo FCD_OPTIONS=wgrad_halo=0
run blocking CUDA_LAUNCH_BLOCKING=1
