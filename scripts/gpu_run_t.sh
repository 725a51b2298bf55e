#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3; do
  timeout -s KILL 1200 python -m pytest tests -m gpu -x -q -p no:cacheprovider 2>&1 | grep -v "^$" > gpurun_out/t_tests_$i.log
  echo "full suite run $i: $(tail -n 1 gpurun_out/t_tests_$i.log | cut -c1-120)"
  grep -n "^E  .*Error\|^FAILED" gpurun_out/t_tests_$i.log | head -5 | cut -c1-300
done
for i in 1 2 3 4 5 6; do
  timeout -s KILL 300 python -m pytest tests/test_modules_gpu.py tests/test_steps_gpu.py tests/test_graph_gpu.py -q -p no:cacheprovider 2>&1 | grep -v "^$" > gpurun_out/t_sub_$i.log
  echo "subset run $i: $(tail -n 1 gpurun_out/t_sub_$i.log | cut -c1-120) $(grep -n '^E  .*Error' gpurun_out/t_sub_$i.log | head -3 | cut -c1-200 | tr '\n' ' ')"
done
