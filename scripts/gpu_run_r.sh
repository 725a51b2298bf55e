#!/bin/bash
mkdir -p gpurun_out
for combo in "tests/test_graph_gpu.py tests/test_modules_gpu.py" "tests/test_losses_gpu.py tests/test_modules_gpu.py" "tests/test_config1_gpu.py tests/test_modules_gpu.py" "tests/test_modules_gpu.py"; do
  echo "== $combo"
  timeout -s KILL 600 python -m pytest $combo -q -p no:cacheprovider 2>&1 | grep -v "^$" | tail -n 3 | cut -c1-200
done
echo "== graph tests one by one, each followed by test_modules"
for k in "whole]" "segmented]" "2streams" "rsss_step_graph_forms_match_eager[whole" "rsss_step_graph_forms_match_eager[segmented"; do
  echo "-- $k"
  timeout -s KILL 600 python -m pytest tests/test_graph_gpu.py tests/test_modules_gpu.py -q -p no:cacheprovider -k "$k or double_conv_down" 2>&1 | grep -v "^$" | tail -n 2 | cut -c1-200
done
