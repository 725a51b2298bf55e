#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | grep -v "^$" > gpurun_out/q_tests.log
tail -n 8 gpurun_out/q_tests.log | cut -c1-300
grep -n "^E  .*Error\|^E  .*assert" gpurun_out/q_tests.log | head -30 | cut -c1-500
