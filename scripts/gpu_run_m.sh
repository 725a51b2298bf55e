#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_ops_gpu.py tests/test_networks_gpu.py -q -p no:cacheprovider 2>&1 | tail -n 6 | cut -c1-300
timeout -s KILL 300 python scripts/bench_conv_variants.py 2>&1 | grep -v Warn | grep "x_slots=auto" | grep "tma_out=1" | cut -c1-200
timeout -s KILL 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline --profile-out gpurun_out/m_table_c2.txt > gpurun_out/m_c2.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/m_c2.json')); print(d['value'], d['ms_per_step'], d['roofline']['kernel'], d['roofline']['frac'], d['clocks'])"
head -8 gpurun_out/m_table_c2.txt
timeout -s KILL 400 python bench.py --config g32 --precision fast --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/m_g32_fast.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/m_g32_fast.json')); print('g32 fast', d['value'], d['ms_per_step'], d['roofline']['step_frac_of_peak'])"
