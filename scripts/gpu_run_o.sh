#!/bin/bash
# full -m gpu suite + smoke + default bench exactly as the driver runs it (+ the reference arm)
mkdir -p gpurun_out
timeout -s KILL 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -n 15 | cut -c1-300
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2 | cut -c1-300
timeout -s KILL 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/o_bench.json 2> gpurun_out/o_bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/o_bench.json
timeout -s KILL 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/o_bench_ref.json 2> gpurun_out/o_bench_ref.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/o_bench_ref.json
