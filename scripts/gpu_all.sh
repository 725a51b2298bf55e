#!/bin/bash
bash scripts/gpu_tests.sh 40
bash scripts/gpu_bench.sh
