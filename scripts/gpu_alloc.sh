#!/bin/bash
timeout -s KILL 200 python scripts/dbg_alloc.py 2>&1 | grep -E "plain|dist" | cut -c1-1500
timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29514 scripts/dbg_alloc.py 2>&1 | grep -E "plain|dist" | cut -c1-1500
