#!/bin/bash
export FCD_DIST_TIMEOUT_S=90
nproc
timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 scripts/dbg_multi2.py 2>&1 | grep "rank"
