"""Operand-reuse experiments for tcgen05.mma (SS mode, K-major, M = 128): cycles per k-step for the compile-time UMMA
'programs' of csrc/probe_tc.cu (fully unrolled issue loop, ~2 instructions per UMMA)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fcdgan_b200 import _lib
dev = torch.device("cuda:0")
ITERS = 2000
PROGS = [
    ("1x  A0 x B0 (N=128)", 64),
    ("2x  same B, same accumulator (N=128)", 128),
    ("2x  same B, separate accumulators (N=128)", 128),
    ("4x  same B, A0..A3 (N=128)", 256),
    ("2x  same A, B0 / B1 (N=128)", 128),
    ("2x  nothing shared (N=128)", 128),
    ("split pair: hi*[hi;lo] (128) + lo*hi (64), same acc", 96),
    ("split pair, lo*hi into a separate accumulator", 96),
    ("two tiles: hi0 hi1 (128) lo0 lo1 (64), same B", 192),
    ("two tiles: hi0 lo0 hi1 lo1", 192),
    ("four tiles: 4 x hi (128) then 4 x lo (64), same B", 384),
    ("1x  N=256", 128),
    ("2x  N=256 same B", 256),
    ("3x  N=64 (unstacked split: hi*hi, lo*hi, hi*lo)", 96),
    ("4x  N=64 same B, A0..A3", 128),
    ("1x  N=64", 32),
    ("4x  N=256 same B", 512),
    ("2x  N=64 nothing shared", 64),
    ("4x  N=128 nothing shared", 256),
    ("2x  N=192 same B", 192),
]
for grid in (1,):
    for i, (name, ideal) in enumerate(PROGS):
        cyc = torch.zeros(grid, dtype=torch.int64, device=dev)
        _lib.call_probe("fcd_debug_umma_prog", i, ITERS, 1024, grid, cyc.data_ptr(), None)
        torch.cuda.synchronize()
        per = cyc.double().mean().item() / (ITERS * 4)
        print(f"{i:2d} {name:58s} {per:7.1f} cyc/k-step   ideal {ideal:4d}  -> {ideal / per * 100:5.1f} %", flush=True)
