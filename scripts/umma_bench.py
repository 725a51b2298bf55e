"""tcgen05.mma issue-rate micro-benchmark (run under gpurun): cycles per k-step for the operand shapes / layouts the conv
kernels use, operands resident in shared memory (no TMA in the loop).  Ideal: M128 x N x K16 bf16 = N/2 cycles."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fcdgan_b200 import _lib
dev = torch.device("cuda:0")
ITERS = 2000

def run(name, grid=1, mn=0, n1=64, n2=0, a_sbo=1024, a_lbo=16, a_kstep=32, a_shift=0, a2_off=16384, b_sbo=1024, b_lbo=16,
        b_kstep=32, stage_stride=49152, stages=4, b_off=32768, a_tmem=0):
    cyc = torch.zeros(grid, dtype=torch.int64, device=dev)
    _lib.call_probe("fcd_debug_umma_bench", mn, n1, n2, a_sbo, a_lbo, a_kstep, a_shift, a2_off, b_sbo, b_lbo, b_kstep, stage_stride,
              stages, b_off, ITERS, a_tmem, grid, cyc.data_ptr(), None)
    torch.cuda.synchronize()
    c = cyc.cpu().double()
    per = c / (ITERS * 4)
    ideal = (n1 + n2) / 2
    print(f"{name:58s} grid={grid:3d} cyc/k-step: min {per.min():7.1f} mean {per.mean():7.1f} max {per.max():7.1f}   ideal {ideal:5.0f}"
          f"  -> {ideal / per.mean() * 100:5.1f} % of peak", flush=True)

for grid in (1, 148):
    # K-major, aligned tiles (A 16 KB hi | 16 KB lo | B 16 KB), as conv_tc_kernel
    run("K-major N=64", grid, n1=64)
    run("K-major N=128", grid, n1=128)
    run("K-major N=256", grid, n1=256, stage_stride=65536, stages=3, b_off=32768)
    run("K-major N=128 + N=64 (split conv: hi*[hi;lo], lo*hi)", grid, n1=128, n2=64)
    run("K-major N=64 x3 (n1=64,n2=64 ~ 2 of 3)", grid, n1=64, n2=64)
    # K-major halo views: 8-px groups at 10-px pitch, shifted start
    run("K-major halo pitch 10 shift 11, N=128 + N=64", grid, n1=128, n2=64, a_sbo=1280, a_shift=11, a2_off=23552,
        stage_stride=65536, stages=3, b_off=49152)
    run("K-major halo pitch 10 shift 0, N=128 + N=64", grid, n1=128, n2=64, a_sbo=1280, a_shift=0, a2_off=23552,
        stage_stride=65536, stages=3, b_off=49152)
    run("K-major halo pitch 10 shift 11, N=256", grid, n1=256, a_sbo=1280, a_shift=11, stage_stride=65536, stages=3, b_off=32768)
    # MN-major (wgrad): A = two 64-ch blocks [64 px][128 B] at LBO 8192, K step 2048
    run("MN-major N=64", grid, mn=1, n1=64, a_lbo=8192, a_kstep=2048, b_lbo=8192, b_kstep=2048)
    run("MN-major N=128", grid, mn=1, n1=128, a_lbo=8192, a_kstep=2048, b_lbo=8192, b_kstep=2048)
    run("MN-major N=128 + N=64 (split wgrad)", grid, mn=1, n1=128, n2=64, a_lbo=8192, a_kstep=2048, b_lbo=8192, b_kstep=2048)
    run("MN-major halo pitch 10 (sbo 1280, lbo 128, kstep 2560) N=128+64", grid, mn=1, n1=128, n2=64, a_sbo=1280, a_lbo=128,
        a_kstep=2560, a_shift=3, a2_off=13312, b_lbo=8192, b_kstep=2048, stage_stride=49152, stages=4, b_off=32768)
    run("MN-major N=256", grid, mn=1, n1=256, a_lbo=8192, a_kstep=2048, b_lbo=8192, b_kstep=2048, stage_stride=65536, stages=3,
        b_off=16384)
    # A from TMEM
    run("A in TMEM, B K-major N=64", grid, n1=64, a_tmem=1)
    run("A in TMEM, B K-major N=128", grid, n1=128, a_tmem=1)
    run("A in TMEM, B K-major N=256", grid, n1=256, a_tmem=1, stage_stride=65536, stages=3, b_off=32768)
