"""Pins the UMMA descriptor semantics the halo-reuse kernels rely on (run under gpurun):
  * K-major A with a stride-byte-offset that is NOT a multiple of 1024 (8-pixel groups at a 10-pixel row pitch) and an
    arbitrary row shift of the start address;
  * MN-major A with SBO = 1280, a per-UMMA start advance of 2560 bytes and a SMALL leading-byte-offset (the two 64-wide
    halves of M = 128 are two tap-shifted views of the same staged tile)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fcdgan_b200 import _lib
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(3)

def run(A, B, n, ksteps, mn, a_shift=0, a_sbo=0, a_lbo=0, a_kstep=0):
    Ab, Bb = A.to(torch.bfloat16).contiguous().to(dev), B.to(torch.bfloat16).contiguous().to(dev)
    D = torch.full((128, n), float("nan"), device=dev)
    _lib.call_probe("fcd_debug_umma_probe", Ab.data_ptr(), Bb.data_ptr(), D.data_ptr(), A.shape[1], B.shape[1], A.shape[0], B.shape[0],
              mn, n, ksteps, a_shift, 0, 0, 0, a_sbo, 0, a_lbo, a_kstep, None)
    torch.cuda.synchronize()
    return D.double().cpu()

# ---- K-major: A rows = pixels of a halo tile with row pitch P (pixels); M = 128 = 16 image rows x 8 px
for P, shift in ((10, 0), (10, 1), (10, 11), (10, 22), (12, 13)):
    R = 16 * P + 40
    A = torch.randint(-4, 5, (1, R, 64), generator=g).float()
    B = torch.randint(-4, 5, (1, 64, 64), generator=g).float()
    D = run(A, B, 64, 4, 0, a_shift=shift, a_sbo=P * 128)
    rows = torch.tensor([shift + (m // 8) * P + (m % 8) for m in range(128)])
    ref = A[0][rows].double() @ B[0].double().t()
    print(f"K-major  pitch={P} shift={shift}: max|err| = {(D - ref).abs().max().item():.3g}", flush=True)

# ---- MN-major: A rows = pixels (K), columns = 64 channels (M half); two tap views at distance lbo_rows
for P, shift, lbo_rows in ((10, 0, 1), (10, 3, 1), (10, 12, 8), (10, 5, 10), (16, 7, 4), (10, 0, 21)):
    R = 8 * P + 64
    A = torch.randint(-4, 5, (1, R, 64), generator=g).float()
    B = torch.randint(-4, 5, (1, 64, 64), generator=g).float()
    # K = 64 pixels = 8 image rows x 8 px at pitch P; UMMA K = 16 = two image rows -> start advance 2*P*128
    D = run(A, B, 64, 4, 1, a_shift=shift, a_sbo=P * 128, a_lbo=lbo_rows * 128, a_kstep=2 * P * 128)
    krows = torch.tensor([shift + (k // 8) * P + (k % 8) for k in range(64)])
    A0 = A[0][krows]                       # [64 k][64 m]  (first half of M)
    A1 = A[0][krows + lbo_rows]            # second half: the view shifted by lbo_rows pixels
    Am = torch.cat([A0, A1], dim=1).double()        # [k][128]
    ref = Am.t() @ B[0].double()                    # B MN-major: [k rows][64 n]
    print(f"MN-major pitch={P} shift={shift} lbo_rows={lbo_rows}: max|err| = {(D - ref).abs().max().item():.3g}", flush=True)
