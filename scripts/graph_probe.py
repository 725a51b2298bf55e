"""Diagnostic for tests/test_graph_gpu.py::test_graph_replay_matches_eager: repeat the eager-vs-graph comparison and print, per
repetition, the loss differences and the fraction of differing elements of the worst parameter tensors (eager vs eager too)."""
import sys

import torch

sys.path.insert(0, ".")
import fcdgan_b200 as fb  # noqa: E402
from fcdgan_b200 import engine as E  # noqa: E402
from fcdgan_b200.graph import GraphedStep  # noqa: E402
from tests import test_graph_gpu as T  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
fb.set_precision("parity")


def off_fracs(got, want):
    out = []
    for (k, a), (_, b) in zip(got, want):
        d = (a - b).abs()
        out.append((float((d > 1e-6 + 1e-4 * b.abs()).float().mean()), float(d.max()), k))
    return sorted(out, reverse=True)[:4]


for rep in range(reps):
    l0, p0 = T._eager_run(3)
    l1, p1 = T._eager_run(3)
    netG, netD, optG, optD, x, y, cmap, zero = T._setup()
    seg_g, seg_d, seg_opt = T._segments(netG, netD, optG, optD, zero)

    def whole(x, y, cmap):
        gl, dl = seg_g(x, y, cmap), seg_d(x, y, cmap)
        seg_opt(x, y, cmap)
        return gl, dl

    step = GraphedStep(whole, [x, y, cmap], warmup=2, modules=[netG, netD], optimizers=[optG, optD], restore_after_warmup=True)
    lg = []
    for i in range(3):
        gl, dl = step()
        lg.append((gl.item(), dl.item()))
    pg = T._params(netG, netD)
    fmt = lambda ls: " ".join(f"{a:.7f}/{b:.7f}" for a, b in ls)
    print(f"rep {rep}\n  eager  {fmt(l0)}\n  eager2 {fmt(l1)}\n  graph  {fmt(lg)}")
    print("  eager2 vs eager:", [(f"{f:.3f}", f"{m:.2e}", k) for f, m, k in off_fracs(p1, p0)])
    print("  graph  vs eager:", [(f"{f:.3f}", f"{m:.2e}", k) for f, m, k in off_fracs(pg, p0)], flush=True)
    del step
    E.clear_caches()
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
