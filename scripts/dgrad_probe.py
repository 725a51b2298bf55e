"""Is the Discriminator's first-iteration gradient reproducible?  Fresh network each run; A: D(x_mask, y_mask), B: D(x_mask, x_mask)
(analytically zero conv-weight gradient: both branches see the same input).  Prints per conv weight the spread of |grad A| over the
runs and the size of |grad B| relative to |grad A|, and lists the outlier runs."""
import sys

import torch

sys.path.insert(0, ".")
import fcdgan_b200 as fb  # noqa: E402
from fcdgan_b200 import engine as E  # noqa: E402
from tests import test_graph_gpu as T  # noqa: E402

runs = int(sys.argv[1]) if len(sys.argv) > 1 else 60
fb.set_precision("parity")
keys = ["net.0.weight", "net.0.bias", "net.2.weight", "net.5.weight", "net.8.weight", "classifier.1.weight"]
for name, bb, im in (("default", True, True), ("two-pass branches", False, True), ("zero-padded first layer", True, False)):
    E.set_batch_branches(bb)
    E.set_im2col(im)
    recA, recB = [], []
    for r in range(runs):
        netG, netD, optG, optD, x, y, cmap, zero = T._setup()
        mx, my, mx2 = fb.soft_mask(x, cmap), fb.soft_mask(y, cmap), fb.soft_mask(x, cmap)
        (-fb.mean(netD(mx, my))).backward()
        gA = {k: p.grad.clone() for k, p in netD.named_parameters()}
        netD.zero_grad(set_to_none=True)
        fb.mean(netD(mx, mx2)).backward()
        gB = {k: p.grad.clone() for k, p in netD.named_parameters()}
        recA.append({k: float(gA[k].double().norm()) for k in keys})
        recB.append({k: float(gB[k].double().norm()) for k in keys})
    print(f"== {name}", flush=True)
    for k in keys:
        a = sorted(v[k] for v in recA)
        med = a[len(a) // 2]
        b = [vb[k] / med for vb in recB]
        outA = [(i, f"{v[k] / med:.4f}") for i, v in enumerate(recA) if abs(v[k] / med - 1) > 1e-3]
        outB = [(i, f"{v:.2e}") for i, v in enumerate(b) if v > 1e-3]
        print(f"  {k:22s} |gA| median {med:.4e} min/max {a[0] / med:.5f} {a[-1] / med:.5f}   |gB|/|gA| median {sorted(b)[len(b) // 2]:.2e} max {max(b):.2e}"
              f"   outliers A {outA[:6]} B {outB[:6]}", flush=True)
E.set_batch_branches(True)
E.set_im2col(True)
