"""Where does an odd EAGER run of the G + D iteration first leave the pack?  Per iteration: D's gradients (relative L2 distance and
sign-flip fraction w.r.t. run 0), D's parameters after the optimizer step, RMSprop's square_avg."""
import sys

import torch

sys.path.insert(0, ".")
import fcdgan_b200 as fb  # noqa: E402
from fcdgan_b200 import engine as E  # noqa: E402
from tests import test_graph_gpu as T  # noqa: E402

runs = int(sys.argv[1]) if len(sys.argv) > 1 else 60
fb.set_precision("parity")
KEYS = ("net.0.weight", "net.2.weight", "net.8.weight", "classifier.1.weight", "classifier.1.bias", "classifier.3.weight", "classifier.3.bias", "net.9.weight", "net.9.bias")


def one():
    netG, netD, optG, optD, x, y, cmap, zero = T._setup()
    seg_g, seg_d, seg_opt = T._segments(netG, netD, optG, optD, zero)
    rec = []
    for it in range(3):
        seg_g(x, y, cmap)
        dl = seg_d(x, y, cmap)
        g = {k: p.grad.detach().clone() for k, p in netD.named_parameters()}
        seg_opt(x, y, cmap)
        p = {k: v.detach().clone() for k, v in netD.named_parameters()}
        rec.append((dl.item(), g, p))
    return rec


for name, bb, pool in (("two-pass branches, zero pool", False, True), ("two-pass branches, one torch.zeros per accumulator", False, False),
                       ("default, one torch.zeros per accumulator", True, False)):
    E.set_batch_branches(bb)
    E._cfg["zero_pool"] = pool
    ref = one()
    nodd = 0
    for r in range(1, runs):
        cur = one()
        fin = cur[2][2]["net.2.weight"], ref[2][2]["net.2.weight"]
        frac = float(((fin[0] - fin[1]).abs() > 1e-6 + 1e-4 * fin[1].abs()).float().mean())
        if frac < 0.1:
            continue
        nodd += 1
        if nodd > 1:
            continue
        print(f"-- {name}: odd run {r} (net.2.weight off fraction {frac:.2f})")
        for it in range(3):
            line = [f"it{it} d_loss {cur[it][0] - ref[it][0]:+.1e}"]
            for k in KEYS:
                ga, gb = cur[it][1][k].double(), ref[it][1][k].double()
                rel = float((ga - gb).norm() / gb.norm().clamp_min(1e-30))
                flips = float((torch.sign(ga) != torch.sign(gb)).float().mean())
                pa, pb = cur[it][2][k], ref[it][2][k]
                line.append(f"{k}: g rel {rel:.1e} flips {flips:.3f} p maxdiff {float((pa - pb).abs().max()):.1e}")
            print("   " + "\n      ".join(line), flush=True)
    print(f"== {name}: {nodd} odd of {runs - 1}", flush=True)
E.set_batch_branches(True)
E._cfg['zero_pool'] = True
