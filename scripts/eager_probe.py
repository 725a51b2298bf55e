"""How often does an EAGER run of the G + D iteration of tests/test_graph_gpu.py leave the pack (D parameters deviating in most
elements from the modal result), per engine switch setting?  python scripts/eager_probe.py <runs>"""
import sys

import torch

sys.path.insert(0, ".")
import fcdgan_b200 as fb  # noqa: E402
from fcdgan_b200 import engine as E  # noqa: E402
from tests import test_graph_gpu as T  # noqa: E402

runs = int(sys.argv[1]) if len(sys.argv) > 1 else 40
fb.set_precision("parity")
for name, bb, im in (("default", True, True), ("two-pass branches", False, True), ("batched, zero-padded first layer", True, False)):
    E.set_batch_branches(bb)
    E.set_im2col(im)
    ref = None
    odd = []
    for r in range(runs):
        losses, params = T._eager_run(3)
        p = dict(params)
        if ref is None:
            ref, ref_l = p, losses
            continue
        frac = {k: float(((p[k] - ref[k]).abs() > 1e-6 + 1e-4 * ref[k].abs()).float().mean()) for k in p}
        worst = max((v, k) for k, v in frac.items() if ref[k].numel() >= 4096)
        if worst[0] > 0.1:
            odd.append((r, [f"{k}={frac[k]:.2f}" for k in ("net.0.weight", "net.2.weight", "net.5.weight", "net.8.weight", "net.9.weight",
                                                           "classifier.1.weight", "classifier.3.weight", "block2.conv1.weight")],
                        [f"{a[1] - b[1]:+.1e}" for a, b in zip(losses, ref_l)]))
    print(f"{name}: {len(odd)} of {runs - 1} runs deviate", flush=True)
    for o in odd[:6] + ([("...",)] if len(odd) > 12 else []) + odd[6:][-6:]:
        print("   ", o, flush=True)
E.set_batch_branches(True)
E.set_im2col(True)
