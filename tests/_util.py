"""Shared helpers for the parity tests."""
import os

import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name), map_location="cpu", weights_only=False)


def rel_err(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def rel_l2(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def check_grad_summary(named_grads, summary, tol, atol=1e-6, what=""):
    """named_grads: dict name -> grad tensor.  summary: oracle/make_golden.grad_summary output."""
    worst = 0.0
    for k, ref in summary.items():
        assert k in named_grads and named_grads[k] is not None, f"{what}: missing grad for {k}"
        g = named_grads[k].detach().float().cpu()
        if "full" in ref:
            r = ref["full"]
            scale = max(r.abs().max().item(), atol / tol)
            err = (g - r).abs().max().item() / scale
        else:
            scale = max(ref["norm"] / (ref["numel"] ** 0.5), atol / tol)
            err = (g.flatten()[:16] - ref["head"]).abs().max().item() / max(ref["head"].abs().max().item(), scale)
            err = max(err, abs(g.norm().item() - ref["norm"]) / max(ref["norm"], atol / tol))
        worst = max(worst, err)
        assert err <= tol, f"{what}: grad {k} rel err {err:.3g} > {tol}"
    return worst
