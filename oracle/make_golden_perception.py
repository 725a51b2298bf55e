"""Golden vectors for the perception loss from the UNMODIFIED reference — TEST INFRASTRUCTURE ONLY.

Runs `Loss.PerceptionLoss`, `Loss.CNetLoss` and `Loss.CGeneratorLoss` of /root/reference (imported through oracle/ref_import.py,
whose `Loss.vgg16` rebind builds torchvision's VGG16 under seed 1234 — the ImageNet weights of Loss.py:25 cannot be downloaded
here) and records values + gradients w.r.t. the generated image and the change-density map.  The VGG weights themselves
(58 MB) are NOT stored: every consumer rebuilds them with `oracle.fcd_oracle.vgg16_features(1234)`; the fixture records a
checksum of the weights so that a torch / torchvision whose initialiser differs is detected instead of mis-reported.

    python oracle/make_golden_perception.py      # -> tests/golden/perception.pt
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import fcd_oracle as O, ref_import  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "perception.pt")


def vgg_checksum(net):
    return float(sum(p.double().abs().sum() for p in net.parameters()))


def main():
    assert ref_import.available(), "needs /root/reference"
    torch.set_num_threads(8)
    M, L, _ = ref_import.load()
    fix = {}
    g = torch.Generator().manual_seed(61)
    # (a) per-band, 4 bands, one feature layer (the demos' setting: Demo_USSS.py:110, Demo_RSSS.py:160), odd-ish size
    B, C, H, W = 2, 4, 72, 56
    t = torch.randn(B, C, H, W, generator=g)
    gen = (t + 0.5 * torch.randn(B, C, H, W, generator=g)).requires_grad_(True)
    cmap = torch.rand(B, 1, H, W, generator=g).requires_grad_(True)
    pl = L.PerceptionLoss(feature_layer=1, perception_perBand=True)
    fix["vgg_checksum"] = vgg_checksum(pl.net)
    assert abs(fix["vgg_checksum"] - vgg_checksum(O.vgg16_features(1234))) < 1e-6 * fix["vgg_checksum"]
    v = pl(t, gen, cmap)
    v.backward()
    fix["perband"] = {"t": t, "g": gen.detach().clone(), "cmap": cmap.detach().clone(), "value": v.item(),
                      "dg": gen.grad.clone(), "dcmap": cmap.grad.clone()}
    # (b) RGB mode (first three bands), all five feature layers
    B, C, H, W = 2, 5, 64, 80
    t = torch.randn(B, C, H, W, generator=g)
    gen = (t + 0.5 * torch.randn(B, C, H, W, generator=g)).requires_grad_(True)
    cmap = torch.rand(B, 1, H, W, generator=g).requires_grad_(True)
    pl5 = L.PerceptionLoss(feature_layer=5, perception_perBand=False)
    v = pl5(t, gen, cmap)
    v.backward()
    fix["rgb5"] = {"t": t, "g": gen.detach().clone(), "cmap": cmap.detach().clone(), "value": v.item(),
                   "dg": gen.grad.clone(), "dcmap": cmap.grad.clone()}
    # (c) CNetLoss / CGeneratorLoss with a live perception weight (USSS 0.4, Demo_USSS.py:40; WSSS 0.5, Demo_WSSS.py:43)
    B, C, H, W = 1, 3, 168, 164
    t = torch.randn(B, C, H, W, generator=g)
    gen = (t + 0.4 * torch.randn(B, C, H, W, generator=g)).requires_grad_(True)
    cmap = torch.rand(B, 1, H, W, generator=g).requires_grad_(True)
    crit = L.CNetLoss(channel=C, perception_layer=1, perception_perBand=True)
    gl, l1, perc, sl = crit(t, gen, cmap)
    (gl + 0.65 * l1 + 0.4 * perc + 0.3 * sl).backward()
    fix["cnet"] = {"t": t, "g": gen.detach().clone(), "cmap": cmap.detach().clone(),
                   "values": (gl.item(), l1.item(), perc.item(), sl.item()), "dg": gen.grad.clone(), "dcmap": cmap.grad.clone()}
    gen2 = gen.detach().clone().requires_grad_(True)
    cmap2 = cmap.detach().clone().requires_grad_(True)
    crit2 = L.CGeneratorLoss(channel=C, perception_layer=2, perception_perBand=False)
    gl2, sl2, perc2 = crit2(t, gen2, cmap2)
    (gl2 + 0.5 * perc2).backward()
    fix["cgen"] = {"values": (gl2.item(), sl2.item(), perc2.item()), "dg": gen2.grad.clone(), "dcmap": cmap2.grad.clone()}
    # (d) hard mask (generator_mask_switch=True, Loss.py:89-90): no gradient reaches cmap through the perception term
    gen3 = gen.detach().clone().requires_grad_(True)
    _, _, perc3, _ = crit(t, gen3, cmap.detach(), generator_mask_switch=True)
    perc3.backward()
    fix["cnet_hard"] = {"value": perc3.item(), "dg": gen3.grad.clone()}
    torch.save(fix, OUT)
    print("perception goldens:", {k: (v["value"] if isinstance(v, dict) and "value" in v else (v["values"] if isinstance(v, dict) else v))
                                   for k, v in fix.items()})


if __name__ == "__main__":
    main()
