"""Live pin of the CPU oracle port (oracle/fcd_oracle.py) against the UNMODIFIED reference classes on FRESH seeds (the golden
files pin it at fixed ones).  Runs wherever the reference can be imported: the mounted /root/reference (build container) or the
byte-identical staged copy oracle/_ref; skipped otherwise.  CPU only."""
import pytest
import torch

from oracle import fcd_oracle as O
from oracle import ref_import
from tests._util import rel_err

pytestmark = pytest.mark.skipif(not ref_import.importable(), reason="reference neither mounted nor staged (oracle/build_ref.py)")
TOL = 2e-5


@pytest.fixture(scope="module")
def ref():
    torch.set_num_threads(8)
    return ref_import.load()


def _pair(B, C, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, C, H, W, generator=g)
    return x, x + 0.3 * torch.randn(B, C, H, W, generator=g)


@pytest.mark.parametrize("seed", [101, 202])
def test_networks_forward_and_gradients(ref, seed):
    M, _, _ = ref
    C, B, H, W = 5, 2, 36, 44
    x, y = _pair(B, C, H, W, seed)
    for kind, make, spec, run in (
            ("G", lambda: M.Generator(C), O.generator_spec(C), lambda sd, t: O.generator(sd, x, train=t)),
            ("S", lambda: M.Segmentor(C, 1, True), O.segmentor_spec(C, 1, True), lambda sd, t: O.segmentor(sd, x, y, True, t)),
            ("D", lambda: M.Discriminator_SRGAN_simple(C), O.discriminator_spec(C), lambda sd, t: O.discriminator(sd, x, y, t))):
        sd0 = O.make_state_dict(spec, seed)
        net = make()
        net.load_state_dict(sd0)
        net.train()
        out = net(x) if kind == "G" else net(x, y)
        r = torch.randn(out.shape, generator=torch.Generator().manual_seed(seed + 1))
        (out * r).sum().backward()
        sd = O.clone_sd(sd0, requires_grad=True)
        out_o = run(sd, True)
        (out_o * r).sum().backward()
        assert rel_err(out_o, out) < TOL, kind
        for k, p in net.named_parameters():
            if p.grad.abs().max() < 1e-6:
                continue
            assert rel_err(sd[k].grad, p.grad) < 2e-3, (kind, k)     # fp32 summation-order noise through BatchNorm over 2x2 maps
        for k, v in net.state_dict().items():
            if "running" in k:
                assert rel_err(sd[k], v) < TOL, (kind, k)


def test_loss_stack_and_perception(ref):
    _, L, _ = ref
    B, C, H, W = 1, 3, 164, 172
    t, gen0 = _pair(B, C, H, W, 303)
    cm0 = torch.rand(B, 1, H, W, generator=torch.Generator().manual_seed(304))
    crit = L.CNetLoss(channel=C, perception_layer=2, perception_perBand=True)
    gen = gen0.clone().requires_grad_(True)
    cmap = cm0.clone().requires_grad_(True)
    gl, l1, perc, sl = crit(t, gen, cmap)
    (gl + 0.65 * l1 + 0.4 * perc + 0.3 * sl).backward()
    g2 = gen0.clone().requires_grad_(True)
    c2 = cm0.clone().requires_grad_(True)
    gl_o, l1_o, sl_o = O.cnet_loss(t, g2, c2)
    vsd = dict(O.vgg16_features(1234).state_dict())
    perc_o = O.perception_loss(vsd, t, g2, c2, 2, True)
    (gl_o + 0.65 * l1_o + 0.4 * perc_o + 0.3 * sl_o).backward()
    for a, b in ((gl_o, gl), (l1_o, l1), (perc_o, perc), (sl_o, sl)):
        assert abs(a.item() - b.item()) <= 1e-4 * max(abs(b.item()), 1e-6)
    assert rel_err(g2.grad, gen.grad) < 1e-4 and rel_err(c2.grad, cmap.grad) < 1e-4
    region = (torch.rand(B, 1, H, W, generator=torch.Generator().manual_seed(305)) > 0.6).float()
    for kind, crit_ in (("l1", torch.nn.L1Loss()), ("mse", torch.nn.MSELoss())):
        assert abs(O.region_loss(cm0, region, kind).item() - L.region_loss(cm0, region, crit_).item()) < TOL
