"""GPU parity of the CUDA networks against golden vectors produced by the UNMODIFIED reference
(oracle/make_golden.py): outputs, input gradients, parameter gradients and BatchNorm running statistics.

Tolerances (parity precision = split-bf16 operands, fp32 accumulate): outputs 1e-3 relative (the north-star
bar on the change-density map; measured error is ~1e-5), gradients 2e-3 of the tensor's max."""
import pytest
import torch

import fcdgan_b200 as fb
from oracle import fcd_oracle as O
from tests._util import check_grad_summary, load_golden, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
OUT_TOL = 1e-3
GRAD_TOL = 2e-3


def _load(net, spec, seed):
    sd = O.make_state_dict(spec, seed)
    net.load_state_dict(sd)
    return net.to(DEV)


def _named_grads(net):
    return {k: p.grad for k, p in net.named_parameters()}


def _check_running(net, f):
    sd = net.state_dict()
    for k, v in f["running"].items():
        assert rel_err(sd[k].float(), v.float()) < 1e-4, k


@pytest.mark.parametrize("name", ["g13_train.pt", "g4_eval.pt"])
def test_generator(name):
    fb.set_precision("parity")
    f = load_golden(name)
    net = _load(fb.Generator(f["C"]), O.generator_spec(f["C"]), f["seed"])
    net.train(f["train"])
    x = f["x"].to(DEV).requires_grad_(True)
    y = net(x)
    assert rel_err(y, f["y"]) < OUT_TOL
    (y * f["r"].to(DEV)).sum().backward()
    assert rel_err(x.grad, f["dx"]) < GRAD_TOL
    check_grad_summary(_named_grads(net), f["grads"], GRAD_TOL, what=name)
    _check_running(net, f)


@pytest.mark.parametrize("name", ["s13_bilinear_even.pt", "s4_bilinear_odd.pt", "s4_convT_odd.pt", "s4_bilinear_eval.pt"])
def test_segmentor(name):
    fb.set_precision("parity")
    f = load_golden(name)
    net = _load(fb.Segmentor(f["C"], 1, f["bilinear"]), O.segmentor_spec(f["C"], 1, f["bilinear"]), f["seed"])
    net.train(f["train"])
    x = f["x"].to(DEV).requires_grad_(True)
    y = f["y"].to(DEV).requires_grad_(True)
    cmap = net(x, y)
    assert rel_err(cmap, f["cmap"]) < OUT_TOL
    (cmap * f["r"].to(DEV)).sum().backward()
    assert rel_err(x.grad, f["dx"]) < GRAD_TOL and rel_err(y.grad, f["dy"]) < GRAD_TOL
    check_grad_summary(_named_grads(net), f["grads"], GRAD_TOL, what=name)
    _check_running(net, f)


@pytest.mark.parametrize("name", ["d13.pt", "d3_odd.pt"])
def test_discriminator(name):
    fb.set_precision("parity")
    f = load_golden(name)
    net = _load(fb.Discriminator_SRGAN_simple(f["C"]), O.discriminator_spec(f["C"]), f["seed"])
    net.train(True)
    x = f["x"].to(DEV).requires_grad_(True)
    y = f["y"].to(DEV).requires_grad_(True)
    out = net(x, y)
    assert out.shape == f["out"].shape
    assert rel_err(out, f["out"]) < OUT_TOL
    (out * f["r"].to(DEV)).sum().backward()
    assert rel_err(x.grad, f["dx"]) < GRAD_TOL and rel_err(y.grad, f["dy"]) < GRAD_TOL
    check_grad_summary(_named_grads(net), f["grads"], GRAD_TOL, what=name)
    _check_running(net, f)


def test_generator_double_backward_and_fast_mode():
    """retain_graph=True + second backward (Demo_USSS.py:327,338) accumulates 2x the gradient; 'fast' precision
    stays within bf16-class error of the reference."""
    f = load_golden("g13_train.pt")
    fb.set_precision("parity")
    net = _load(fb.Generator(13), O.generator_spec(13), f["seed"]).train()
    x = f["x"].to(DEV)
    y = net(x)
    loss = (y * f["r"].to(DEV)).sum()
    loss.backward(retain_graph=True)
    g1 = {k: p.grad.clone() for k, p in net.named_parameters()}
    loss.backward()
    for k, p in net.named_parameters():
        assert rel_err(p.grad, 2 * g1[k]) < 1e-5, k
    fb.set_precision("fast")
    try:
        net2 = _load(fb.Generator(13), O.generator_spec(13), f["seed"]).train()
        y2 = net2(x)
        assert rel_err(y2, f["y"]) < 5e-2
    finally:
        fb.set_precision("parity")


def test_no_grad_inference():
    f = load_golden("s4_bilinear_eval.pt")
    net = _load(fb.Segmentor(f["C"], 1, True), O.segmentor_spec(f["C"], 1, True), f["seed"]).eval()
    with torch.no_grad():
        cmap = net(f["x"].to(DEV), f["y"].to(DEV))
    assert not cmap.requires_grad and rel_err(cmap, f["cmap"]) < OUT_TOL
