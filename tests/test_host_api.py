"""CPU: the host-side mirror of the reference interface — class names, constructor signatures, state_dict keys and
shapes (checked against the oracle's restatement of the reference layouts AND, where /root/reference is mounted,
against the real reference modules), error behaviour without a GPU."""
import inspect

import pytest
import torch

import fcdgan_b200 as fb
from fcdgan_b200._lib import FcdError
from oracle import fcd_oracle as O
from oracle import ref_import


def _shapes(sd):
    return {k: tuple(v.shape) for k, v in sd.items()}


@pytest.mark.parametrize("C", [3, 4, 13])
def test_state_dict_layouts_match_reference_spec(C):
    assert _shapes(fb.Generator(C).state_dict()) == _shapes(O.make_state_dict(O.generator_spec(C), 0))
    assert _shapes(fb.Discriminator_SRGAN_simple(C).state_dict()) == _shapes(O.make_state_dict(O.discriminator_spec(C), 0))
    for bil in (True, False):
        assert _shapes(fb.Segmentor(C, 1, bil).state_dict()) == _shapes(O.make_state_dict(O.segmentor_spec(C, 1, bil), 0))
    assert len(fb.Generator(13).state_dict()) == 87 and len(fb.Discriminator_SRGAN_simple(13).state_dict()) == 27
    assert len(fb.Segmentor(13, 1, True).state_dict()) == 128
    assert sum(p.numel() for p in fb.Generator(13).parameters()) == 542483
    assert sum(p.numel() for p in fb.Segmentor(13, 1, True).parameters()) == 40833729
    assert sum(p.numel() for p in fb.Discriminator_SRGAN_simple(13).parameters()) == 2084865


@pytest.mark.skipif(not ref_import.available(), reason="reference not mounted (build container only)")
def test_state_dicts_interchange_with_the_real_reference():
    M, L, S = ref_import.load()
    for ours, ref in ((fb.Generator(4), M.Generator(4)), (fb.Segmentor(4, 1, True), M.Segmentor(4, 1, True)),
                      (fb.Segmentor(4, 1, False), M.Segmentor(4, 1, False)),
                      (fb.Discriminator_SRGAN_simple(4), M.Discriminator_SRGAN_simple(4))):
        assert _shapes(ours.state_dict()) == _shapes(ref.state_dict())
        ours.load_state_dict(ref.state_dict())          # reference .pkl -> ours
        ref.load_state_dict(ours.state_dict())          # ours -> reference
    # constructor / forward signatures
    for name in ("Generator", "Segmentor", "Discriminator_SRGAN_simple", "DoubleConv", "Down", "Up", "OutConv", "ResidualBlock"):
        a = inspect.signature(getattr(M, name).__init__)
        b = inspect.signature(getattr(fb, name).__init__)
        assert list(a.parameters) == list(b.parameters), name
        assert [p.default for p in a.parameters.values()] == [p.default for p in b.parameters.values()], name
        fa = list(inspect.signature(getattr(M, name).forward).parameters)
        fbb = list(inspect.signature(getattr(fb, name).forward).parameters)
        assert fa == fbb, name
    for name in ("SSIM", "MS_SSIM"):
        a, b = inspect.signature(getattr(S, name).__init__), inspect.signature(getattr(fb, name).__init__)
        assert list(a.parameters) == list(b.parameters)
        assert [p.default for p in a.parameters.values()] == [p.default for p in b.parameters.values()]
    assert list(inspect.signature(S.ms_ssim).parameters) == list(inspect.signature(fb.ms_ssim).parameters)
    assert list(inspect.signature(S.ssim).parameters) == list(inspect.signature(fb.ssim).parameters)
    assert list(inspect.signature(L.region_loss).parameters) == list(inspect.signature(fb.region_loss).parameters)
    for name in ("CNetLoss", "CGeneratorLoss"):
        a = list(inspect.signature(getattr(L, name).forward).parameters)
        b = list(inspect.signature(getattr(fb, name).forward).parameters)
        assert a == b, name
    # the Gaussian window and default MS-SSIM weights are the reference's
    import sys
    assert torch.equal(S._fspecial_gauss_1d(11, 1.5), sys.modules["fcdgan_b200.ssim"]._fspecial_gauss_1d(11, 1.5))


def test_gauss_window_matches_oracle():
    import sys
    ssim_mod = sys.modules["fcdgan_b200.ssim"]      # (the package attribute `ssim` is the function, like the reference's)
    assert torch.equal(ssim_mod._fspecial_gauss_1d(11, 1.5).flatten(), O.gauss_window(11, 1.5))
    m = fb.MS_SSIM(data_range=1.0, channel=13)
    assert tuple(m.win.shape) == (13, 1, 1, 11) and m.data_range == 1.0 and m.weights is None


def test_no_cpu_fallback_and_input_validation():
    g = fb.Generator(3)
    with pytest.raises(FcdError):
        g(torch.rand(1, 3, 16, 16))                     # CPU tensor: loud failure, no fallback
    with pytest.raises(ValueError):
        g(torch.rand(1, 4, 16, 16))                     # wrong band count
    s = fb.Segmentor(3, 1, True)
    with pytest.raises(ValueError):
        s(torch.rand(1, 3, 32, 32), torch.rand(1, 3, 32, 16))
    with pytest.raises(ValueError):
        s(torch.rand(1, 3, 8, 8), torch.rand(1, 3, 8, 8))
    with pytest.raises(ValueError):
        fb.ms_ssim(torch.rand(1, 3, 200, 200), torch.rand(1, 2, 200, 200))
    with pytest.raises(AssertionError):
        fb.ms_ssim(torch.rand(1, 3, 160, 160), torch.rand(1, 3, 160, 160))
    with pytest.raises(ValueError):
        fb.ssim(torch.rand(1, 3, 32, 32), torch.rand(1, 3, 32, 32), win_size=8)
    with pytest.raises(NotImplementedError):
        fb.region_loss(torch.rand(1, 1, 8, 8), torch.rand(1, 1, 8, 8), torch.nn.BCELoss())
    with pytest.raises(ValueError):
        fb.set_precision("fp8")


def test_train_eval_and_optimizer_plumbing():
    g = fb.Generator(4)
    assert g.training and not g.eval().training and g.train().training
    opt = torch.optim.Adam(g.parameters(), lr=2e-4, betas=(0.9, 0.99))   # Demo_USSS.py:121
    assert len(opt.param_groups[0]["params"]) == len(list(g.parameters()))
    assert float(g.block1[1].weight) == 0.25                             # nn.PReLU() default slope, Module.py:147
