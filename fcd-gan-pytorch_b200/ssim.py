"""SSIM / MS-SSIM with the reference's call surface (ssim.py:95-311, vendored pytorch-msssim) on fused sm_100a
kernels (csrc/losses.cu): one launch per pyramid level instead of ~30, Gaussian window uploaded once.

    ssim(X, Y, data_range=255, size_average=True, win_size=11, win_sigma=1.5, win=None, K=(0.01, 0.03),
         nonnegative_ssim=False)
    ms_ssim(X, Y, data_range=255, size_average=True, win_size=11, win_sigma=1.5, win=None, weights=None, K=(0.01, 0.03))
    SSIM(...), MS_SSIM(...)  — same constructor arguments and attributes as the reference classes.

Differences (documented, checked): only 4-d (N, C, H, W) CUDA fp32 inputs (the reference's 5-d conv3d branch is not
on the FCD-GAN path); win_size <= 11; one window shared by all channels (the reference's is too: it repeats one
1-D kernel).  Error behaviour follows ssim.py:120-137,182-197 (ValueError / AssertionError, same messages).
"""
from __future__ import annotations

import warnings
from typing import Dict, Optional, Tuple

import torch

from . import _lib
from .engine import _call

_MS_WEIGHTS = [0.0448, 0.2856, 0.3001, 0.2363, 0.1333]   # ssim.py:199


def _fspecial_gauss_1d(size, sigma):
    """1-D Gaussian, built in fp32 exactly like ssim.py:9-23 -> shape (1, 1, size)."""
    coords = torch.arange(size).to(dtype=torch.float)
    coords -= size // 2
    g = torch.exp(-(coords ** 2) / (2 * sigma ** 2))
    g /= g.sum()
    return g.unsqueeze(0).unsqueeze(0)


_win_cache: Dict[Tuple, torch.Tensor] = {}
_small_cache: Dict[Tuple, torch.Tensor] = {}


def _device_window(win: torch.Tensor, device) -> torch.Tensor:
    """First row of a (C, 1, [1,] ws) window as a cached device float[ws]."""
    w = win.detach().reshape(win.shape[0], -1)[0].to(torch.float32).cpu()
    key = (tuple(w.tolist()), str(device))
    d = _win_cache.get(key)
    if d is None:
        d = w.to(device)
        _win_cache[key] = d
    return d


def _cached_small(kind: str, values: Tuple[float, ...], dtype, device) -> torch.Tensor:
    key = (kind, values, dtype, str(device))
    t = _small_cache.get(key)
    if t is None:
        t = torch.tensor(list(values), dtype=dtype).to(device)
        _small_cache[key] = t
    return t


class _MsSsimFunction(torch.autograd.Function):
    """levels == 1 -> single-scale SSIM (ssim.py:95-150); otherwise MS-SSIM (ssim.py:153-225)."""

    @staticmethod
    def forward(ctx, X, Y, win_dev, win_size, C1, C2, weights, size_average, use_relu):
        if not (X.is_cuda and Y.is_cuda and X.dtype == torch.float32):
            raise _lib.FcdError("fcdgan_b200 SSIM takes fp32 CUDA tensors (there is no CPU path)")
        X, Y = X.contiguous(), Y.contiguous()
        B, C, H, W = X.shape
        planes = B * C
        dev = X.device
        levels = len(weights)
        ws = win_size
        sums = torch.empty((levels, 2, planes), dtype=torch.float64, device=dev)
        Xs, Ys, counts = [X], [Y], []
        for l in range(levels):
            h, w = Xs[l].shape[2], Xs[l].shape[3]
            if h < ws or w < ws:
                warnings.warn(f"Skipping Gaussian Smoothing for input: {tuple(Xs[l].shape)} and win size: {ws}")
            oh, ow = (h - ws + 1 if h >= ws else h), (w - ws + 1 if w >= ws else w)
            counts.append(float(oh * ow))
            _call("fcd_ssim_level_fwd", Xs[l].data_ptr(), Ys[l].data_ptr(), planes, h, w, win_dev.data_ptr(), ws, C1, C2,
                  sums[l].data_ptr(), None, 0)
            if l < levels - 1:
                ph, pw = h % 2, w % 2
                nh, nw = (h + 2 * ph - 2) // 2 + 1, (w + 2 * pw - 2) // 2 + 1
                xn = torch.empty((B, C, nh, nw), dtype=torch.float32, device=dev)
                yn = torch.empty_like(xn)
                _call("fcd_avgpool2_fwd", Xs[l].data_ptr(), planes, h, w, ph, pw, xn.data_ptr())
                _call("fcd_avgpool2_fwd", Ys[l].data_ptr(), planes, h, w, ph, pw, yn.data_ptr())
                Xs.append(xn)
                Ys.append(yn)
        cnt = _cached_small("counts", tuple(counts), torch.float64, dev)
        wts = _cached_small("weights", tuple(float(v) for v in weights), torch.float32, dev)
        prod = torch.empty(planes, dtype=torch.float32, device=dev)
        out = torch.empty(1 if size_average else B, dtype=torch.float32, device=dev)
        _call("fcd_msssim_combine_fwd", sums.data_ptr(), cnt.data_ptr(), wts.data_ptr(), levels, planes, C,
              1 if size_average else 0, 1 if use_relu else 0, prod.data_ptr(), out.data_ptr())
        ctx.state = (Xs, Ys, sums, cnt, wts, prod, win_dev, ws, C1, C2, levels, size_average, use_relu)
        return out.reshape(()) if size_average else out

    @staticmethod
    def backward(ctx, gout):
        Xs, Ys, sums, cnt, wts, prod, win_dev, ws, C1, C2, levels, size_average, use_relu = ctx.state
        B, C = Xs[0].shape[0], Xs[0].shape[1]
        planes = B * C
        dev = Xs[0].device
        gout = gout.contiguous().to(torch.float32)
        coef = torch.empty((levels, planes), dtype=torch.float32, device=dev)
        _call("fcd_msssim_combine_bwd", sums.data_ptr(), cnt.data_ptr(), wts.data_ptr(), levels, planes, C,
              1 if size_average else 0, 1 if use_relu else 0, prod.data_ptr(), gout.data_ptr(), coef.data_ptr())
        dXn = dYn = None
        for l in range(levels - 1, -1, -1):
            h, w = Xs[l].shape[2], Xs[l].shape[3]
            oh, ow = (h - ws + 1 if h >= ws else h), (w - ws + 1 if w >= ws else w)
            dmaps = torch.empty((5, planes, oh, ow), dtype=torch.float32, device=dev)
            _call("fcd_ssim_level_fwd", Xs[l].data_ptr(), Ys[l].data_ptr(), planes, h, w, win_dev.data_ptr(), ws, C1, C2, None,
                  dmaps.data_ptr(), 1 if l == levels - 1 else 0)
            dX = torch.empty_like(Xs[l])
            dY = torch.empty_like(Ys[l])
            _call("fcd_ssim_level_bwd", dmaps.data_ptr(), Xs[l].data_ptr(), Ys[l].data_ptr(), planes, h, w, win_dev.data_ptr(),
                  ws, coef[l].data_ptr(), dX.data_ptr(), dY.data_ptr(), 0)
            del dmaps
            if dXn is not None:
                ph, pw = h % 2, w % 2
                _call("fcd_avgpool2_bwd", dXn.data_ptr(), planes, h, w, ph, pw, dX.data_ptr(), 1)
                _call("fcd_avgpool2_bwd", dYn.data_ptr(), planes, h, w, ph, pw, dY.data_ptr(), 1)
            dXn, dYn = dX, dY
        return dXn, dYn, None, None, None, None, None, None, None


def _prepare(X, Y):
    if not X.shape == Y.shape:
        raise ValueError("Input images should have the same dimensions.")
    for d in range(len(X.shape) - 1, 1, -1):
        X = X.squeeze(dim=d)
        Y = Y.squeeze(dim=d)
    return X, Y


def ssim(X, Y, data_range=255, size_average=True, win_size=11, win_sigma=1.5, win=None, K=(0.01, 0.03),
         nonnegative_ssim=False):
    """Interface of ssim — ssim.py:95-150."""
    X, Y = _prepare(X, Y)
    if len(X.shape) not in (4, 5):
        raise ValueError(f"Input images should be 4-d or 5-d tensors, but got {X.shape}")
    if len(X.shape) == 5:
        raise ValueError("fcdgan_b200.ssim: 5-d (conv3d) inputs are not on the FCD-GAN path and are not supported")
    if not X.type() == Y.type():
        raise ValueError("Input images should have the same dtype.")
    if win is not None:
        win_size = win.shape[-1]
    if not (win_size % 2 == 1):
        raise ValueError("Window size should be odd.")
    if win is None:
        win = _fspecial_gauss_1d(win_size, win_sigma)
    C1, C2 = (K[0] * data_range) ** 2, (K[1] * data_range) ** 2
    return _MsSsimFunction.apply(X, Y, _device_window(win, X.device), int(win_size), float(C1), float(C2), (1.0,),
                                 bool(size_average), bool(nonnegative_ssim))


def ms_ssim(X, Y, data_range=255, size_average=True, win_size=11, win_sigma=1.5, win=None, weights=None,
            K=(0.01, 0.03)):
    """Interface of ms-ssim — ssim.py:153-225."""
    X, Y = _prepare(X, Y)
    if not X.type() == Y.type():
        raise ValueError("Input images should have the same dtype.")
    if len(X.shape) == 5:
        raise ValueError("fcdgan_b200.ms_ssim: 5-d (conv3d) inputs are not on the FCD-GAN path and are not supported")
    if len(X.shape) != 4:
        raise ValueError(f"Input images should be 4-d or 5-d tensors, but got {X.shape}")
    if win is not None:
        win_size = win.shape[-1]
    if not (win_size % 2 == 1):
        raise ValueError("Window size should be odd.")
    smaller_side = min(X.shape[-2:])
    assert smaller_side > (win_size - 1) * (2 ** 4), \
        "Image size should be larger than %d due to the 4 downsamplings in ms-ssim" % ((win_size - 1) * (2 ** 4))
    if weights is None:
        weights = _MS_WEIGHTS
    if isinstance(weights, torch.Tensor):
        weights = weights.detach().cpu().tolist()
    # the reference stores the weights in fp32 (torch.FloatTensor, ssim.py:201)
    weights = tuple(float(torch.tensor(float(v), dtype=torch.float32)) for v in weights)
    if win is None:
        win = _fspecial_gauss_1d(win_size, win_sigma)
    C1, C2 = (K[0] * data_range) ** 2, (K[1] * data_range) ** 2
    return _MsSsimFunction.apply(X, Y, _device_window(win, X.device), int(win_size), float(C1), float(C2), weights,
                                 bool(size_average), True)


class SSIM(torch.nn.Module):
    """ssim.py:228-268."""

    def __init__(self, data_range=255, size_average=True, win_size=11, win_sigma=1.5, channel=3, spatial_dims=2,
                 K=(0.01, 0.03), nonnegative_ssim=False):
        super().__init__()
        self.win_size = win_size
        self.win = _fspecial_gauss_1d(win_size, win_sigma).repeat([channel, 1] + [1] * spatial_dims)
        self.size_average = size_average
        self.data_range = data_range
        self.K = K
        self.nonnegative_ssim = nonnegative_ssim

    def forward(self, X, Y):
        return ssim(X, Y, data_range=self.data_range, size_average=self.size_average, win=self.win, K=self.K,
                    nonnegative_ssim=self.nonnegative_ssim)


class MS_SSIM(torch.nn.Module):
    """ssim.py:271-311."""

    def __init__(self, data_range=255, size_average=True, win_size=11, win_sigma=1.5, channel=3, spatial_dims=2,
                 weights=None, K=(0.01, 0.03)):
        super().__init__()
        self.win_size = win_size
        self.win = _fspecial_gauss_1d(win_size, win_sigma).repeat([channel, 1] + [1] * spatial_dims)
        self.size_average = size_average
        self.data_range = data_range
        self.weights = weights
        self.K = K

    def forward(self, X, Y):
        return ms_ssim(X, Y, data_range=self.data_range, size_average=self.size_average, win=self.win,
                       weights=self.weights, K=self.K)
