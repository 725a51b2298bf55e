"""Internal NHWC execution engine behind the reference-named nn.Modules (modules.py).

Inside a network everything is NHWC and channel padded; NCHW fp32 exists only at the module boundary
(SURVEY.md §8(b)).  A network forward is a sequence of C-ABI kernel launches (include/fcd_b200.h) recorded
on a `Tape`; the tape's closures, run in reverse, are the hand-written backward pass.  torch supplies
device memory (caching allocator), the current stream and the autograd hook (`NetFunction`) only.

Tensors
  * `Act`  split activation: two bf16 NHWC planes (hi, lo), value = hi + lo (lo is None in "fast" precision).
           Convolution operands.  May be a channel slice of a concatenation buffer (pitch `ld` > Cp) or the images of one
           siamese branch inside a batch that carries both (`batch_view`).
  * `Z`    raw convolution output, fp32 NHWC, plus the per-channel sum / sum-of-squares for BatchNorm.
  * gradients w.r.t. an `Act` are fp32 NHWC (`Act.grad`), gradients w.r.t. a `Z` are split (`Z.dz`), because
    they are the operands of the dgrad / wgrad convolutions.

There is no CPU or torch fallback here: every arithmetic step is a libfcd_b200.so call and raises if the
library is missing (`_lib.FcdError`).
"""
from __future__ import annotations

from typing import Callable, Dict, List, Optional, Tuple

import torch

from . import _lib

ACT_NONE, ACT_RELU, ACT_PRELU, ACT_LEAKY = 0, 1, 2, 3
ENGINE_AUTO, ENGINE_SIMT, ENGINE_TC = 0, 1, 2
BN_EPS = 1e-5        # nn.BatchNorm2d defaults relied on by Module.py:27,30,156,...
BN_MOMENTUM = 0.1

_cfg = {"split": True, "engine": ENGINE_AUTO, "fuse_stats": True, "im2col": True, "streams": 1, "sync_bn": None,
        "batch_branches": True, "rowpack": True, "zero_pool": True}
# A/B switches without code changes: FCD_ENGINE="rowpack=0,batch_branches=0" (bench.py reports the dictionary in its line)
import os as _os

for _item in filter(None, _os.environ.get("FCD_ENGINE", "").split(",")):
    _k, _, _v = _item.partition("=")
    if _k.strip() not in ("rowpack", "batch_branches", "im2col", "fuse_stats", "zero_pool"):
        raise ValueError(f"FCD_ENGINE: unknown switch {_k!r}")
    _cfg[_k.strip()] = bool(int(_v or 1))

DEBUG_CAPTURE = None  # set to a list to record (kind, tensor) pairs from the backward pass (scripts/dbg_g2.py)
launch_count = 0     # number of libfcd_b200 kernels-launching calls (bench.py reports it)


def set_precision(mode: str) -> None:
    """'parity': split-bf16 operands, three tcgen05 MMAs per product (fp32-class accuracy, the mode the 1e-3
    change-density parity bar is checked in).  'fast': single bf16 plane, one MMA (throughput mode)."""
    if mode not in ("parity", "fast"):
        raise ValueError("precision must be 'parity' or 'fast'")
    _cfg["split"] = mode == "parity"


def set_streams(n: int) -> None:
    """Number of CUDA streams a network pass may use (1 = everything on the current stream).  With 2, independent work —
    the two siamese branches of the Segmentor encoder / the Discriminator, and every weight gradient next to the
    data-gradient chain of the backward pass — is forked onto a side stream and joined with events; under CUDA-graph
    capture the forks become parallel branches of the graph."""
    if n not in (1, 2):
        raise ValueError("streams must be 1 or 2")
    _cfg["streams"] = n


def get_streams() -> int:
    return _cfg["streams"]


def set_sync_bn(enabled: bool, group=None) -> None:
    """Synchronised BatchNorm for data-parallel runs (no reference counterpart — the reference is single-device; SURVEY.md
    §8(e)): every train-mode BatchNorm call all-reduces its batch sums (2*C doubles forward, 2*C doubles backward) over
    `group`, so an N-rank run with equal shards computes exactly the statistics, running statistics and gradients of the
    single-process run on the concatenated batch (tests/test_dp_gpu.py, 2 ranks on NCCL).  Off by default (per-rank
    statistics, the DDP default).  Eager launches only: under CUDA-graph capture it raises."""
    import torch.distributed as dist

    if enabled and not dist.is_initialized():
        raise RuntimeError("set_sync_bn(True) needs an initialised torch.distributed process group")
    _cfg["sync_bn"] = (group,) if enabled else None


def _sync_world() -> int:
    import torch.distributed as dist

    return dist.get_world_size(_cfg["sync_bn"][0]) if _cfg["sync_bn"] is not None else 1


def _all_reduce_sum(t: torch.Tensor) -> None:
    import torch.distributed as dist

    if _capturing():
        # measured (round 2): SyncBN's all-reduces captured inside graph segments, mixed with the eager gradient all-reduces
        # between the segments, hung on torch 2.11 / NCCL 2.28.9 — refuse instead of hanging a multi-GPU job
        raise _lib.FcdError("set_sync_bn(True) is not supported under CUDA-graph capture: run SyncBN steps eagerly")
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=_cfg["sync_bn"][0])


def get_precision() -> str:
    return "parity" if _cfg["split"] else "fast"


def im2col_enabled() -> bool:
    """Receptive-field-packed first discriminator layer for data inputs (conv_im2col_s2); tcgen05 engine only."""
    return _cfg["im2col"] and _cfg["engine"] != ENGINE_SIMT


def set_im2col(on: bool) -> None:
    _cfg["im2col"] = bool(on)


def batch_branches_enabled() -> bool:
    """The Discriminator's two siamese branches run as one batch through the convolutions (modules.Discriminator_SRGAN_simple).
    Off under SyncBN (its grouped BatchNorm form keeps per-rank statistics only)."""
    return _cfg["batch_branches"] and _cfg["sync_bn"] is None


def set_batch_branches(on: bool) -> None:
    _cfg["batch_branches"] = bool(on)


def set_rowpack(on: bool) -> None:
    """Tight row-packed operand layout for the few-band 9x9 layers (row_pack_pixels); off = the 4-pixel form everywhere."""
    _cfg["rowpack"] = bool(on)


def set_engine(engine: int) -> None:
    _cfg["engine"] = engine


PROFILE = None       # set to a list: every C-ABI call is bracketed by CUDA events -> (name, tag, flops, bytes, e0, e1)


def _raw_stream() -> int:
    """cudaStream_t of torch's current stream on the current device (the capture stream under CUDA-graph capture).  The
    networks make their tensors' device current for the duration of a pass (NetFunction), so this is the tensors' device."""
    return torch._C._cuda_getCurrentRawStream(torch.cuda.current_device())


def _call(name, *args, tag=None, flops=0.0, nbytes=0.0):
    global launch_count
    launch_count += 1
    if PROFILE is None:
        return _lib.call(name, *args, _raw_stream())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = _lib.call(name, *args, _raw_stream())
    e1.record()
    PROFILE.append((name, tag or name, flops, nbytes, e0, e1))
    return rc


def _conv_engine_name(Cin_p, Cout_p, KH, KW, stride) -> str:
    if _cfg["engine"] == ENGINE_SIMT:
        return "simt"
    return "tc" if _lib.load().fcd_conv2d_tc_supported(Cin_p, Cout_p, KH, KW, stride) else "simt"


def pad_ch(c: int, tc: bool = True) -> int:
    """Channel padding rule.  tc=True: round up to a multiple of 64 so the tensor lands on the tcgen05 engine (the
    13-band head / tail of the Generator and the Segmentor's first layer are zero-padded to 64 channels: the padded
    MMA work is cheaper than leaving 9x9 convolutions on CUDA cores).  tc=False: multiples of 16 (SIMT engine —
    the stride-2 discriminator layers)."""
    return (c + 63) // 64 * 64 if (tc or c >= 64) else (c + 15) // 16 * 16


# --------------------------------------------------------------------------------------------------
class Act:
    """Split NHWC activation view (possibly a channel slice of a wider buffer)."""

    __slots__ = ("hi", "lo", "N", "H", "W", "C", "Cp", "ld", "parent", "off", "n0", "_grad", "_ready", "name")

    def __init__(self, hi, lo, N, H, W, C, Cp, ld, parent=None, off=0, name="", n0=None):
        self.hi, self.lo = hi, lo
        self.N, self.H, self.W, self.C, self.Cp, self.ld = N, H, W, C, Cp, ld
        self.parent, self.off = parent, off
        self.n0 = n0             # not None: a view of images [n0, n0 + N) of `parent` (batch_view)
        self._grad = None
        self._ready = False
        self.name = name

    @staticmethod
    def empty(N, H, W, C, device, Cp=None, name="") -> "Act":
        Cp = Cp or pad_ch(C)
        hi = torch.empty((N, H, W, Cp), dtype=torch.bfloat16, device=device)
        lo = torch.empty_like(hi) if _cfg["split"] else None
        return Act(hi, lo, N, H, W, C, Cp, Cp, name=name)

    def slice(self, off: int, C: int) -> "Act":
        assert off % 8 == 0 and C % 8 == 0 and off + C <= self.Cp
        return Act(self.hi[..., off:off + C], None if self.lo is None else self.lo[..., off:off + C], self.N,
                   self.H, self.W, C, C, self.ld, parent=self, off=off, name=f"{self.name}[{off}:{off + C}]")

    def batch_view(self, n0: int, n: int) -> "Act":
        """Images [n0, n0 + n) of this activation (one siamese branch of a batch that carries both)."""
        assert 0 <= n0 and n0 + n <= self.N
        return Act(self.hi[n0:n0 + n], None if self.lo is None else self.lo[n0:n0 + n], n, self.H, self.W, self.C, self.Cp,
                   self.ld, parent=self, off=0, name=f"{self.name}[n{n0}:{n0 + n}]", n0=n0)

    @property
    def npix(self) -> int:
        return self.N * self.H * self.W

    def p_hi(self):
        return self.hi.data_ptr()

    def p_lo(self):
        return None if self.lo is None else self.lo.data_ptr()

    # ---- gradient buffer (fp32 NHWC, same pitch as the data) ----
    @property
    def grad(self) -> torch.Tensor:
        if self._grad is None:
            if self.parent is not None:
                g = self.parent.grad
                if self.n0 is not None:
                    g = g[self.n0:self.n0 + self.N]
                self._grad = g[..., self.off:self.off + self.Cp]
            else:
                self._grad = torch.empty((self.N, self.H, self.W, self.Cp), dtype=torch.float32, device=self.hi.device)
        return self._grad

    @property
    def ready(self) -> bool:
        return self._ready or (self.parent is not None and self.parent.ready)

    def mark_ready(self):
        self._ready = True

    def reset_grad(self):
        self._grad = None
        self._ready = False


class Z:
    """fp32 NHWC convolution output + BatchNorm statistics."""

    __slots__ = ("t", "N", "H", "W", "C", "Cp", "ld", "sum", "sqsum", "stats", "sum_local", "dz", "pack_m", "pack_p", "bias_param", "db_done")

    def __init__(self, t, N, H, W, C, Cp):
        self.t, self.N, self.H, self.W, self.C, self.Cp, self.ld = t, N, H, W, C, Cp, Cp
        self.sum = self.sqsum = None
        self.stats = None        # the (2, Cp) double tensor behind sum / sqsum (SyncBN all-reduces it in one call)
        self.sum_local = None    # SyncBN: this rank's own sum, kept for the closed-form conv-bias gradient
        self.bias_param = None   # the producing convolution's bias: bn_act's backward may deliver its gradient (db_done)
        self.db_done = False
        self.dz: Optional[Act] = None
        self.pack_m = 0          # > 0: the producer wants its output gradient as a PackedAct with this left margin
        self.pack_p = 4          # ... and this many pixels per pack (4: 16-slot form, else the tight row-packed form)

    @property
    def npix(self):
        return self.N * self.H * self.W


PACK_M = 4   # left margin (pixels) of a 4-pixel channel-packed tensor; must be >= the convolution padding


def row_pack_pixels(C: int, KW: int) -> int:
    """Pixels per pack for a KW-wide filter over a C-band image: KW (one tap of K = pad64(KW*C) per filter row, bands packed
    tightly) when that needs fewer 64-channel K chunks per filter row than the 4-pixel form's ceil(KW/4) taps, else 4.
    13 bands, 9x9 (Module.py:146,158): 117 -> 2 chunks instead of 3; 3 or 4 bands, 9x9: 1 instead of 3; 3x3 filters stay 4."""
    if not _cfg["rowpack"]:
        return 4
    return KW if (KW * C + 63) // 64 < (KW + 3) // 4 and KW * C <= 256 else 4


class PackedAct:
    """Split NHWC tensor (N, H, W + M, Kp) holding a few-band image with P horizontally adjacent pixels packed into the
    channel axis.  P = 4 (Kp = 64, <= 16 bands): v[n, h, w'', j*16 + c] = image[n, c, h, w'' - M + j]
    (fcd_stage_nchw_to_split_pack4) — a K x K convolution over it needs ceil(K/4) taps of 64 channels per filter row instead
    of K taps.  P != 4: bands packed tightly, v[n, h, w'', j*C + c] = image[n, c, h, w'' - M + j], j < P, Kp = pad64(P*C)
    (fcd_stage_nchw_to_split_rowpack) — with P = K a whole filter row is one tap."""

    __slots__ = ("hi", "lo", "N", "H", "W", "Wp", "C", "M", "P", "Kp")

    def __init__(self, x_nchw: torch.Tensor, M: int = PACK_M, P: int = 4):
        N, C, H, W = x_nchw.shape
        assert C <= 16
        self.N, self.C, self.H, self.W, self.M, self.Wp, self.P = N, C, H, W, M, W + M, P
        self.Kp = 64 if P == 4 else pad_ch(P * C)
        self.hi = torch.empty((N, H, self.Wp, self.Kp), dtype=torch.bfloat16, device=x_nchw.device)
        self.lo = torch.empty_like(self.hi) if _cfg["split"] else None
        x_nchw = x_nchw.contiguous()
        if P == 4:
            _call("fcd_stage_nchw_to_split_pack4", x_nchw.data_ptr(), N, C, H, W, M, self.hi.data_ptr(), _lib.ptr(self.lo))
        else:
            _call("fcd_stage_nchw_to_split_rowpack", x_nchw.data_ptr(), N, C, H, W, M, P, self.Kp, self.hi.data_ptr(),
                  _lib.ptr(self.lo))

    def p_hi(self):
        return self.hi.data_ptr()

    def p_lo(self):
        return None if self.lo is None else self.lo.data_ptr()


# --------------------------------------------------------------------------------------------------
_workspace: Dict[tuple, torch.Tensor] = {}      # (device, stream) -> split-K scratch of the weight-gradient kernels
_side_streams: Dict[torch.device, "torch.cuda.Stream"] = {}
_weight_epoch = 0     # bumped whenever parameters may have changed WITHOUT their version counter moving (graph replay)


def clear_caches():
    _workspace.clear()


def _side(device) -> "torch.cuda.Stream":
    """The engine's side stream on `device` (set_streams(2)): weight gradients run there, next to the data-gradient chain."""
    device = torch.device(device)
    st = _side_streams.get(device)
    if st is None:
        st = _side_streams[device] = torch.cuda.Stream(device=device)
    return st


def bump_weight_epoch() -> None:
    """Invalidate every cached packed weight.  A CUDA-graph replay updates parameters on the device without
    touching `Tensor._version`, so `graph.GraphedStep` calls this after each replay."""
    global _weight_epoch
    _weight_epoch += 1


def invalidate_weight_cache() -> None:
    """Public form of bump_weight_epoch().  The packed copies are keyed by (storage pointer, `Tensor._version`); writes
    that go through `.data` (`p.data.clamp_(-1, 1)` — the WGAN clip commented out at Demo_RSSS.py:308-309 —, `p.data.copy_`,
    EMA updates, `dist.broadcast(p.data)`) do NOT move the version counter, so call this after any such write.  Optimizer
    steps (fused ones included, through torch's global post-step hook), `load_state_dict`, `.to()` and in-place ops on the
    parameter itself are detected without it."""
    bump_weight_epoch()


def _after_any_optimizer_step(*_args, **_kwargs) -> None:
    bump_weight_epoch()


# torch's fused optimizers (`torch.optim.Adam(fused=True)`, ...) update the parameters WITHOUT moving `Tensor._version`
# (checked on torch 2.11: the version counter is the same before and after `step()`), so the (pointer, version) stamp of the
# packed copies cannot see them: every optimizer step in the process invalidates the cache through torch's global post-step hook.
try:
    from torch.optim.optimizer import register_optimizer_step_post_hook as _reg_post_hook

    _reg_post_hook(_after_any_optimizer_step)
except ImportError:          # older torch: the stamp check alone (un-fused optimizers move the version counter)
    pass


def _capturing() -> bool:
    return torch.cuda.is_current_stream_capturing()


def _packed(w: torch.Tensor, Cout_p: int, Cin_p: int, mode: int, tag: str = ""):
    """Packed split-bf16 copy of an OIHW weight.  The cache lives ON the parameter object and is valid while
    (storage pointer, in-place version counter) are unchanged — an optimizer step, load_state_dict or .to()
    invalidates it."""
    cache = getattr(w, "_fcd_pack", None)
    if cache is None:
        cache = {}
        try:
            w._fcd_pack = cache
        except AttributeError:
            pass
    key = (mode, _cfg["split"], Cout_p, Cin_p, tag)
    stamp = (w.data_ptr(), w._version, _weight_epoch)
    capturing = _capturing()      # under graph capture always re-pack (the packing kernel must be part of the graph)
    ent = None if capturing else cache.get(key)
    if ent is not None and ent[0] == stamp:
        if ent[3] is not None and ent[3][0] != torch.cuda.current_stream().cuda_stream:
            torch.cuda.current_stream().wait_event(ent[3][1])     # packed on another stream: order this stream after it
        return ent[1], ent[2]
    Cout, Cin, KH, KW = w.shape
    rows, cols = (Cout_p, Cin_p) if mode == 0 else (Cin_p, Cout_p)
    hi = torch.empty((KH * KW, rows, cols), dtype=torch.bfloat16, device=w.device)
    lo = torch.empty_like(hi) if _cfg["split"] else None
    _call("fcd_pack_conv_weight", w.data_ptr(), Cout, Cin, KH, KW, Cout_p, Cin_p, mode, hi.data_ptr(), _lib.ptr(lo))
    if not capturing:
        ev = None
        if _cfg["streams"] > 1:
            e = torch.cuda.Event()
            e.record()
            ev = (torch.cuda.current_stream().cuda_stream, e)
        cache[key] = (stamp, hi, lo, ev)
    return hi, lo


def _derived(w: torch.Tensor, key: str, fn) -> torch.Tensor:
    """A tensor derived from parameter `w` by `fn` (small torch reshuffles of the weights), cached on the parameter
    like the packed copies."""
    cache = getattr(w, "_fcd_derived", None)
    if cache is None:
        cache = {}
        try:
            w._fcd_derived = cache
        except AttributeError:
            pass
    stamp = (w.data_ptr(), w._version, _weight_epoch)
    capturing = _capturing()
    ent = None if capturing else cache.get(key)
    if ent is not None and ent[0] == stamp:
        return ent[1]
    t = fn(w.detach())
    if not capturing:
        cache[key] = (stamp, t)
    return t


def _w_pack4_in(w: torch.Tensor) -> torch.Tensor:
    """[Cout][C<=16][KH][KW] -> fake OIHW [Cout][64][KH][ceil(KW/4)] with W'[co][j*16+c][r][g] = w[co][c][r][4g+j]."""
    Cout, C, KH, KW = w.shape
    ng = (KW + 3) // 4
    wp = w.new_zeros((Cout, 16, KH, ng * 4))
    wp[:, :C, :, :KW] = w
    return wp.view(Cout, 16, KH, ng, 4).permute(0, 4, 1, 2, 3).reshape(Cout, 64, KH, ng).contiguous()


def _w_unpack4_in(dwp: torch.Tensor, C: int, KW: int) -> torch.Tensor:
    """inverse of _w_pack4_in for a gradient: [Cout][64][KH][ng] -> [Cout][C][KH][KW]."""
    Cout, _, KH, ng = dwp.shape
    return dwp.view(Cout, 4, 16, KH, ng).permute(0, 2, 3, 4, 1).reshape(Cout, 16, KH, ng * 4)[:, :C, :, :KW]


def _w_rowpack_in(w: torch.Tensor) -> torch.Tensor:
    """[Cout][C][KH][KW] -> fake OIHW [Cout][pad64(KW*C)][KH][1] with W'[co][j*C+c][r][0] = w[co][c][r][j]."""
    Cout, C, KH, KW = w.shape
    wp = w.new_zeros((Cout, pad_ch(KW * C), KH, 1))
    wp[:, :KW * C, :, 0] = w.permute(0, 3, 1, 2).reshape(Cout, KW * C, KH)
    return wp


def _w_unrowpack_in(dwp: torch.Tensor, C: int, KW: int) -> torch.Tensor:
    """inverse of _w_rowpack_in for a gradient: [Cout][Kp][KH][1] -> [Cout][C][KH][KW]."""
    Cout, _, KH, _ = dwp.shape
    return dwp[:, :KW * C, :, 0].reshape(Cout, KW, C, KH).permute(0, 2, 3, 1)


def _w_unrowpack_out(dwp: torch.Tensor, Cout: int, KW: int) -> torch.Tensor:
    """Weight gradient taken against a row-packed OUTPUT gradient: T[j*Cout+co][ci][r][0] = dw[co][ci][r][KW-1-j]
    -> [Cout][Cin][KH][KW]."""
    _, Cin, KH, _ = dwp.shape
    return dwp[:KW * Cout, :, :, 0].reshape(KW, Cout, Cin, KH).flip(0).permute(1, 2, 3, 0)


def _w_pack4_out(w: torch.Tensor) -> torch.Tensor:
    """[Cout<=16][Cin][KH][KW] -> fake OIHW [64][Cin][KH][KW+3] with Wq[j*16+co][ci][r][s''] = w[co][ci][r][s''-j]:
    column j*16+co produces output pixel 4q+j, channel co."""
    Cout, Cin, KH, KW = w.shape
    wq = w.new_zeros((4, 16, Cin, KH, KW + 3))
    for j in range(4):
        wq[j, :Cout, :, :, j:j + KW] = w
    return wq.view(64, Cin, KH, KW + 3)


def _ws(device, nbytes: int) -> torch.Tensor:
    if _capturing():              # graph-private memory must not leak into the eager cache
        return torch.empty(max(nbytes, 16), dtype=torch.uint8, device=device)
    key = (torch.device(device), torch.cuda.current_stream(device).cuda_stream)     # one scratch per stream: launches on
    cur = _workspace.get(key)                                                       # different streams may run concurrently
    if cur is None or cur.numel() < nbytes:
        cur = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8, device=device)
        _workspace[key] = cur
    return cur


def _padded_vec(v: torch.Tensor, n: int) -> torch.Tensor:
    if v.numel() == n:
        return v
    out = torch.zeros(n, dtype=v.dtype, device=v.device)
    out[:v.numel()] = v
    return out


class Tape:
    """Records backward closures during a network forward; `run()` replays them in reverse.  The same tape can
    be replayed several times (Demo_USSS.py:327 uses backward(retain_graph=True) followed by a second
    backward), so closures never free or overwrite forward state."""

    def __init__(self, device, record: bool):
        self.device = device
        self.record = record
        self.ops: List[Callable[[], None]] = []
        self.acts: List[Act] = []
        self.pgrads: Dict[int, torch.Tensor] = {}   # id(param) -> grad (this replay)
        self.params: Dict[int, torch.Tensor] = {}
        self.param_grads = True                     # False: no parameter of the network requires a gradient (a pass that only
                                                    # carries a gradient THROUGH the network): weight-gradient launches are skipped
        self.forked = False                         # work of this replay is in flight on the side stream
        self.held: List = []                        # buffers that work reads: kept alive until it has been joined
        self._arena = None                          # zero-initialised float64 pool the accumulators are carved from (zeros64)
        self._arena_used = 0

    def push(self, fn):
        if self.record:
            self.ops.append(fn)

    ARENA = 1 << 16     # doubles per pool block (512 KB: every accumulator of a Segmentor pass; one fill instead of ~50)

    def zeros64(self, *shape) -> torch.Tensor:
        """Zero-initialised float64 tensor for a kernel's atomics / sums (BatchNorm statistics, backward reductions), carved
        from a pool that is filled once per block: a pass needed one tiny fill kernel per convolution and per BatchNorm
        backward.  Slices are never handed out twice (a replayed tape gets fresh zeros), 16-byte aligned."""
        if not _cfg["zero_pool"]:
            return torch.zeros(shape, dtype=torch.float64, device=self.device)
        n = 1
        for d in shape:
            n *= d
        n_al = (n + 1) // 2 * 2
        if self._arena is None or self._arena_used + n_al > self._arena.numel():
            self._arena = torch.zeros(max(self.ARENA, n_al), dtype=torch.float64, device=self.device)
            self._arena_used = 0
        v = self._arena[self._arena_used:self._arena_used + n].view(shape)
        self._arena_used += n_al
        return v

    def track(self, a: Act) -> Act:
        if self.record:
            self.acts.append(a)
        return a

    def new_act(self, N, H, W, C, Cp=None, name="") -> Act:
        return self.track(Act.empty(N, H, W, C, self.device, Cp, name))

    def pgrad(self, p: torch.Tensor) -> Tuple[torch.Tensor, int]:
        """(gradient buffer for parameter p, accumulate flag) — a parameter used by several layers / siamese
        branches accumulates after its first use in this replay."""
        k = id(p)
        if k in self.pgrads:
            return self.pgrads[k], 1
        g = torch.empty_like(p, dtype=torch.float32)
        self.pgrads[k] = g
        self.params[k] = p
        return g, 0

    def side_join(self):
        """Make the current stream wait for the side stream's work of this replay and release the buffers it was reading."""
        if self.forked:
            torch.cuda.current_stream().wait_stream(_side(self.device))
            self.forked = False
        self.held.clear()

    def on_side(self, fn, *hold):
        """Run `fn()` on the side stream, ordered after everything issued so far on the current stream.  One piece of side
        work is in flight at a time (the previous one is joined first), `hold` stays referenced until the next join."""
        self.side_join()
        side = _side(self.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fn()
        self.forked = True
        self.held.extend(hold)

    def run(self):
        for a in self.acts:
            a.reset_grad()
        self.pgrads = {}
        try:
            for fn in reversed(self.ops):
                fn(self)     # the tape is passed in (closures must not capture it: tape <-> closure cycles would
                             # keep a whole iteration's device buffers alive until Python's cyclic GC runs)
        finally:
            self.side_join()
        for a in self.acts:      # gradient buffers are per replay
            a.reset_grad()
        return self.pgrads


# --------------------------------------------------------------------------------------------------
# layer primitives (forward + recorded backward)
# --------------------------------------------------------------------------------------------------
def stage_input(tape: Tape, x: torch.Tensor, need_grad: bool, mask: Optional[torch.Tensor] = None, tc: bool = True) -> Act:
    """NCHW fp32 boundary tensor -> split NHWC (K15)."""
    N, C, H, W = x.shape
    x = x.contiguous()
    a = tape.new_act(N, H, W, C, Cp=pad_ch(C, tc), name="input")
    _call("fcd_stage_nchw_to_split", x.data_ptr(), _lib.ptr(mask), N, C, H, W, a.p_hi(), a.p_lo(), a.ld, a.Cp)
    return a


def stage_input_into(tape: Tape, x: torch.Tensor, dst: Act) -> None:
    """Stage an NCHW fp32 tensor into an existing (slice of an) activation buffer."""
    N, C, H, W = x.shape
    assert (N, H, W) == (dst.N, dst.H, dst.W) and C <= dst.Cp
    x = x.contiguous()
    _call("fcd_stage_nchw_to_split", x.data_ptr(), None, N, C, H, W, dst.p_hi(), dst.p_lo(), dst.ld, dst.Cp)


def act_to_nchw(tape: Tape, a: Act, grad_slot: dict) -> torch.Tensor:
    """Split activation -> NCHW fp32 boundary tensor (output of a stand-alone building block)."""
    out = torch.empty((a.N, a.C, a.H, a.W), dtype=torch.float32, device=tape.device)
    _call("fcd_unstage_split_to_nchw", a.p_hi(), a.p_lo(), a.ld, a.N, a.C, a.H, a.W, out.data_ptr())

    def backward(tape):
        dout = grad_slot["dout"].contiguous()
        assert not a.ready
        _call("fcd_stage_nchw_to_f32", dout.data_ptr(), a.N, a.C, a.H, a.W, a.grad.data_ptr(), a.ld, a.Cp)
        a.mark_ready()

    tape.push(backward)
    return out


def unstage_grad(a: Act) -> torch.Tensor:
    """fp32 NHWC gradient of a staged input -> NCHW."""
    out = torch.empty((a.N, a.C, a.H, a.W), dtype=torch.float32, device=a.hi.device)
    _call("fcd_unstage_f32_to_nchw", a.grad.data_ptr(), a.ld, a.N, a.C, a.H, a.W, out.data_ptr(), 0)
    return out


def conv(tape: Tape, x: Act, w: torch.Tensor, b: Optional[torch.Tensor], stride: int, pad: int, stats: bool,
         x_needs_grad: bool = True, wtag: str = "", frozen: bool = False) -> Z:
    """nn.Conv2d forward (Module.py:26-216) + recorded wgrad/dgrad.  `frozen`: the weights are constants (the VGG16 of the
    perception loss, Loss.py:25-27) — no weight / bias gradient is computed."""
    Cout, Cin, KH, KW = w.shape
    assert Cin == x.C, f"conv: weight expects {Cin} channels, activation has {x.C}"
    Cout_p, Cin_p = pad_ch(Cout), x.Cp
    N, H, W = x.N, x.H, x.W
    OH = (H + 2 * pad - KH) // stride + 1
    OW = (W + 2 * pad - KW) // stride + 1
    w_hi, w_lo = _packed(w, Cout_p, Cin_p, 0, wtag)
    bias = None if b is None else _padded_vec(b, Cout_p)
    zt = torch.empty((N, OH, OW, Cout_p), dtype=torch.float32, device=tape.device)
    z = Z(zt, N, OH, OW, Cout, Cout_p)
    fuse = stats and _cfg["fuse_stats"]
    if stats:
        st = tape.zeros64(2, Cout_p)
        z.sum, z.sqsum, z.stats = st[0], st[1], st
    eng = _conv_engine_name(Cin_p, Cout_p, KH, KW, stride)
    flops = 2.0 * N * OH * OW * Cout * Cin * KH * KW          # algorithmic (un-padded) multiply-adds x 2
    shape = f"{KH}x{KW}s{stride} {Cin}->{Cout}"
    _call("fcd_conv2d_fwd", x.p_hi(), x.p_lo(), x.ld, w_hi.data_ptr(), _lib.ptr(w_lo), _lib.ptr(bias), None, 0,
          zt.data_ptr(), z.ld, N, H, W, Cin_p, Cout_p, KH, KW, stride, pad,
          z.sum.data_ptr() if fuse else None, z.sqsum.data_ptr() if fuse else None, _cfg["engine"],
          tag=f"conv_fwd_{eng} {shape}", flops=flops)
    if stats and not fuse:
        _call("fcd_bn_stats", zt.data_ptr(), z.ld, z.npix, Cout_p, z.sum.data_ptr(), z.sqsum.data_ptr())
    z.bias_param = b

    def backward(tape):
        dz = z.dz
        assert dz is not None, "conv backward: output gradient missing"

        def wgrad():
            nbytes = _lib.load().fcd_conv2d_wgrad_workspace(N, H, W, Cin_p, Cout_p, KH, KW, stride, pad, _cfg["engine"])
            ws = _ws(tape.device, nbytes)
            _call("fcd_conv2d_wgrad", x.p_hi(), x.p_lo(), x.ld, dz.p_hi(), dz.p_lo(), dz.ld, gw.data_ptr(), _lib.ptr(gb),
                  N, H, W, Cin, Cin_p, Cout, Cout_p, KH, KW, stride, pad, acc, ws.data_ptr(), nbytes, _cfg["engine"],
                  tag=f"conv_wgrad_{eng} {shape}", flops=flops)

        want_w = not frozen and tape.param_grads
        if want_w:
            gw, acc = tape.pgrad(w)
            gb = None
            if b is not None and not z.db_done:
                gb, accb = tape.pgrad(b)
                assert accb == acc
        z.db_done = False
        side = _cfg["streams"] > 1 and want_w
        if want_w and not side:
            wgrad()
        if x_needs_grad:
            g = x.grad
            addend = g.data_ptr() if x.ready else None
            if stride == 1:
                wd_hi, wd_lo = _packed(w, Cout_p, Cin_p, 1, wtag)
                _call("fcd_conv2d_fwd", dz.p_hi(), dz.p_lo(), dz.ld, wd_hi.data_ptr(), _lib.ptr(wd_lo), None, addend,
                      x.ld, g.data_ptr(), x.ld, N, OH, OW, Cout_p, Cin_p, KH, KW, 1, KH - 1 - pad, None, None,
                      _cfg["engine"], tag=f"conv_dgrad_{_conv_engine_name(Cout_p, Cin_p, KH, KW, 1)} {shape}", flops=flops)
            else:
                wd_hi, wd_lo = _packed(w, Cout_p, Cin_p, 1, wtag)
                deng = _conv_engine_name(Cout_p, Cin_p, KH, KW, stride)
                _call("fcd_conv2d_dgrad_strided", dz.p_hi(), dz.p_lo(), dz.ld, w_hi.data_ptr(), _lib.ptr(w_lo),
                      wd_hi.data_ptr(), _lib.ptr(wd_lo), addend, x.ld, g.data_ptr(), x.ld, N, H, W, Cin_p, Cout_p, KH, KW,
                      stride, pad, _cfg["engine"], tag=f"conv_dgrad_{deng} {shape}", flops=flops)
            x.mark_ready()
        if side:
            # the weight gradient is off the critical path (nothing in this replay reads it): it runs on the side stream
            # AFTER the data gradient (both want every SM) and so overlaps the HBM-bound BatchNorm backward of the next layer
            tape.on_side(wgrad, dz)
        z.dz = None

    tape.push(backward)
    return z


def conv_small_in(tape: Tape, xp: PackedAct, w: torch.Tensor, b: Optional[torch.Tensor], pad: int, stats: bool) -> Z:
    """Stride-1 convolution whose INPUT has <= 16 channels (Generator head Module.py:146, Segmentor first layer
    Module.py:26), on a channel-packed input: ceil(KW/4) taps of K = 64 per filter row in the 4-pixel form, ONE tap of
    K = pad64(KW*C) in the row-packed form (xp.P == KW).  No input gradient (the input is data)."""
    Cout, C, KH, KW = w.shape
    assert C == xp.C
    M, N, Kp = xp.M, xp.N, xp.Kp
    row = xp.P != 4
    assert pad <= M and (not row or xp.P == KW)
    ng, sstep = (1, 1) if row else ((KW + 3) // 4, 4)
    Cout_p = pad_ch(Cout)
    OH, OW = xp.H + 2 * pad - KH + 1, xp.W + 2 * pad - KW + 1
    form = "rowpack" if row else "pack4"
    wf = _derived(w, form + "_in", _w_rowpack_in if row else _w_pack4_in)
    w_hi, w_lo = _packed(wf, Cout_p, Kp, 0, form + "_in")
    bias = None if b is None else _padded_vec(b, Cout_p)
    zt = torch.empty((N, OH, OW, Cout_p), dtype=torch.float32, device=tape.device)
    z = Z(zt, N, OH, OW, Cout, Cout_p)
    flops = 2.0 * N * OH * OW * Cout * C * KH * KW
    shape = f"{KH}x{KW}s1 {C}->{Cout}"
    _call("fcd_conv2d_taps_fwd", xp.p_hi(), xp.p_lo(), Kp, xp.H, xp.Wp, w_hi.data_ptr(), _lib.ptr(w_lo), _lib.ptr(bias), None, 0,
          zt.data_ptr(), z.ld, N, OH, OW, Kp, Cout_p, KH, ng, -pad, 1, -pad + M, sstep, 1, 1,
          tag=f"conv_fwd_tc_{form} {shape}", flops=flops)
    if stats:
        st = tape.zeros64(2, Cout_p)
        z.sum, z.sqsum, z.stats = st[0], st[1], st
        _call("fcd_bn_stats", zt.data_ptr(), z.ld, z.npix, Cout_p, z.sum.data_ptr(), z.sqsum.data_ptr())
    z.bias_param = b

    def backward(tape):
        dz = z.dz
        assert dz is not None
        if not tape.param_grads:         # no input gradient here either (the input is data): nothing to do
            z.db_done = False
            z.dz = None
            return
        gw, acc = tape.pgrad(w)
        dwp = torch.empty((Cout, Kp, KH, ng), dtype=torch.float32, device=tape.device)
        need_db = b is not None and not z.db_done        # else: delivered by bn_act's backward (fcd_bn_bwd_finalize)
        z.db_done = False
        dbt = torch.empty((Cout,), dtype=torch.float32, device=tape.device) if need_db else None
        nbytes = _lib.load().fcd_conv2d_taps_wgrad_workspace(Kp, Cout_p, KH, ng)
        ws = _ws(tape.device, nbytes)
        _call("fcd_conv2d_taps_wgrad", xp.p_hi(), xp.p_lo(), Kp, xp.H, xp.Wp, dz.p_hi(), dz.p_lo(), dz.ld, OH, OW,
              dwp.data_ptr(), _lib.ptr(dbt), N, Kp, Kp, Cout, Cout_p, KH, ng, -pad, -pad + M, sstep, 0, ws.data_ptr(), nbytes,
              tag=f"conv_wgrad_tc_{form} {shape}", flops=flops)
        g = _w_unrowpack_in(dwp, C, KW) if row else _w_unpack4_in(dwp, C, KW)
        gw.add_(g) if acc else gw.copy_(g)
        if need_db:
            gb, accb = tape.pgrad(b)
            gb.add_(dbt) if accb else gb.copy_(dbt)
        z.dz = None

    tape.push(backward)
    return z


def conv_im2col_s2(tape: Tape, x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor]) -> Z:
    """3x3 / stride 2 / pad 1 convolution of an NCHW DATA tensor with few channels (the first discriminator layer,
    Module.py:196: 13 -> 64): the staging kernel writes every output pixel's 3x3xC receptive field as one K = 9C row
    (padded to a multiple of 64), so forward and weight gradient are single 1x1 GEMMs over a quarter of the pixels
    instead of 9 taps over a 64-channel zero-padded tensor.  No input gradient (callers use `conv` when x needs one).
    `x` may be a tuple of equally shaped tensors (the two siamese branches, Module.py:219-220): they become one batch."""
    Cout, C, KH, KW = w.shape
    xs = tuple(x) if isinstance(x, (tuple, list)) else (x,)     # several tensors of one shape: staged into ONE batch
    x = xs[0]
    assert (KH, KW) == (3, 3) and all(t.shape == x.shape for t in xs) and x.shape[1] == C
    Nx, _, H, W = x.shape
    N = Nx * len(xs)
    OH, OW = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    K = 9 * C
    Kp, Cout_p = pad_ch(K), pad_ch(Cout)
    a = tape.new_act(N, OH, OW, K, Cp=Kp, name="im2col")
    for i, t in enumerate(xs):
        t = t.contiguous()
        _call("fcd_stage_im2col3x3s2", t.data_ptr(), Nx, C, H, W, a.hi[i * Nx:].data_ptr(),
              None if a.lo is None else a.lo[i * Nx:].data_ptr(), Kp)
    wf = _derived(w, "im2col", lambda t: t.permute(0, 2, 3, 1).reshape(Cout, K, 1, 1).contiguous())
    w_hi, w_lo = _packed(wf, Cout_p, Kp, 0, "im2col")
    bias = None if b is None else _padded_vec(b, Cout_p)
    zt = torch.empty((N, OH, OW, Cout_p), dtype=torch.float32, device=tape.device)
    z = Z(zt, N, OH, OW, Cout, Cout_p)
    flops = 2.0 * N * OH * OW * Cout * C * 9
    shape = f"3x3s2 {C}->{Cout}"
    _call("fcd_conv2d_fwd", a.p_hi(), a.p_lo(), a.ld, w_hi.data_ptr(), _lib.ptr(w_lo), _lib.ptr(bias), None, 0,
          zt.data_ptr(), z.ld, N, OH, OW, Kp, Cout_p, 1, 1, 1, 0, None, None, _cfg["engine"],
          tag=f"conv_fwd_tc_im2col {shape}", flops=flops)
    z.bias_param = b

    def backward(tape):
        dz = z.dz
        assert dz is not None
        if not tape.param_grads:
            z.db_done = False
            z.dz = None
            return
        gw, acc = tape.pgrad(w)
        need_db = b is not None and not z.db_done
        z.db_done = False
        dwp = torch.empty((Cout, K, 1, 1), dtype=torch.float32, device=tape.device)
        dbt = torch.empty((Cout,), dtype=torch.float32, device=tape.device) if need_db else None
        nbytes = _lib.load().fcd_conv2d_wgrad_workspace(N, OH, OW, Kp, Cout_p, 1, 1, 1, 0, _cfg["engine"])
        ws = _ws(tape.device, nbytes)
        _call("fcd_conv2d_wgrad", a.p_hi(), a.p_lo(), a.ld, dz.p_hi(), dz.p_lo(), dz.ld, dwp.data_ptr(), _lib.ptr(dbt),
              N, OH, OW, K, Kp, Cout, Cout_p, 1, 1, 1, 0, 0, ws.data_ptr(), nbytes, _cfg["engine"],
              tag=f"conv_wgrad_tc_im2col {shape}", flops=flops)
        g = dwp.view(Cout, 3, 3, C).permute(0, 3, 1, 2)
        gw.add_(g) if acc else gw.copy_(g)
        if need_db:
            gb, accb = tape.pgrad(b)
            gb.add_(dbt) if accb else gb.copy_(dbt)
        z.dz = None

    tape.push(backward)
    return z


def conv_small_out(tape: Tape, x: Act, w: torch.Tensor, b: torch.Tensor, pad: int) -> Z:
    """Stride-1 convolution whose OUTPUT has <= 16 channels (Generator tail Module.py:158): the 64-wide N dimension
    of the MMA produces FOUR adjacent output pixels x 16 channel slots per row (TMA reads every 4th input pixel), the
    output is a plain NHWC tensor with 16 channel slots; dgrad / wgrad run on the channel-packed output gradient (4-pixel
    form, or the tight row-packed form when row_pack_pixels says so: one tap per filter row).  Requires OW % 4 == 0."""
    Cout, Cin, KH, KW = w.shape
    assert Cin == x.C and Cout <= 16
    N, H, W = x.N, x.H, x.W
    OH, OW = H + 2 * pad - KH + 1, W + 2 * pad - KW + 1
    assert OW % 4 == 0 and pad <= PACK_M
    M = PACK_M
    P = row_pack_pixels(Cout, KW)
    if P != 4 and (KW - 1) - M - pad > 0:      # the row-packed weight gradient drops packed columns left of 0: fine only
        P = 4                                  # while the x column they pair with is out of bounds too
    wq = _derived(w, "pack4_out", _w_pack4_out)
    w_hi, w_lo = _packed(wq, 64, x.Cp, 0, "pack4_out")
    bias = _derived(b, "pack4_bias", lambda t: torch.cat([_padded_vec(t, 16)] * 4))
    zt = torch.empty((N, OH, OW, 16), dtype=torch.float32, device=tape.device)
    z = Z(zt, N, OH, OW, Cout, 16)
    z.pack_m, z.pack_p = M, P
    flops = 2.0 * N * OH * OW * Cout * Cin * KH * KW
    shape = f"{KH}x{KW}s1 {Cin}->{Cout}"
    _call("fcd_conv2d_taps_fwd", x.p_hi(), x.p_lo(), x.ld, H, W, w_hi.data_ptr(), _lib.ptr(w_lo), bias.data_ptr(), None, 0,
          zt.data_ptr(), 64, N, OH, OW // 4, x.Cp, 64, KH, KW + 3, -pad, 1, -pad, 1, 1, 4,
          tag=f"conv_fwd_tc_pack4 {shape}", flops=flops)

    def backward(tape):
        dzp = z.dz
        assert isinstance(dzp, PackedAct) and dzp.P == P, "conv_small_out backward: packed output gradient missing"
        dev = tape.device
        gw, acc = tape.pgrad(w)
        gb, accb = tape.pgrad(b)
        row = P != 4
        Kp = dzp.Kp
        ng, sstep = (1, 1) if row else ((KW + 3) // 4, 4)
        form = "rowpack" if row else "pack4"
        # wgrad, 4-pixel form: T[(r,g)][ci][j*16+co] = sum x[oh+r-pad, ow''+4g+3-M-pad, ci] * dzp[oh, ow'', j*16+co] = dw[co][ci][r][4g+3-j]
        # row-packed form:     T[r][ci][j*Cout+co]   = sum x[oh+r-pad, ow''+KW-1-M-pad, ci] * dzp[oh, ow'', j*Cout+co] = dw[co][ci][r][KW-1-j]
        dwp = torch.empty((Kp, Cin, KH, ng), dtype=torch.float32, device=dev)
        dbk = torch.empty((Kp,), dtype=torch.float32, device=dev)
        nbytes = _lib.load().fcd_conv2d_taps_wgrad_workspace(x.Cp, Kp, KH, ng)
        ws = _ws(dev, nbytes)
        _call("fcd_conv2d_taps_wgrad", x.p_hi(), x.p_lo(), x.ld, H, W, dzp.p_hi(), dzp.p_lo(), Kp, OH, dzp.Wp, dwp.data_ptr(),
              dbk.data_ptr(), N, Cin, x.Cp, Kp, Kp, KH, ng, -pad, (KW - 1 if row else 3) - M - pad, sstep, 0, ws.data_ptr(), nbytes,
              tag=f"conv_wgrad_tc_{form} {shape}", flops=flops)
        if row:
            g = _w_unrowpack_out(dwp, Cout, KW)
        else:
            g = dwp.view(4, 16, Cin, KH, ng).flip(0).permute(1, 2, 3, 4, 0).reshape(16, Cin, KH, ng * 4)[:Cout, :, :, :KW]
        gw.add_(g) if acc else gw.copy_(g)
        gb.add_(dbk[:Cout]) if accb else gb.copy_(dbk[:Cout])        # pack slot j = 0 sees every pixel once
        # dgrad: dx[h,w,ci] = sum_{r',g} dzp[h+r'-ph, w+sstep*g-pw+M, :] . W''[ci][:][r'][g],  W'' = pack_in(flip(w)^T)
        wd = _derived(w, form + "_dgrad", lambda t: (_w_rowpack_in if row else _w_pack4_in)(t.flip(2, 3).permute(1, 0, 2, 3)))
        wd_hi, wd_lo = _packed(wd, x.Cp, Kp, 0, form + "_dgrad")
        gx = x.grad
        addend = gx.data_ptr() if x.ready else None
        _call("fcd_conv2d_taps_fwd", dzp.p_hi(), dzp.p_lo(), Kp, OH, dzp.Wp, wd_hi.data_ptr(), _lib.ptr(wd_lo), None, addend, x.ld,
              gx.data_ptr(), x.ld, N, H, W, Kp, x.Cp, KH, ng, -(KH - 1 - pad), 1, -(KW - 1 - pad) + M, sstep, 1, 1,
              tag=f"conv_dgrad_tc_{form} {shape}", flops=flops)
        x.mark_ready()
        z.dz = None

    tape.push(backward)
    return z


class BN:
    """Parameter / buffer bundle of one nn.BatchNorm2d."""

    __slots__ = ("weight", "bias", "running_mean", "running_var", "num_batches_tracked")

    def __init__(self, weight, bias, running_mean, running_var, num_batches_tracked):
        self.weight, self.bias = weight, bias
        self.running_mean, self.running_var, self.num_batches_tracked = running_mean, running_var, num_batches_tracked


def bn_act(tape: Tape, z: Z, bn: Optional[BN], training: bool, act: int, slope: Optional[torch.Tensor] = None,
           slope_const: float = 0.0, residual: Optional[Act] = None, out: Optional[Act] = None, groups: int = 1) -> Act:
    """[BatchNorm2d] -> activation -> [+ residual], written as a split activation (K6/K7).
    BatchNorm statistics are those of THIS call (per siamese branch).  `groups` > 1: the batch holds that
    many equally sized calls of the reference back to back (the two branches of Module.py:219-220 run as one batch through
    the convolutions); every group gets its own batch statistics, running-statistics update and backward sums, in call order."""
    C, Cp, G = z.C, z.Cp, groups
    assert z.N % G == 0
    Ng, npix = z.N // G, z.npix // G          # images / pixels per statistics group
    dev = tape.device
    if G > 1:
        assert residual is None and out is None and _cfg["sync_bn"] is None, "bn_act: grouped form is plain BN + activation"
    if out is None:
        out = tape.new_act(z.N, z.H, z.W, C, Cp)
    zp = [z.t[g * Ng:].data_ptr() for g in range(G)]
    o_hi = [out.hi[g * Ng:].data_ptr() for g in range(G)]
    o_lo = [None if out.lo is None else out.lo[g * Ng:].data_ptr() for g in range(G)]
    vec = None
    gstats = None
    if bn is not None:
        vec = torch.empty((G, 6, Cp), dtype=torch.float32, device=dev)  # per group: scale, shift, mean, invstd, c1, c2
        count = float(npix)
        if training:
            if G > 1:
                gstats = tape.zeros64(G, 2, Cp)
                for g in range(G):
                    _call("fcd_bn_stats", zp[g], z.ld, npix, Cp, gstats[g, 0].data_ptr(), gstats[g, 1].data_ptr())
            else:
                assert z.sum is not None
                if _cfg["sync_bn"] is not None and z.sum_local is None:      # statistics of the whole (all-rank) batch
                    z.sum_local = z.sum.clone()
                    _all_reduce_sum(z.stats)
                count = float(npix) * _sync_world()
        for g in range(G):
            gsum, gsq = (gstats[g, 0], gstats[g, 1]) if gstats is not None else (z.sum, z.sqsum)
            _call("fcd_bn_finalize", _lib.ptr(gsum), _lib.ptr(gsq), count, bn.weight.data_ptr(),
                  bn.bias.data_ptr(), bn.running_mean.data_ptr(), bn.running_var.data_ptr(), C, Cp, BN_MOMENTUM, BN_EPS,
                  1 if training else 0, vec[g, 0].data_ptr(), vec[g, 1].data_ptr(), vec[g, 2].data_ptr(), vec[g, 3].data_ptr())
        if training and bn.num_batches_tracked is not None:
            bn.num_batches_tracked.add_(G)
    sp = None if slope is None else slope.data_ptr()
    for g in range(G):
        _call("fcd_bn_act_fwd", zp[g], z.ld, None if vec is None else vec[g, 0].data_ptr(),
              None if vec is None else vec[g, 1].data_ptr(), act, sp, slope_const,
              None if residual is None else residual.p_hi(), None if residual is None else residual.p_lo(),
              0 if residual is None else residual.ld, o_hi[g], o_lo[g], out.ld, npix, Cp)

    def backward(tape):
        assert out.ready, f"bn_act backward: gradient of {out.name} missing"
        da = out.grad
        dz = Act.empty(z.N, z.H, z.W, C, dev, Cp)
        need_reduce = bn is not None or act == ACT_PRELU
        red_all = tape.zeros64(G, 2, Cp + 8) if need_reduce else None
        c_all = vec if vec is not None else (torch.empty((G, 6, Cp), dtype=torch.float32, device=dev) if need_reduce else None)
        for g in range(G):
            dap = da[g * Ng:].data_ptr()
            v = (lambda i: vec[g, i].data_ptr()) if vec is not None else (lambda i: None)
            if need_reduce:
                red = red_all[g]
                s1, s2, ds = red[0, :Cp], red[1, :Cp], red[0, Cp:]
                _call("fcd_bn_act_bwd_reduce", dap, out.ld, zp[g], z.ld, v(0), v(1), v(2), v(3), act, sp,
                      slope_const, npix, Cp, s1.data_ptr(), s2.data_ptr(), ds.data_ptr() if act == ACT_PRELU else None)
                dgam = dbet = dsl = None
                acc = 0
                if bn is not None:
                    dgam, acc = tape.pgrad(bn.weight)
                    dbet, _ = tape.pgrad(bn.bias)
                if act == ACT_PRELU:
                    dsl, acc_s = tape.pgrad(slope)
                    assert bn is None or acc_s == acc
                    acc = acc_s
                c = c_all[g]
                # the producing convolution's bias gradient (= per-channel sum of dz) comes out of the same sums in closed form
                gb = acc_b = None
                if z.bias_param is not None:
                    gb, acc_b = tape.pgrad(z.bias_param)
                    z.db_done = True
                gsum = gstats[g, 0] if gstats is not None else z.sum
                use_stats = bn is not None and training and gsum is not None
                zsum = gsum
                gl = None
                if use_stats and z.sum_local is not None:        # SyncBN: dz needs the means over ALL ranks' batches
                    zsum = z.sum_local
                    gl = red[:, :Cp].contiguous()
                    _all_reduce_sum(gl)
                _call("fcd_bn_bwd_finalize", s1.data_ptr(), s2.data_ptr(), float(npix), 1 if (training and bn is not None) else 0,
                      C, Cp, c[4].data_ptr(), c[5].data_ptr(), _lib.ptr(dgam), _lib.ptr(dbet), acc,
                      ds.data_ptr() if act == ACT_PRELU else None, _lib.ptr(dsl),
                      v(0), zsum.data_ptr() if use_stats else None, v(2) if use_stats else None, v(3) if use_stats else None,
                      _lib.ptr(gb), acc_b or 0, None if gl is None else gl[0].data_ptr(), None if gl is None else gl[1].data_ptr(),
                      float(npix) * _sync_world())
            _call("fcd_bn_act_bwd_apply", dap, out.ld, zp[g], z.ld, v(0), v(1), v(2), v(3), v(4), v(5), act,
                  sp, slope_const, dz.hi[g * Ng:].data_ptr(), None if dz.lo is None else dz.lo[g * Ng:].data_ptr(), dz.ld, npix, Cp)
        z.dz = dz
        if DEBUG_CAPTURE is not None:
            DEBUG_CAPTURE.append(("da", da[..., :C].permute(0, 3, 1, 2).clone()))
            j = dz.hi.float() + (dz.lo.float() if dz.lo is not None else 0)
            DEBUG_CAPTURE.append(("dz", j[..., :C].permute(0, 3, 1, 2).clone()))
        if residual is not None:
            if residual.ready:
                _call("fcd_add_f32", residual.grad.data_ptr(), residual.ld, da.data_ptr(), out.ld, z.npix, Cp)
            elif residual.parent is None and out.parent is None:
                residual._grad = da          # alias: `out.grad` is dead after this closure
                residual.mark_ready()
            else:
                residual.grad.copy_(da)
                residual.mark_ready()
        elif out.parent is None:
            out.reset_grad()             # release the consumed gradient buffer early

    tape.push(backward)
    return out


def maxpool2(tape: Tape, x: Act) -> Act:
    """nn.MaxPool2d(2) (Module.py:44)."""
    out = tape.new_act(x.N, x.H // 2, x.W // 2, x.C, x.Cp)
    _call("fcd_maxpool2_fwd", x.p_hi(), x.p_lo(), x.ld, x.N, x.H, x.W, x.Cp, out.p_hi(), out.p_lo(), out.ld)

    def backward(tape):
        assert out.ready
        g = x.grad
        _call("fcd_maxpool2_bwd", out.grad.data_ptr(), out.ld, x.p_hi(), x.p_lo(), x.ld, x.N, x.H, x.W, x.Cp,
              g.data_ptr(), x.ld, 1 if x.ready else 0)
        x.mark_ready()

    tape.push(backward)
    return out


def upsample2x_into(tape: Tape, x: Act, dst: Act) -> None:
    """nn.Upsample(x2, bilinear, align_corners=True) + F.pad + cat slot (Module.py:60,70-78)."""
    dY, dX = dst.H - 2 * x.H, dst.W - 2 * x.W
    pt, pl = dY // 2, dX // 2
    assert dY >= 0 and dX >= 0 and dst.Cp == x.Cp
    _call("fcd_upsample2x_bilinear_fwd", x.p_hi(), x.p_lo(), x.ld, x.N, x.H, x.W, x.Cp, dst.p_hi(), dst.p_lo(), dst.ld,
          dst.H, dst.W, pt, pl)

    def backward(tape):
        assert dst.ready and not x.ready
        _call("fcd_upsample2x_bilinear_bwd", dst.grad.data_ptr(), dst.ld, x.N, x.H, x.W, x.Cp, dst.H, dst.W, pt, pl,
              x.grad.data_ptr(), x.ld)
        x.mark_ready()

    tape.push(backward)


def conv_transpose2x2_into(tape: Tape, x: Act, w: torch.Tensor, b: torch.Tensor, dst: Act) -> None:
    """nn.ConvTranspose2d(Cin, Cout, 2, 2) (Module.py:63) + F.pad + cat slot: four 1x1 tcgen05 convolutions
    (one per output sub-pixel) and a pixel-shuffle store."""
    Cin, Cout = w.shape[0], w.shape[1]
    assert x.C == Cin and dst.Cp == pad_ch(Cout)
    N, h, wd = x.N, x.H, x.W
    Cout_p = dst.Cp
    dY, dX = dst.H - 2 * h, dst.W - 2 * wd
    pt, pl = dY // 2, dX // 2
    planes = torch.empty((4, N, h, wd, Cout_p), dtype=torch.float32, device=tape.device)
    # (Cin, Cout, 2, 2) -> four OIHW (Cout, Cin, 1, 1) matrices; tiny weight-side reshuffle
    w4 = w.detach().permute(2, 3, 1, 0).reshape(4, Cout, Cin, 1, 1).contiguous()
    bias = _padded_vec(b, Cout_p)
    packed = []
    for s in range(4):
        hi = torch.empty((1, Cout_p, x.Cp), dtype=torch.bfloat16, device=tape.device)
        lo = torch.empty_like(hi) if _cfg["split"] else None
        _call("fcd_pack_conv_weight", w4[s].data_ptr(), Cout, Cin, 1, 1, Cout_p, x.Cp, 0, hi.data_ptr(), _lib.ptr(lo))
        packed.append((hi, lo))
        _call("fcd_conv2d_fwd", x.p_hi(), x.p_lo(), x.ld, hi.data_ptr(), _lib.ptr(lo), bias.data_ptr(), None, 0,
              planes[s].data_ptr(), Cout_p, N, h, wd, x.Cp, Cout_p, 1, 1, 1, 0, None, None, _cfg["engine"])
    _call("fcd_convT2x2_shuffle_fwd", planes.data_ptr(), planes.stride(0), Cout_p, N, h, wd, Cout_p, dst.p_hi(),
          dst.p_lo(), dst.ld, dst.H, dst.W, pt, pl)
    del planes

    def backward(tape):
        assert dst.ready and not x.ready
        dev = tape.device
        g_hi = torch.empty((4, N, h, wd, Cout_p), dtype=torch.bfloat16, device=dev)
        g_lo = torch.empty_like(g_hi) if _cfg["split"] else None
        _call("fcd_convT2x2_shuffle_bwd", dst.grad.data_ptr(), dst.ld, N, h, wd, Cout_p, dst.H, dst.W, pt, pl,
              g_hi.data_ptr(), _lib.ptr(g_lo), g_hi.stride(0), Cout_p)
        dw4 = torch.empty((4, Cout, Cin), dtype=torch.float32, device=dev)
        db4 = torch.empty((4, Cout), dtype=torch.float32, device=dev)
        gx = x.grad
        nbytes = _lib.load().fcd_conv2d_wgrad_workspace(N, h, wd, x.Cp, Cout_p, 1, 1, 1, 0, _cfg["engine"])
        ws = _ws(dev, nbytes)
        for s in range(4):
            ghi = g_hi[s].data_ptr()
            glo = None if g_lo is None else g_lo[s].data_ptr()
            _call("fcd_conv2d_wgrad", x.p_hi(), x.p_lo(), x.ld, ghi, glo, Cout_p, dw4[s].data_ptr(), db4[s].data_ptr(), N,
                  h, wd, Cin, x.Cp, Cout, Cout_p, 1, 1, 1, 0, 0, ws.data_ptr(), nbytes, _cfg["engine"])
            wd_hi = torch.empty((1, x.Cp, Cout_p), dtype=torch.bfloat16, device=dev)
            wd_lo = torch.empty_like(wd_hi) if _cfg["split"] else None
            _call("fcd_pack_conv_weight", w4[s].data_ptr(), Cout, Cin, 1, 1, Cout_p, x.Cp, 1, wd_hi.data_ptr(),
                  _lib.ptr(wd_lo))
            _call("fcd_conv2d_fwd", ghi, glo, Cout_p, wd_hi.data_ptr(), _lib.ptr(wd_lo), None,
                  gx.data_ptr() if s > 0 else None, x.ld, gx.data_ptr(), x.ld, N, h, wd, Cout_p, x.Cp, 1, 1, 1, 0, None,
                  None, _cfg["engine"])
        x.mark_ready()
        gw, acc = tape.pgrad(w)
        gb, _ = tape.pgrad(b)
        dw = dw4.reshape(2, 2, Cout, Cin).permute(3, 2, 0, 1)   # -> (Cin, Cout, 2, 2)
        if acc:
            gw.add_(dw)
            gb.add_(db4.sum(0))
        else:
            gw.copy_(dw)
            gb.copy_(db4.sum(0))

    tape.push(backward)


def outconv_sigmoid(tape: Tape, x: Act, w: torch.Tensor, b: torch.Tensor, grad_slot: dict) -> torch.Tensor:
    """OutConv (Module.py:82-90): 1x1 conv + sigmoid -> NCHW fp32 density map.  `grad_slot['dout']` must hold the
    NCHW output gradient when the tape is replayed."""
    n_out, Cin = w.shape[0], w.shape[1]
    assert Cin == x.C
    out = torch.empty((x.N, n_out, x.H, x.W), dtype=torch.float32, device=tape.device)
    w2 = w.reshape(n_out, Cin)
    _call("fcd_outconv_sigmoid_fwd", x.p_hi(), x.p_lo(), x.ld, Cin, w2.data_ptr(), b.data_ptr(), n_out, x.N, x.H, x.W,
          out.data_ptr())
    result, out = out, out.detach()   # the closure keeps a grad_fn-free alias: capturing the RETURNED tensor would close
                                      # the cycle output -> grad_fn -> tape -> closure -> output and defer freeing to the GC

    def backward(tape):
        dout = grad_slot["dout"].contiguous()
        gw, acc = tape.pgrad(w)
        gb, _ = tape.pgrad(b)
        scratch = torch.empty(n_out * Cin + n_out, dtype=torch.float64, device=tape.device)
        assert not x.ready
        _call("fcd_outconv_sigmoid_bwd", dout.data_ptr(), out.data_ptr(), x.p_hi(), x.p_lo(), x.ld, Cin, w2.data_ptr(),
              n_out, x.N, x.H, x.W, x.grad.data_ptr(), x.ld, gw.data_ptr(), gb.data_ptr(), acc, scratch.data_ptr())
        x.mark_ready()

    tape.push(backward)
    return result


def disc_head(tape: Tape, fx: Act, fy: Act, w1, b1, w2, b2, grad_slot: dict) -> torch.Tensor:
    """Discriminator classifier on (fx - fy) (Module.py:212-223): GAP -> 1x1(512->1024) -> LeakyReLU(0.2) ->
    1x1(1024->1) -> sigmoid -> (B,)."""
    N, HW, C = fx.N, fx.H * fx.W, fx.C
    dev = tape.device
    O1 = w1.shape[0]
    pooled = torch.empty((N, C), dtype=torch.float32, device=dev)
    _call("fcd_gap_diff_fwd", fx.p_hi(), fx.p_lo(), fy.p_hi(), fy.p_lo(), fx.ld, N, HW, C, pooled.data_ptr())
    pre1 = torch.empty((N, O1), dtype=torch.float32, device=dev)
    h1 = torch.empty_like(pre1)
    _call("fcd_fc_fwd", pooled.data_ptr(), w1.data_ptr(), b1.data_ptr(), N, C, O1, 3, pre1.data_ptr(), h1.data_ptr())
    out = torch.empty((N,), dtype=torch.float32, device=dev)
    _call("fcd_fc_fwd", h1.data_ptr(), w2.data_ptr(), b2.data_ptr(), N, O1, 1, 4, None, out.data_ptr())
    result, out = out, out.detach()   # see outconv_sigmoid: no reference cycle through the returned tensor

    def backward(tape):
        dout = grad_slot["dout"].contiguous()
        gw2, a2 = tape.pgrad(w2)
        gb2, _ = tape.pgrad(b2)
        gw1, a1 = tape.pgrad(w1)
        gb1, _ = tape.pgrad(b1)
        du2 = torch.empty((N,), dtype=torch.float32, device=dev)
        dh1 = torch.empty((N, O1), dtype=torch.float32, device=dev)
        _call("fcd_fc_bwd", dout.data_ptr(), None, out.data_ptr(), h1.data_ptr(), w2.data_ptr(), N, O1, 1, 4,
              du2.data_ptr(), dh1.data_ptr(), gw2.data_ptr(), gb2.data_ptr(), a2)
        du1 = torch.empty((N, O1), dtype=torch.float32, device=dev)
        dpooled = torch.empty((N, C), dtype=torch.float32, device=dev)
        _call("fcd_fc_bwd", dh1.data_ptr(), pre1.data_ptr(), h1.data_ptr(), pooled.data_ptr(), w1.data_ptr(), N, C, O1, 3,
              du1.data_ptr(), dpooled.data_ptr(), gw1.data_ptr(), gb1.data_ptr(), a1)
        assert not fx.ready and not fy.ready and fx.ld == fy.ld
        _call("fcd_gap_diff_bwd", dpooled.data_ptr(), N, HW, C, fx.grad.data_ptr(), fy.grad.data_ptr(), fx.ld)
        fx.mark_ready()
        fy.mark_ready()

    tape.push(backward)
    return result


class BatchHalf:
    """Gradient view of images [n0, n0 + n) of a staged batch (what `unstage_grad` needs of an input activation)."""

    def __init__(self, a: Act, n0: int, n: int):
        self.a, self.n0 = a, n0
        self.N, self.C, self.H, self.W, self.ld, self.hi = n, a.C, a.H, a.W, a.ld, a.hi

    @property
    def grad(self) -> torch.Tensor:
        return self.a.grad[self.n0:self.n0 + self.N]


def stage_two_inputs(tape: Tape, x: torch.Tensor, y: torch.Tensor) -> Act:
    """Two NCHW fp32 tensors of equal shape -> ONE split NHWC activation holding x in the first half of the batch and y in the
    second (the perception loss runs its frozen VGG16 on target and generated images in one pass)."""
    N, C, H, W = x.shape
    assert y.shape == x.shape
    x, y = x.contiguous(), y.contiguous()
    a = tape.new_act(2 * N, H, W, C, Cp=pad_ch(C), name="input2")
    lo = a.lo
    _call("fcd_stage_nchw_to_split", x.data_ptr(), None, N, C, H, W, a.hi[:N].data_ptr(), None if lo is None else lo[:N].data_ptr(),
          a.ld, a.Cp)
    _call("fcd_stage_nchw_to_split", y.data_ptr(), None, N, C, H, W, a.hi[N:].data_ptr(), None if lo is None else lo[N:].data_ptr(),
          a.ld, a.Cp)
    return a


def mse_halves(tape: Tape, f: Act, weight: float, acc: torch.Tensor, grad_slot: dict) -> float:
    """nn.MSELoss between the two batch halves of a feature tensor (Loss.py:36,48,59): `acc` (double[1], zero-initialised)
    receives the sum of squared differences; returns the factor weight / numel that turns it into the weighted mean.
    Backward adds gout * 2 * weight / numel * (a - b) to the first half's gradient and its negative to the second's."""
    assert f.N % 2 == 0 and f.parent is None
    npix = (f.N // 2) * f.H * f.W
    half = npix * f.ld
    coef = weight / float(npix * f.C)
    _call("fcd_mse_halves_fwd", f.p_hi(), f.p_lo(), f.ld, half, npix, f.Cp, acc.data_ptr())

    def backward(tape):
        gout = grad_slot["dout"].contiguous()
        g = f.grad
        _call("fcd_mse_halves_bwd", f.p_hi(), f.p_lo(), f.ld, half, npix, f.Cp, gout.data_ptr(), 2.0 * coef, g.data_ptr(), f.ld,
              half, 1 if f.ready else 0)
        f.mark_ready()

    tape.push(backward)
    return coef


def z_to_nchw(tape: Tape, z: Z, grad_slot: dict) -> torch.Tensor:
    """Final convolution output (no BN / activation) -> NCHW fp32 boundary tensor (Generator block8, Module.py:168)."""
    out = torch.empty((z.N, z.C, z.H, z.W), dtype=torch.float32, device=tape.device)
    _call("fcd_unstage_f32_to_nchw", z.t.data_ptr(), z.ld, z.N, z.C, z.H, z.W, out.data_ptr(), 0)

    def backward(tape):
        dout = grad_slot["dout"].contiguous()
        if z.pack_m:
            z.dz = PackedAct(dout, z.pack_m, z.pack_p)
            return
        dz = Act.empty(z.N, z.H, z.W, z.C, tape.device, z.Cp)
        _call("fcd_stage_nchw_to_split", dout.data_ptr(), None, z.N, z.C, z.H, z.W, dz.p_hi(), dz.p_lo(), dz.ld, dz.Cp)
        z.dz = dz

    tape.push(backward)
    return out


# --------------------------------------------------------------------------------------------------
# autograd bridge
# --------------------------------------------------------------------------------------------------
class Runner:
    """(forward recorder, the nn.Parameters it uses) — parameter gradients are matched by object identity."""

    def __init__(self, fn, params):
        self.fn, self.params = fn, list(params)


def run_net(module: torch.nn.Module, fn, *inputs) -> torch.Tensor:
    params = [p for p in module.parameters()]
    return NetFunction.apply(Runner(fn, params), len(inputs), *inputs, *params)


class NetFunction(torch.autograd.Function):
    """One network forward = one autograd node.  `runner(tape, inputs, needs_input_grad) -> (outputs, finish)`
    launches the forward kernels and records the backward closures; `finish(grad_outputs) -> input grads`
    is called after the tape has been replayed."""

    @staticmethod
    def forward(ctx, runner, n_inputs: int, *tensors):
        inputs = tensors[:n_inputs]
        params = tensors[n_inputs:]
        for t in inputs:
            if not (t.is_cuda and t.dtype == torch.float32):
                raise _lib.FcdError("fcdgan_b200 networks take fp32 CUDA tensors (there is no CPU path)")
        record = any(ctx.needs_input_grad[2:])
        tape = Tape(inputs[0].device, record)
        tape.param_grads = any(ctx.needs_input_grad[2 + n_inputs:])
        with torch.cuda.device(inputs[0].device):       # kernels launch on the TENSORS' device, whatever device is current
            outs, slot, input_acts = runner.fn(tape, inputs, ctx.needs_input_grad[2:2 + n_inputs])
        ctx.tape, ctx.slot, ctx.input_acts = tape, slot, input_acts
        ctx.n_inputs = n_inputs
        ctx.param_ids = [id(p) for p in runner.params]
        assert len(runner.params) == len(params)
        return outs

    @staticmethod
    def backward(ctx, gout):
        tape = ctx.tape
        ctx.slot["dout"] = gout
        need_in = ctx.needs_input_grad[2:2 + ctx.n_inputs]
        # input gradients must be extracted before Tape.run() resets the gradient buffers -> do it as a final op
        results = {}

        def grab(_tape):
            for i, a in enumerate(ctx.input_acts):
                if need_in[i]:
                    results[i] = unstage_grad(a)

        tape.ops.insert(0, grab)
        try:
            with torch.cuda.device(tape.device):
                pg = tape.run()
        finally:
            tape.ops.pop(0)
            ctx.slot["dout"] = None
        gin = [results.get(i) for i in range(ctx.n_inputs)]
        gp = []
        for k, pid in enumerate(ctx.param_ids):
            g = pg.get(pid)
            gp.append(g if (g is not None and ctx.needs_input_grad[2 + ctx.n_inputs + k]) else None)
        return (None, None, *gin, *gp)
