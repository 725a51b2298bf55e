"""Rasters either side of the hot path, resident on the GPU.

The reference walks a large bi-temporal raster in overlapping patches through GDAL + numpy on the host
(`GDALDataset`, data_utils.py:14-213), normalises every band in a Python loop (`NORMALIZE`, CommonFunc.py:199-224),
copies each predicted change-density tile back to the host to write its centre crop (`GDALwriteDefault`), and thresholds
/ counts the confusion matrix per sample in numpy (Demo_USSS.py:349-362, metrics.py:74-80).  Here the two rasters are
uploaded once and everything between "raster" and "tile batch" is one kernel launch each way:

    grid = TileGrid(xsize, ysize, patch_size=(220, 220), overlap_padding=(10, 10))     # GDALDataset geometry
    pair = RasterPair(x, y, grid, ref=ref)                # x, y: [C][H][W] uint8/uint16/int16/float32, numpy or torch
    meanX, stdX, meanY, stdY = pair.meanstd()             # Dataset_meanstd, on the device
    xt, yt, rt = pair.tiles(items, (meanX, stdX, meanY, stdY))      # == default-collated GDALDataset.__getitem__
    pair.write_default(cmap, items)                       # GDALwriteDefault -> pair.out [H][W] float32
    acc = Evaluator(2); acc.add_batch_map(rt, cmap, grid, items, prob_thresh, gt_map, pre_map)

File I/O (GDAL) stays with the caller: hand in arrays, read `pair.out` back.  No CPU fallback: device work goes through
libfcd_b200.so and raises if it is missing.
"""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np
import torch

from . import _lib
from .engine import _raw_stream

_DTYPES = {torch.float32: 0, torch.uint16: 1, torch.int16: 2, torch.uint8: 3}


class TileGrid:
    """Patch geometry of `GDALDataset` (data_utils.py:57-63, 139-176): same names, same `> 0` edge rules."""

    def __init__(self, xsize: int, ysize: int, patch_size=(200, 200), overlap_padding=(10, 10)):
        px, py = patch_size
        ox, oy = overlap_padding
        if px - 2 * ox <= 0 or py - 2 * oy <= 0:
            raise ValueError("patch_size must exceed twice the overlap_padding")
        self.xsize, self.ysize = int(xsize), int(ysize)
        self.patch_size, self.overlap_padding = (int(px), int(py)), (int(ox), int(oy))
        self.xstart = list(range(0, xsize, px - 2 * ox))
        self.xend = [x + px - 2 * ox for x in self.xstart if x + px - 2 * ox < xsize] + [xsize]
        self.ystart = list(range(0, ysize, py - 2 * oy))
        self.yend = [y + py - 2 * oy for y in self.ystart if y + py - 2 * oy < ysize] + [ysize]

    def __len__(self) -> int:
        return len(self.xstart) * len(self.ystart)

    def patch_count(self) -> Tuple[int, int]:
        return len(self.xstart), len(self.ystart)

    def item_xy(self, item: int) -> Tuple[int, int]:
        yc = len(self.ystart)
        return math.floor(item / yc), item % yc

    def slice_assign(self, item_x: int, item_y: int):
        pad = self.overlap_padding
        xstart, xend = self.xstart[item_x], self.xend[item_x]
        ystart, yend = self.ystart[item_y], self.yend[item_y]
        sl = (xstart, ystart, xend - xstart, yend - ystart)
        x_ori = 0 if xstart - pad[0] > 0 else pad[0]
        y_ori = 0 if ystart - pad[1] > 0 else pad[1]
        xstart = xstart - pad[0] if xstart - pad[0] > 0 else 0
        ystart = ystart - pad[1] if ystart - pad[1] > 0 else 0
        xend = xend + pad[0] if xend + pad[0] < self.xsize else self.xsize
        yend = yend + pad[1] if yend + pad[1] < self.ysize else self.ysize
        return sl, (xstart, ystart, xend - xstart, yend - ystart), (x_ori, y_ori, xend - xstart, yend - ystart)

    # ---- geometry tables for the kernels: int32 [n_tiles][6], built once, cached per device ---------------------------
    def gather_geom(self) -> np.ndarray:
        """row t = (read_x, read_y, read_w, read_h, write_x, write_y) of tile t."""
        if getattr(self, "_gather_geom", None) is None:
            g = np.empty((len(self), 6), dtype=np.int32)
            for it in range(len(self)):
                _, rd, wr = self.slice_assign(*self.item_xy(it))
                if wr[0] + rd[2] > self.patch_size[0] or wr[1] + rd[3] > self.patch_size[1]:
                    raise ValueError(f"tile {it}: read window {rd} does not fit the patch at offset {wr[:2]} "
                                     "(the reference raises a broadcast error here)")
                g[it] = (rd[0], rd[1], rd[2], rd[3], wr[0], wr[1])
            self._gather_geom = g
        return self._gather_geom

    def crop_geom(self) -> np.ndarray:
        """row t = (pad_x, pad_y, slice_x, slice_y, slice_w, slice_h) of tile t (the centre crop GDALwriteDefault keeps)."""
        if getattr(self, "_crop_geom", None) is None:
            g = np.empty((len(self), 6), dtype=np.int32)
            pad = self.overlap_padding
            for it in range(len(self)):
                sl, _, _ = self.slice_assign(*self.item_xy(it))
                g[it] = (pad[0], pad[1], sl[0], sl[1], sl[2], sl[3])
            self._crop_geom = g
        return self._crop_geom

    def device_tables(self, device) -> Tuple[torch.Tensor, torch.Tensor]:
        cache = self.__dict__.setdefault("_dev_tables", {})
        key = str(device)
        if key not in cache:
            cache[key] = (torch.from_numpy(self.gather_geom()).to(device), torch.from_numpy(self.crop_geom()).to(device))
        return cache[key]

    def device_items(self, items, device) -> torch.Tensor:
        """Tile indices of a batch as an int32 device tensor (range-checked on the host when they arrive from the host)."""
        if isinstance(items, torch.Tensor) and items.is_cuda:
            return items.to(device=device, dtype=torch.int32)
        lst = items.tolist() if isinstance(items, torch.Tensor) else [int(i) for i in items]
        if not lst or min(lst) < 0 or max(lst) >= len(self):
            raise IndexError(f"tile index out of range (grid has {len(self)} tiles)")
        return torch.tensor(lst, dtype=torch.int32).to(device)


def _as_device_raster(a, device) -> torch.Tensor:
    if isinstance(a, np.ndarray):
        if a.dtype == np.float64:
            a = a.astype(np.float32)
        a = torch.from_numpy(np.ascontiguousarray(a))
    if a.dtype not in _DTYPES:
        raise TypeError(f"unsupported raster dtype {a.dtype}; use uint8, uint16, int16 or float32")
    if a.dim() == 2:
        a = a[None]
    return a.contiguous().to(device)


class RasterPair:
    """Two co-registered rasters (+ optional reference map) resident in HBM, served as tile batches."""

    def __init__(self, x, y, grid: TileGrid, ref=None, device="cuda:0"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("RasterPair needs a CUDA device (no CPU fallback on the fcdgan_b200 path)")
        self.grid = grid
        self.x, self.y = _as_device_raster(x, self.device), _as_device_raster(y, self.device)
        if self.x.shape != self.y.shape:
            raise ValueError("Image sizes don't match")                       # data_utils.py:50-52
        C, H, W = self.x.shape
        if (W, H) != (grid.xsize, grid.ysize):
            raise ValueError("raster size does not match the tile grid")
        self.ref = None
        if ref is not None:
            self.ref = _as_device_raster(ref, self.device)
            if self.ref.shape != (1, H, W):
                raise ValueError("Reference sizes don't match image")         # data_utils.py:83-85
        self.out = torch.zeros((H, W), dtype=torch.float32, device=self.device)     # GDALwriteDefault's output band

    def size(self) -> Tuple[int, int, int]:
        C, H, W = self.x.shape
        return W, H, C

    def __len__(self) -> int:
        return len(self.grid)

    def _gather(self, raster: torch.Tensor, items_dev: torch.Tensor, mean=None, std=None) -> torch.Tensor:
        C, H, W = raster.shape
        B = items_dev.numel()
        geom_dev, _ = self.grid.device_tables(self.device)
        pw, ph = self.grid.patch_size
        out = torch.empty((B, C, ph, pw), dtype=torch.float32, device=self.device)
        stats = None
        if mean is not None:
            if len(mean) < C or len(std) < C:
                raise ValueError("The input channel doesn't match the stats list")     # CommonFunc.py:211-213
            key = (tuple(map(float, mean[:C])), tuple(map(float, std[:C])))
            cache = self.__dict__.setdefault("_stats_cache", {})
            if key not in cache:
                cache[key] = torch.tensor([list(key[0]), list(key[1])], dtype=torch.float64).to(self.device)
            stats = cache[key]
        _lib.call("fcd_tiles_gather", raster.data_ptr(), _DTYPES[raster.dtype], C, H, W, geom_dev.data_ptr(), items_dev.data_ptr(), B, pw, ph,
                  None if stats is None else stats[0].data_ptr(), None if stats is None else stats[1].data_ptr(),
                  out.data_ptr(), _raw_stream())
        return out

    def tiles(self, items, stats=None):
        """-> (x_tiles, y_tiles, ref_tiles): what a default-collated batch of `GDALDataset.__getitem__` holds
        (data_utils.py:94-137); stats = (meanX, stdX, meanY, stdY) applies `NORMALIZE` (switch 1 for x, 2 for y)."""
        it = self.grid.device_items(items, self.device)
        mX = sX = mY = sY = None
        if stats is not None:
            mX, sX, mY, sY = stats
        xt = self._gather(self.x, it, mX, sX)
        yt = self._gather(self.y, it, mY, sY)
        if self.ref is not None:
            rt = self._gather(self.ref, it)
        else:
            pw, ph = self.grid.patch_size
            rt = torch.zeros((it.numel(), 1, ph, pw), dtype=torch.float32, device=self.device)
        return xt, yt, rt

    def meanstd(self, batch: int = 64):
        """`Dataset_mean` + `Dataset_std` (CommonFunc.py:436-499) over the tiles of THIS grid (the reference uses a grid
        with overlap_padding (0, 0) for the statistics, Demo_USSS.py:88-93): per-tile statistics over the pixels whose
        band sum of x is non-zero, combined with weights npixel / N (mean) and npixel / (N - 1) (variance).
        -> (meanX, stdX, meanY, stdY) as Python float lists, like the reference."""
        C = self.x.shape[0]
        n = len(self.grid)
        pw, ph = self.grid.patch_size
        sums = torch.zeros((n, 2, C), dtype=torch.float64, device=self.device)
        counts = torch.zeros((n,), dtype=torch.int64, device=self.device)

        def sweep(centre, sums_, counts_):
            for s in range(0, n, batch):
                items = list(range(s, min(n, s + batch)))
                xt, yt, _ = self.tiles(items)
                _lib.call("fcd_tiles_moments", xt.data_ptr(), yt.data_ptr(), len(items), C, ph * pw,
                          None if centre is None else centre.data_ptr(), sums_[s].data_ptr(),
                          None if counts_ is None else counts_[s:].data_ptr(), _raw_stream())

        sweep(None, sums, counts)
        npix = counts.double()
        total = npix.sum()
        tile_mean = sums / npix.clamp(min=1)[:, None, None]
        mean = (tile_mean * (npix / total)[:, None, None]).sum(dim=0)            # [2][C]
        sq = torch.zeros_like(sums)
        sweep(mean.reshape(-1).contiguous(), sq, None)
        tile_var = sq / npix.clamp(min=1)[:, None, None]
        std = torch.sqrt((tile_var * (npix / (total - 1))[:, None, None]).sum(dim=0))
        # the reference's statistics are float32 torch tensors converted with .numpy().tolist() (CommonFunc.py:403-406)
        f32 = lambda t: t.float().cpu().numpy().tolist()
        return f32(mean[0]), f32(std[0]), f32(mean[1]), f32(std[1])

    def write_default(self, out_tiles: torch.Tensor, items) -> torch.Tensor:
        """`GDALwriteDefault` for a whole batch (data_utils.py:178-213): the centre crop of every [1][ph][pw] tile goes to
        its place in `self.out` ([H][W] float32, the GDT_Float32 band the reference creates)."""
        it = self.grid.device_items(items, self.device)
        pw, ph = self.grid.patch_size
        t = out_tiles.detach()
        if tuple(t.shape) != (it.numel(), 1, ph, pw) or t.dtype != torch.float32 or not t.is_cuda:
            raise ValueError(f"write_default: expected a CUDA float32 tensor of shape {(it.numel(), 1, ph, pw)}")
        t = t.contiguous()
        _, crop = self.grid.device_tables(self.device)
        H, W = self.out.shape
        _lib.call("fcd_tiles_scatter", t.data_ptr(), crop.data_ptr(), it.data_ptr(), it.numel(), pw, ph, self.out.data_ptr(),
                  H, W, _raw_stream())
        return self.out


class Evaluator:
    """`metrics.Evaluator` (metrics.py:6-82) with the confusion matrix accumulated on the device: no per-sample
    device->host copies inside the training loop (Demo_USSS.py:349-362).  Scores are computed on the host from the 2 x 2
    matrix when asked for (one 32-byte read)."""

    def __init__(self, num_class: int = 2, device="cuda:0"):
        if num_class != 2:
            raise ValueError("the change-detection loops use num_class = len(gt_map) = 2")
        self.num_class = num_class
        self.device = torch.device(device)
        self._counts = torch.zeros((4,), dtype=torch.int64, device=self.device)

    def reset(self) -> None:
        self._counts.zero_()

    def add_batch_map(self, ref_tiles: torch.Tensor, cmap: torch.Tensor, grid: TileGrid, items, prob_thresh: float = 0.5,
                      gt_map=(0, 1), pre_map=(0, 1)) -> None:
        """`cmask[cmap > prob_thresh] = 1` + centre crop + `add_batch_map` for every sample of the batch."""
        assert len(gt_map) == len(pre_map) == self.num_class                   # metrics.py:84-86
        it = grid.device_items(items, self.device)
        pw, ph = grid.patch_size
        if ref_tiles.shape != cmap.shape or tuple(cmap.shape) != (it.numel(), 1, ph, pw):
            raise ValueError("add_batch_map: ref and cmap must both be [B][1][patch_h][patch_w]")
        c, r = cmap.detach().float().contiguous(), ref_tiles.detach().float().contiguous()
        if not (c.is_cuda and r.is_cuda):
            raise RuntimeError("add_batch_map: tensors must live on the GPU (no CPU fallback)")
        _, crop = grid.device_tables(self.device)
        _lib.call("fcd_confusion_accumulate", c.data_ptr(), r.data_ptr(), crop.data_ptr(), it.data_ptr(), it.numel(), pw, ph,
                  float(prob_thresh), int(gt_map[0]), int(gt_map[1]), int(pre_map[0]), int(pre_map[1]),
                  self._counts.data_ptr(), _raw_stream())

    @property
    def confusion_matrix(self) -> np.ndarray:
        return self._counts.cpu().numpy().reshape(2, 2).astype(np.float64)

    # ---- scores: the reference's formulas, verbatim semantics (metrics.py:11-50) --------------------------------------
    def Pixel_Accuracy(self):
        cm = self.confusion_matrix
        return np.diag(cm).sum() / cm.sum()

    def Pixel_Kappa(self):
        cm = self.confusion_matrix
        po = np.diag(cm).sum() / cm.sum()
        pe = np.dot(cm.sum(axis=0), cm.sum(axis=1)) / np.square(cm.sum())
        return (po - pe) / (1 - pe)

    def Pixel_Accuracy_Class(self):
        cm = self.confusion_matrix
        acc = np.diag(cm) / cm.sum(axis=1)
        return np.nanmean(acc), acc

    def Pixel_Precision_Rate(self):
        cm = self.confusion_matrix
        return cm[1, 1] / (cm[0, 1] + cm[1, 1])

    def Pixel_Recall_Rate(self):
        cm = self.confusion_matrix
        return cm[1, 1] / (cm[1, 0] + cm[1, 1])

    def Pixel_F1_score(self):
        rec, pre = self.Pixel_Recall_Rate(), self.Pixel_Precision_Rate()
        return 2 * rec * pre / (rec + pre)

    def Mean_Intersection_over_Union(self):
        cm = self.confusion_matrix
        iou = np.diag(cm) / (cm.sum(axis=1) + cm.sum(axis=0) - np.diag(cm))
        return np.nanmean(iou), iou[1].copy()

    def Frequency_Weighted_Intersection_over_Union(self):
        cm = self.confusion_matrix
        freq = cm.sum(axis=1) / cm.sum()
        iu = np.diag(cm) / (cm.sum(axis=1) + cm.sum(axis=0) - np.diag(cm))
        return (freq[freq > 0] * iu[freq > 0]).sum()
