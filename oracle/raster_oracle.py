"""CPU restatement (numpy) of the reference's raster staging / tiled writer / accuracy accumulation — the data formats
either side of the hot path.  TEST INFRASTRUCTURE ONLY: imported by tests/, never by
the product path.  Pinned against outputs of the UNMODIFIED reference classes (GDALDataset, NORMALIZE, Dataset_meanstd,
Evaluator) run over an in-memory GDAL stand-in: oracle/make_golden_raster.py -> tests/golden/raster.npz, checked by
tests/test_raster_oracle.py.
"""
from __future__ import annotations

import math

import numpy as np


# ---- synthetic scenes shared by the golden generator and the tests (numpy PCG64 streams are stable across versions) ------
SCENES = {
    # tag: (xsize, ysize, dtype, bands, patch, overlap)      a = BASELINE config 1 shape; b = ragged edges, no-data border
    "a": (256, 256, np.float32, 4, (220, 220), (10, 10)),
    "b": (463, 431, np.uint16, 4, (220, 220), (10, 10)),
    "c": (150, 131, np.uint8, 3, (64, 48), (6, 4)),
}


def make_scene(tag: str):
    """-> dict(X, Y [C][ysize][xsize], REF [ysize][xsize] float32 in {0, 1, 2}, cmap [n][1][ph][pw] float32, patch, pad)."""
    xs, ys, dt, C, patch, pad = SCENES[tag]
    rng = np.random.default_rng({"a": 2024, "b": 2025, "c": 2026}[tag])
    hi = 250 if dt == np.uint8 else 2000
    X = rng.uniform(1, hi, (C, ys, xs))
    Y = X + rng.normal(0, hi * 0.03, (C, ys, xs))
    if tag != "a":
        X[:, :7, :] = 0
        X[:, :, -9:] = 0                                  # no-data border (all bands zero)
        Y = np.clip(Y, 0, np.iinfo(dt).max)
    X, Y = X.astype(dt), Y.astype(dt)
    REF = (rng.uniform(0, 1, (ys, xs)) > 0.8).astype(np.float32) + 1     # gt_map = [1, 2]
    REF[rng.uniform(0, 1, (ys, xs)) > 0.97] = 0                          # unlabeled pixels (ignored by the map)
    grid = tile_grid(xs, ys, patch, pad)
    n = len(grid["xstart"]) * len(grid["ystart"])
    cmap = rng.uniform(0, 1, (n, 1, patch[1], patch[0])).astype(np.float32)
    return {"X": X, "Y": Y, "REF": REF, "cmap": cmap, "patch": patch, "pad": pad, "grid": grid, "n": n}


def digest(a: np.ndarray) -> str:
    import hashlib
    a = np.ascontiguousarray(a)
    return hashlib.sha256(str(a.dtype).encode() + str(a.shape).encode() + a.tobytes()).hexdigest()


# ---- tile geometry: GDALDataset.__init__ (data_utils.py:57-63) and slice_assign (data_utils.py:151-176) -------------
def tile_grid(xsize: int, ysize: int, patch_size, overlap_padding):
    px, py = patch_size
    ox, oy = overlap_padding
    xstart = list(range(0, xsize, px - 2 * ox))
    xend = [x + px - 2 * ox for x in xstart if x + px - 2 * ox < xsize] + [xsize]
    ystart = list(range(0, ysize, py - 2 * oy))
    yend = [y + py - 2 * oy for y in ystart if y + py - 2 * oy < ysize] + [ysize]
    return {"xsize": xsize, "ysize": ysize, "patch": (px, py), "pad": (ox, oy), "xstart": xstart, "xend": xend,
            "ystart": ystart, "yend": yend}


def patch_count(grid):
    return len(grid["xstart"]), len(grid["ystart"])


def slice_assign(grid, item_x: int, item_y: int):
    """-> (slice, slice_read, slice_write), each (x, y, w, h); the `> 0` tests (not `>= 0`) are the reference's."""
    pad, xsize, ysize = grid["pad"], grid["xsize"], grid["ysize"]
    xstart, xend = grid["xstart"][item_x], grid["xend"][item_x]
    ystart, yend = grid["ystart"][item_y], grid["yend"][item_y]
    sl = (xstart, ystart, xend - xstart, yend - ystart)
    x_ori = 0 if xstart - pad[0] > 0 else pad[0]
    y_ori = 0 if ystart - pad[1] > 0 else pad[1]
    xstart = xstart - pad[0] if xstart - pad[0] > 0 else 0
    ystart = ystart - pad[1] if ystart - pad[1] > 0 else 0
    xend = xend + pad[0] if xend + pad[0] < xsize else xsize
    yend = yend + pad[1] if yend + pad[1] < ysize else ysize
    return sl, (xstart, ystart, xend - xstart, yend - ystart), (x_ori, y_ori, xend - xstart, yend - ystart)


def item_xy(grid, item: int):
    _, yc = patch_count(grid)
    return math.floor(item / yc), item % yc              # data_utils.py:92-95


# ---- N2: GDALDataset.__getitem__ (data_utils.py:94-123) + NORMALIZE.forward (CommonFunc.py:208-224) ------------------
def gather_tile(raster: np.ndarray, grid, item: int, mean=None, std=None) -> np.ndarray:
    """raster [C][H][W] -> float32 [C][patch_h][patch_w]; normalisation runs in float64 like the reference's numpy code."""
    ix, iy = item_xy(grid, item)
    _, rd, wr = slice_assign(grid, ix, iy)
    tmp = np.array(raster[:, rd[1]:rd[1] + rd[3], rd[0]:rd[0] + rd[2]], dtype=float)
    if mean is not None:
        for b in range(tmp.shape[0]):
            tmp[b] = (tmp[b] - mean[b]) / std[b]
    px, py = grid["patch"]
    out = np.zeros((raster.shape[0], py, px), dtype=float)
    out[:, wr[1]:wr[1] + wr[3], wr[0]:wr[0] + wr[2]] = tmp
    return out.astype(np.float32)


# ---- N2: Dataset_mean / Dataset_std (CommonFunc.py:436-499), float64 restatement -------------------------------------
def dataset_meanstd(raster_x: np.ndarray, raster_y: np.ndarray, grid):
    n = patch_count(grid)[0] * patch_count(grid)[1]
    tiles = [(gather_tile(raster_x, grid, i), gather_tile(raster_y, grid, i)) for i in range(n)]
    npix, mx, my = [], [], []
    for x, y in tiles:
        idx = x.sum(axis=0, dtype=np.float32) != 0
        npix.append(int(idx.sum()))
        mx.append(x[:, idx].astype(np.float64).mean(axis=1))
        my.append(y[:, idx].astype(np.float64).mean(axis=1))
    npix = np.array(npix, dtype=np.float64)
    total = npix.sum()
    mean_x = (np.array(mx) * (npix / total)[:, None]).sum(axis=0)
    mean_y = (np.array(my) * (npix / total)[:, None]).sum(axis=0)
    vx, vy = [], []
    for x, y in tiles:
        idx = x.sum(axis=0, dtype=np.float32) != 0
        vx.append(np.square(x[:, idx].astype(np.float64) - mean_x[:, None]).mean(axis=1))
        vy.append(np.square(y[:, idx].astype(np.float64) - mean_y[:, None]).mean(axis=1))
    std_x = np.sqrt((np.array(vx) * (npix / (total - 1))[:, None]).sum(axis=0))
    std_y = np.sqrt((np.array(vy) * (npix / (total - 1))[:, None]).sum(axis=0))
    return mean_x, std_x, mean_y, std_y


# ---- N3: GDALDataset.GDALwriteDefault (data_utils.py:178-213) ---------------------------------------------------------
def scatter_tile(out_raster: np.ndarray, tile: np.ndarray, grid, item: int) -> None:
    """tile [1][patch_h][patch_w] -> out_raster[H][W] (centre crop, in place)."""
    ix, iy = item_xy(grid, item)
    sl, _, _ = slice_assign(grid, ix, iy)
    pad = grid["pad"]
    out_raster[sl[1]:sl[1] + sl[3], sl[0]:sl[0] + sl[2]] = tile[0, pad[1]:pad[1] + sl[3], pad[0]:pad[0] + sl[2]]


# ---- N4: Demo_USSS.py:349-362 + Evaluator (metrics.py:6-82) -----------------------------------------------------------
def confusion_tile(ref_tile: np.ndarray, cmap_tile: np.ndarray, grid, item: int, prob_thresh: float, gt_map, pre_map):
    """ref_tile, cmap_tile [patch_h][patch_w] float32 -> int64 [2][2] counts over the tile's centre crop."""
    ix, iy = item_xy(grid, item)
    sl, _, _ = slice_assign(grid, ix, iy)
    pad = grid["pad"]
    cmask = np.zeros_like(cmap_tile)
    cmask[cmap_tile > np.float32(prob_thresh)] = 1
    gt = ref_tile[pad[1]:pad[1] + sl[3], pad[0]:pad[0] + sl[2]].astype(np.int16)
    pre = cmask[pad[1]:pad[1] + sl[3], pad[0]:pad[0] + sl[2]].astype(np.int16)
    cm = np.zeros((len(gt_map), len(pre_map)), dtype=np.int64)
    for i in range(len(gt_map)):
        for j in range(len(pre_map)):
            cm[i, j] = np.sum((gt == gt_map[i]) & (pre == pre_map[j]))
    return cm


def evaluator_scores(cm: np.ndarray) -> dict:
    """Evaluator's scores on a 2x2 confusion matrix (metrics.py:11-50)."""
    cm = cm.astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        acc = np.diag(cm).sum() / cm.sum()
        pe = np.dot(cm.sum(axis=0), cm.sum(axis=1)) / np.square(cm.sum())
        pre = cm[1, 1] / (cm[0, 1] + cm[1, 1])
        rec = cm[1, 1] / (cm[1, 0] + cm[1, 1])
        iou = np.diag(cm) / (cm.sum(axis=1) + cm.sum(axis=0) - np.diag(cm))
        return {"Pixel_Accuracy": acc, "Pixel_Kappa": (acc - pe) / (1 - pe), "Pixel_Precision_Rate": pre,
                "Pixel_Recall_Rate": rec, "Pixel_F1_score": 2 * rec * pre / (rec + pre),
                "Mean_Intersection_over_Union": (np.nanmean(iou), iou[1])}
