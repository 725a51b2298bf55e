"""GPU parity of the raster staging / tiled writer / confusion-matrix kernels through
fcdgan_b200.raster (C ABI: fcd_tiles_gather / fcd_tiles_moments / fcd_tiles_scatter / fcd_confusion_accumulate) against the
numpy oracle and the reference-generated fixture (tests/golden/raster.npz).  Byte / integer / index work is compared BIT FOR
BIT (the normalisation runs in float64 and rounds once to float32 exactly like the reference's numpy code); the dataset
statistics within 1e-5 relative (the reference accumulates them in float32)."""
import os

import numpy as np
import pytest
import torch

from fcdgan_b200 import raster as R
from oracle import raster_oracle as RO

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "raster.npz"))


def _pair(tag, overlap=None):
    sc = RO.make_scene(tag)
    grid = R.TileGrid(sc["X"].shape[2], sc["X"].shape[1], sc["patch"], sc["pad"] if overlap is None else overlap)
    return sc, grid, R.RasterPair(sc["X"], sc["Y"], grid, ref=sc["REF"], device=DEV)


@pytest.mark.parametrize("tag", list(RO.SCENES))
def test_tiles_bit_exact(tag):
    sc, grid, pair = _pair(tag)
    assert list(grid.patch_count()) == GOLD[f"{tag}_counts"].tolist() and len(pair) == sc["n"]
    stats = [row.tolist() for row in GOLD[f"{tag}_meanstd"]]
    items = list(range(sc["n"]))
    xt, yt, rt = pair.tiles(torch.tensor(items), stats)
    assert RO.digest(xt.cpu().numpy()) == str(GOLD[f"{tag}_xt_sha"])
    assert RO.digest(yt.cpu().numpy()) == str(GOLD[f"{tag}_yt_sha"])
    assert RO.digest(rt.cpu().numpy()) == str(GOLD[f"{tag}_rt_sha"])
    # a shuffled, repeated batch (DataLoader(shuffle=True)) against the oracle, un-normalised
    sel = [items[-1], items[0], items[len(items) // 2], items[0]]
    xt2, _, _ = pair.tiles(sel)
    want = np.stack([RO.gather_tile(sc["X"], sc["grid"], i) for i in sel])
    assert np.array_equal(xt2.cpu().numpy(), want)


@pytest.mark.parametrize("tag", list(RO.SCENES))
def test_meanstd(tag):
    sc, grid, pair = _pair(tag, overlap=(0, 0))                   # Demo_USSS.py:88-89
    got = np.array(pair.meanstd(batch=5))
    np.testing.assert_allclose(got, GOLD[f"{tag}_meanstd"], rtol=1e-5)


@pytest.mark.parametrize("tag", list(RO.SCENES))
def test_write_default_and_confusion(tag):
    sc, grid, pair = _pair(tag)
    n = sc["n"]
    cmap = torch.from_numpy(sc["cmap"]).to(DEV)
    acc = R.Evaluator(2, device=DEV)
    for s in range(0, n, 5):                                       # ragged last batch
        items = list(range(s, min(n, s + 5)))
        _, _, rt = pair.tiles(items)
        pair.write_default(cmap[s:s + len(items)], items)
        acc.add_batch_map(rt, cmap[s:s + len(items)], grid, items, 0.5, [1, 2], [0, 1])
    assert RO.digest(pair.out.cpu().numpy()) == str(GOLD[f"{tag}_stitched_sha"])
    assert np.array_equal(acc.confusion_matrix.astype(np.int64), GOLD[f"{tag}_confusion"])
    miou, ciou = acc.Mean_Intersection_over_Union()
    got = [acc.Pixel_Accuracy(), acc.Pixel_Kappa(), acc.Pixel_Precision_Rate(), acc.Pixel_Recall_Rate(), acc.Pixel_F1_score(),
           miou, ciou]
    np.testing.assert_allclose(got, GOLD[f"{tag}_scores"], rtol=1e-12)
    acc.reset()
    assert acc.confusion_matrix.sum() == 0


def test_full_size_roundtrip_properties():
    """Size-independent properties at a production-sized scene (8192 x 6000 x 13 uint16, patch 256, overlap 16):
    gather -> scatter of band 0 reproduces the raster (every pixel is written exactly once by a centre crop), and the
    confusion counts of a map against itself land on the diagonal and sum to the number of labelled pixels."""
    g = torch.Generator(device=DEV).manual_seed(5)
    H, W, C = 6000, 8192, 13
    x = torch.randint(1, 4000, (C, H, W), generator=g, dtype=torch.int32, device=DEV).to(torch.uint16)
    grid = R.TileGrid(W, H, (256, 256), (16, 16))
    ref = (torch.rand(1, H, W, generator=g, device=DEV) > 0.7).float()
    pair = R.RasterPair(x, x, grid, ref=ref, device=DEV)
    acc = R.Evaluator(2, device=DEV)
    n = len(grid)
    for s in range(0, n, 256):
        items = list(range(s, min(n, s + 256)))
        xt, _, rt = pair.tiles(items)
        pair.write_default(xt[:, :1].contiguous(), items)
        acc.add_batch_map(rt, rt, grid, items, 0.5, [0, 1], [0, 1])
    assert torch.equal(pair.out, x[0].float())
    cm = acc.confusion_matrix
    assert cm[0, 1] == 0 and cm[1, 0] == 0 and cm.sum() == H * W and cm[1, 1] == ref.sum().item()


def test_errors():
    sc, grid, pair = _pair("c")
    with pytest.raises(ValueError):
        pair.tiles([0], ([0.0], [1.0], [0.0], [1.0]))               # fewer statistics than bands (CommonFunc.py:211-213)
    with pytest.raises(ValueError):
        R.RasterPair(sc["X"], sc["Y"][:, :-1], grid, device=DEV)
    with pytest.raises(RuntimeError):
        R.RasterPair(sc["X"], sc["Y"], grid, device="cpu")
    with pytest.raises(ValueError):
        pair.write_default(torch.zeros(1, 1, 3, 3, device=DEV), [0])
