"""The numpy raster oracle (oracle/raster_oracle.py) against outputs of the UNMODIFIED reference data path
(GDALDataset, NORMALIZE, Dataset_meanstd, GDALwriteDefault, Evaluator; oracle/make_golden_raster.py -> tests/golden/raster.npz).
Tile batches and the stitched raster must match BIT FOR BIT (SHA-256 of the arrays), the confusion matrix exactly;
statistics within 1e-5 relative (the reference accumulates them in float32)."""
import os

import numpy as np
import pytest

from oracle import raster_oracle as RO

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "raster.npz"))


@pytest.mark.parametrize("tag", list(RO.SCENES))
def test_geometry_and_tiles_bit_exact(tag):
    sc = RO.make_scene(tag)
    grid = sc["grid"]
    assert list(RO.patch_count(grid)) == GOLD[f"{tag}_counts"].tolist()
    mX, sX, mY, sY = GOLD[f"{tag}_meanstd"]
    xt = np.stack([RO.gather_tile(sc["X"], grid, i, mX, sX) for i in range(sc["n"])])
    yt = np.stack([RO.gather_tile(sc["Y"], grid, i, mY, sY) for i in range(sc["n"])])
    rt = np.stack([RO.gather_tile(sc["REF"][None], grid, i) for i in range(sc["n"])])
    assert np.array_equal(xt[:, :, :3, :5], GOLD[f"{tag}_xt_head"])
    assert RO.digest(xt) == str(GOLD[f"{tag}_xt_sha"])
    assert RO.digest(yt) == str(GOLD[f"{tag}_yt_sha"])
    assert RO.digest(rt) == str(GOLD[f"{tag}_rt_sha"])


@pytest.mark.parametrize("tag", list(RO.SCENES))
def test_meanstd(tag):
    sc = RO.make_scene(tag)
    grid0 = RO.tile_grid(sc["X"].shape[2], sc["X"].shape[1], sc["patch"], (0, 0))      # Demo_USSS.py:88-89
    got = np.array(RO.dataset_meanstd(sc["X"], sc["Y"], grid0))
    np.testing.assert_allclose(got, GOLD[f"{tag}_meanstd"], rtol=1e-5)


@pytest.mark.parametrize("tag", list(RO.SCENES))
def test_stitch_and_confusion(tag):
    sc = RO.make_scene(tag)
    grid = sc["grid"]
    out = np.zeros((sc["X"].shape[1], sc["X"].shape[2]), dtype=np.float32)
    cm = np.zeros((2, 2), dtype=np.int64)
    for i in range(sc["n"]):
        RO.scatter_tile(out, sc["cmap"][i], grid, i)
        ref_tile = RO.gather_tile(sc["REF"][None], grid, i)[0]
        cm += RO.confusion_tile(ref_tile, sc["cmap"][i, 0], grid, i, 0.5, [1, 2], [0, 1])
    assert RO.digest(out) == str(GOLD[f"{tag}_stitched_sha"])
    assert np.array_equal(cm, GOLD[f"{tag}_confusion"])
    s = RO.evaluator_scores(cm)
    got = [s["Pixel_Accuracy"], s["Pixel_Kappa"], s["Pixel_Precision_Rate"], s["Pixel_Recall_Rate"], s["Pixel_F1_score"],
           *s["Mean_Intersection_over_Union"]]
    np.testing.assert_allclose(got, GOLD[f"{tag}_scores"], rtol=1e-12)


def test_slice_assign_edge_quirk():
    """slice_assign uses `> 0` (data_utils.py:160-164): a tile whose start equals the padding reads from 0 but is written at
    offset pad — kept as the reference does it."""
    grid = RO.tile_grid(100, 100, (40, 40), (10, 10))           # starts 0, 20, 40, ...: the second tile has xstart - pad = 10 > 0
    sl, rd, wr = RO.slice_assign(grid, 1, 0)
    assert sl == (20, 0, 20, 20) and rd == (10, 0, 40, 30) and wr == (0, 10, 40, 30)
    grid = RO.tile_grid(100, 100, (30, 30), (10, 10))           # starts 0, 10, 20: xstart - pad == 0 for the second tile
    sl, rd, wr = RO.slice_assign(grid, 1, 0)
    assert rd[0] == 0 and wr[0] == 10
