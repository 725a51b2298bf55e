"""Generate tests/golden/raster.npz by running the UNMODIFIED reference data path (GDALDataset, NORMALIZE,
Dataset_meanstd, GDALwriteDefault, Evaluator) over the numpy-backed GDAL stand-in (oracle/_shim/osgeo).
Run in the build container only:  python oracle/make_golden_raster.py

Scenes: oracle/raster_oracle.py SCENES (a = BASELINE config 1 shape: 256 x 256 x 4, patch 220, overlap 10 -> 2 x 2 tiles;
b = 463 x 431 uint16 with a no-data border -> 3 x 3 ragged tiles; c = small uint8 scene with a non-square patch).
The inputs are regenerated from their seeds by the tests; the fixture stores SHA-256 digests of the large reference
outputs (tile batches, stitched raster) and the small ones in full (statistics, confusion matrix, scores).
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_import  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "raster.npz")


def main():
    ref_import.load()
    from osgeo import gdal
    import CommonFunc
    import data_utils
    import metrics

    from oracle import raster_oracle as RO

    res = {}
    for tag in RO.SCENES:
        sc = RO.make_scene(tag)
        X, Y, REF, cmap, patch, pad = sc["X"], sc["Y"], sc["REF"], sc["cmap"], sc["patch"], sc["pad"]
        gdal.register(f"{tag}_x.tif", X); gdal.register(f"{tag}_y.tif", Y); gdal.register(f"{tag}_ref.tif", REF)
        # Demo_USSS.py:88-95: statistics over un-padded tiles, then the normalising dataset
        ds0 = data_utils.GDALDataset(f"{tag}_x.tif", f"{tag}_y.tif", outPath=f"{tag}_o.tif", patch_size=patch,
                                     overlap_padding=(0, 0))
        with tempfile.TemporaryDirectory() as td:
            mX, sX, mY, sY = CommonFunc.Dataset_meanstd(os.path.join(td, "1.txt"), os.path.join(td, "2.txt"), ds0)
        scaler = CommonFunc.NORMALIZE(mX, sX, mY, sY)
        ds = data_utils.GDALDataset(f"{tag}_x.tif", f"{tag}_y.tif", refPath=f"{tag}_ref.tif", outPath=f"{tag}_o.tif",
                                    enhance=scaler, patch_size=patch, overlap_padding=pad)
        n = len(ds)
        assert n == sc["n"]
        xt, yt, rt = [], [], []
        for i in range(n):
            x, y, item, ref = ds[i]
            assert int(item) == i
            xt.append(x.numpy()); yt.append(y.numpy()); rt.append(ref.numpy())
        xt, yt, rt = np.stack(xt), np.stack(yt), np.stack(rt)
        gt_map, pre_map, thresh = [1, 2], [0, 1], 0.5            # Demo_USSS.py:64-67
        acc = metrics.Evaluator(num_class=2)
        xc, yc = ds.patch_count()
        for i in range(n):
            ds.GDALwriteDefault(cmap[i], i)
            cmask = np.zeros_like(cmap[i]); cmask[cmap[i] > thresh] = 1
            sl, _, _ = ds.slice_assign(i // yc, i % yc)
            acc.add_batch_map(rt[i][0, pad[1]:pad[1] + sl[3], pad[0]:pad[0] + sl[2]].astype(np.int16),
                              cmask[0, pad[1]:pad[1] + sl[3], pad[0]:pad[0] + sl[2]].astype(np.int16), gt_map, pre_map)
        stitched = gdal.store[f"{tag}_o.tif"].array[0].copy()
        miou, ciou = acc.Mean_Intersection_over_Union()
        res.update({f"{tag}_meanstd": np.array([mX, sX, mY, sY], dtype=np.float64), f"{tag}_counts": np.array([xc, yc]),
                    f"{tag}_xt_sha": RO.digest(xt), f"{tag}_yt_sha": RO.digest(yt), f"{tag}_rt_sha": RO.digest(rt),
                    f"{tag}_stitched_sha": RO.digest(stitched), f"{tag}_xt_head": xt[:, :, :3, :5].copy(),
                    f"{tag}_confusion": acc.confusion_matrix.astype(np.int64),
                    f"{tag}_scores": np.array([acc.Pixel_Accuracy(), acc.Pixel_Kappa(), acc.Pixel_Precision_Rate(),
                                               acc.Pixel_Recall_Rate(), acc.Pixel_F1_score(), miou, ciou])})
        print(tag, "tiles", n, (xc, yc), "mean", np.round(mX, 3), "confusion", acc.confusion_matrix.ravel())
    np.savez_compressed(OUT, **res)
    print("wrote", OUT, os.path.getsize(OUT) // 1024, "KiB")


if __name__ == "__main__":
    main()
