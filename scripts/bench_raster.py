"""Measures the raster-side kernels against the HBM roofline on a production-sized scene:
13-band uint16 rasters, 256 x 256 patches with 16 px overlap, batches of 256 tiles.  Algorithmic bytes per launch:
  gather   : read B*C*ph*pw source elements (2 B for uint16) + write B*C*ph*pw fp32
  scatter  : read + write the centre crops (2 * 4 B per cropped pixel)
  confusion: read cmap + ref centre crops (2 * 4 B per cropped pixel)
Prints one JSON line per kernel; CPU column = the numpy oracle (the reference's own algorithm) on one tile batch slice."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from fcdgan_b200 import raster as R
from oracle import raster_oracle as RO

dev = "cuda:0"
H, W, C, B = 6000, 8192, 13, 256
g = torch.Generator(device=dev).manual_seed(1)
x = torch.randint(1, 4000, (C, H, W), generator=g, dtype=torch.int32, device=dev).to(torch.uint16)
y = torch.randint(1, 4000, (C, H, W), generator=g, dtype=torch.int32, device=dev).to(torch.uint16)
ref = (torch.rand(1, H, W, generator=g, device=dev) > 0.7).float()
grid = R.TileGrid(W, H, (256, 256), (16, 16))
pair = R.RasterPair(x, y, grid, ref=ref, device=dev)
stats = ([1000.0 + i for i in range(C)], [500.0 + i for i in range(C)]) * 2
items = list(range(100, 100 + B))
acc = R.Evaluator(2, device=dev)
peak = bench.peaks()["hbm_gbs"]
crop = int(sum(g_[4] * g_[5] for g_ in grid.crop_geom()[items]))
read_px = int(sum(g_[2] * g_[3] for g_ in grid.gather_geom()[items]))
items_dev = grid.device_items(items, dev)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


xt, yt, rt = pair.tiles(items, stats)
cm = torch.rand(B, 1, 256, 256, device=dev)
rows = []
ms = timeit(lambda: pair._gather(pair.x, items_dev, stats[0], stats[1]))
rows.append(("fcd_tiles_gather (uint16 -> normalised fp32 tiles)", ms, read_px * C * 2 + B * C * 256 * 256 * 4))
ms = timeit(lambda: pair.write_default(cm, items_dev))
rows.append(("fcd_tiles_scatter", ms, crop * 8))
ms = timeit(lambda: acc.add_batch_map(rt, cm, grid, items_dev, 0.5, [0, 1], [0, 1]))
rows.append(("fcd_confusion_accumulate", ms, crop * 8))
t0 = time.perf_counter()
xs = x.cpu().numpy()
sub = items[:8]
t0 = time.perf_counter()
for i in sub:
    RO.gather_tile(xs, RO.tile_grid(W, H, (256, 256), (16, 16)), i, stats[0], stats[1])
cpu_ms_per_tile = (time.perf_counter() - t0) * 1e3 / len(sub)
for name, ms, nbytes in rows:
    print(json.dumps({"kernel": name, "ms_per_launch": round(ms, 4), "tiles_per_launch": B, "algorithmic_MB": round(nbytes / 1e6, 1),
                      "achieved_GBs": round(nbytes / ms / 1e6, 1), "hbm_peak_GBs": peak, "frac": round(nbytes / ms / 1e6 / peak, 3),
                      "tiles_per_s": round(B / ms * 1e3)}))
print(json.dumps({"cpu_oracle_gather_ms_per_tile": round(cpu_ms_per_tile, 2), "cpu_tiles_per_s": round(1e3 / cpu_ms_per_tile, 1),
                  "note": "numpy restatement of GDALDataset.__getitem__ + NORMALIZE on one host core, 13 x 256 x 256 uint16"}))
