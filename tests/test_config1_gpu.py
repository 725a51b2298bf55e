"""BASELINE.json configs[0] end to end on the GPU path: the 3-stage Demo_USSS flow (Demo_USSS.py:88-122 set-up, 124-189
generator warm-up, 192-286 segmentor warm-up, 289-400 joint training, 404-473 inference + raster write + accuracy) on a
synthetic 256 x 256 x 4 bi-temporal raster pair -> 4 tiles of 220 x 220 (patch 220, overlap 10), one epoch per stage with the
demo's batch size 10 (= one iteration per stage), through `fcdgan_b200.raster` + `fcdgan_b200.steps` — against the SAME flow
run by the CPU oracles (oracle/raster_oracle.py for the rasters, oracle/fcd_oracle.py + torch.optim.Adam for the networks).

Checks: tile batches bit-exact; every stage's losses within 1e-3; the change-density raster stitched from the centre crops
within 2e-3 absolute of the oracle's after three optimizer steps (Adam's sign-like first steps amplify 1e-7 gradient noise into
lr-sized parameter differences; the map itself is in [0, 1]); the confusion matrix bit-exact against the numpy Evaluator
restatement applied to the same stitched map."""
import numpy as np
import pytest
import torch

import fcdgan_b200 as fb
from fcdgan_b200 import raster as R
from oracle import fcd_oracle as O
from oracle import raster_oracle as RO

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_usss_three_stage_flow_on_a_256x256x4_raster():
    fb.set_precision("parity")
    sc = RO.make_scene("a")                                   # 256 x 256 x 4 float32, patch 220, overlap 10 -> 4 tiles
    C, n = 4, sc["n"]
    assert n == 4
    grid = R.TileGrid(256, 256, sc["patch"], sc["pad"])
    grid0 = R.TileGrid(256, 256, sc["patch"], (0, 0))         # statistics grid, Demo_USSS.py:88-89
    stats = R.RasterPair(sc["X"], sc["Y"], grid0, device=DEV).meanstd()
    stats_o = RO.dataset_meanstd(sc["X"], sc["Y"], RO.tile_grid(256, 256, sc["patch"], (0, 0)))
    np.testing.assert_allclose(np.array(stats), np.array(stats_o), rtol=1e-5)
    pair = R.RasterPair(sc["X"], sc["Y"], grid, ref=sc["REF"], device=DEV)
    items = [2, 0, 3, 1]                                      # one shuffled batch of all four tiles (batch_size 10)
    x, y, ref = pair.tiles(items, stats)
    xo = torch.from_numpy(np.stack([RO.gather_tile(sc["X"], sc["grid"], i, stats[0], stats[1]) for i in items]))
    yo = torch.from_numpy(np.stack([RO.gather_tile(sc["Y"], sc["grid"], i, stats[2], stats[3]) for i in items]))
    assert torch.equal(x.cpu(), xo) and torch.equal(y.cpu(), yo)

    sdG, sdS = O.make_state_dict(O.generator_spec(C), 11), O.make_state_dict(O.segmentor_spec(C, 1, True), 12)
    netG, netS = fb.Generator(C), fb.Segmentor(C, 1, True)                   # Demo_USSS.py:109-111
    netG.load_state_dict(sdG); netS.load_state_dict(sdS)
    netG.to(DEV).train(); netS.to(DEV).train()
    crit = fb.CNetLoss(channel=C, perception_layer=1, perception_perBand=True)
    optG = torch.optim.Adam(netG.parameters(), lr=2e-4, betas=(0.9, 0.99))   # Demo_USSS.py:121-122
    optS = torch.optim.Adam(netS.parameters(), lr=2e-4, betas=(0.9, 0.99))
    ssim_w, l1_w = 0.3, 0.65
    s1 = fb.usss_g_step(netG, x, y, crit, optG, ssim_weight=ssim_w)
    s2 = fb.usss_s_step(netG, netS, x, y, crit, optS, ssim_weight=ssim_w, l1_weight=l1_w)
    s3 = fb.usss_step(netG, netS, x, y, crit, optG, optS, ssim_weight=ssim_w, l1_weight=l1_w)
    netS.eval()                                                               # Demo_USSS.py:404-435
    with torch.no_grad():
        cmap = netS(x, y)
    pair.write_default(cmap, items)
    acc = R.Evaluator(2, device=DEV)
    acc.add_batch_map(ref, cmap, grid, items, 0.5, [1, 2], [0, 1])

    # ---- the same flow on the CPU oracle -------------------------------------------------------------------------------
    oG, oS = O.clone_sd(sdG, requires_grad=True), O.clone_sd(sdS, requires_grad=True)
    pG = [v for v in oG.values() if v.requires_grad]
    pS = [v for v in oS.values() if v.requires_grad]
    aG = torch.optim.Adam(pG, lr=2e-4, betas=(0.9, 0.99))
    aS = torch.optim.Adam(pS, lr=2e-4, betas=(0.9, 0.99))
    zero = torch.zeros(n, 1, 220, 220)
    aG.zero_grad()
    gl, l1, sl = O.cnet_loss(yo, O.generator(oG, xo, train=True), zero)
    (gl + ssim_w * sl).backward()
    aG.step()
    t1 = (gl.item(), sl.item())
    gl, l1, sl = O.cnet_loss(yo, O.generator(oG, xo, train=True), O.segmentor(oS, xo, yo, True, True))
    aS.zero_grad()
    (gl + l1_w * l1 + ssim_w * sl).backward()
    aS.step()
    t2 = (gl.item(), l1.item(), sl.item())
    gl, l1, sl = O.cnet_loss(yo, O.generator(oG, xo, train=True), O.segmentor(oS, xo, yo, True, True))
    loss = gl + ssim_w * sl
    aG.zero_grad()
    loss.backward(retain_graph=True)
    aS.zero_grad()
    (loss + l1_w * l1).backward()
    aG.step(); aS.step()
    t3 = (gl.item(), l1.item(), sl.item())
    with torch.no_grad():
        cmap_o = O.segmentor(oS, xo, yo, True, False)

    def close(a, b):
        return abs(float(a) - b) <= 1e-3 * max(abs(b), 1e-3)

    assert close(s1["generator_loss"], t1[0]) and close(s1["ssim_loss"], t1[1]), (s1, t1)
    assert close(s2["generator_loss"], t2[0]) and close(s2["l1_loss"], t2[1]) and close(s2["ssim_loss"], t2[2]), (s2, t2)
    assert close(s3["generator_loss"], t3[0]) and close(s3["l1_loss"], t3[1]) and close(s3["ssim_loss"], t3[2]), (s3, t3)
    diff = (cmap.cpu() - cmap_o).abs().max().item()
    print(f"[config 1] change-density map after 3 optimizer steps: max |diff| {diff:.2e}")
    assert diff < 2e-3
    # stitched raster and confusion matrix: bit-exact against the numpy restatement applied to OUR tiles
    out_o = np.zeros((256, 256), np.float32)
    cm_o = np.zeros((2, 2), np.int64)
    cm_np = cmap.cpu().numpy()
    ref_np = ref.cpu().numpy()
    for k, it in enumerate(items):
        RO.scatter_tile(out_o, cm_np[k], sc["grid"], it)
        cm_o += RO.confusion_tile(ref_np[k, 0], cm_np[k, 0], sc["grid"], it, 0.5, [1, 2], [0, 1])
    assert np.array_equal(pair.out.cpu().numpy(), out_o)
    assert np.array_equal(acc.confusion_matrix.astype(np.int64), cm_o)
    assert 0.0 <= acc.Pixel_Accuracy() <= 1.0
