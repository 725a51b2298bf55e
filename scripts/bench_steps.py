"""Times the USSS joint iteration (Demo_USSS.py:305-341) and the RSSS adversarial iteration (Demo_RSSS.py:270-332) on
synthetic 13-band tiles and dumps the per-call CUDA-event tables (BASELINE configs 3/4 at single-GPU batch sizes).

    python scripts/bench_steps.py [--batch 8] [--size 256] [--steps 5] [--precision parity] [--out gpurun_out]
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn
import bench
import fcdgan_b200 as fb
from fcdgan_b200 import engine as E

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=8)
ap.add_argument("--size", type=int, default=256)
ap.add_argument("--steps", type=int, default=5)
ap.add_argument("--precision", default="parity")
ap.add_argument("--out", default="gpurun_out")
ap.add_argument("--which", default="usss,rsss")
ap.add_argument("--graph", action="store_true", help="replay each step from one CUDA graph (fcdgan_b200.graph.GraphedStep)")
args = ap.parse_args()
dev = torch.device("cuda:0")
fb.set_precision(args.precision)
B, C = args.batch, 13
bench.H = bench.W = args.size
x, y, region, _ = bench.synth(B, 7, device=dev)
torch.manual_seed(0)
netG, netS, netD = fb.Generator(C).to(dev), fb.Segmentor(C, 1, True).to(dev), fb.Discriminator_SRGAN_simple(C).to(dev)
GF = {"usss": 203.6 + 786.7, "rsss": 889.0}   # algorithmic GFLOP per pair at 13x256^2; scaled by (size/256)^2
scale = (args.size / 256.0) ** 2


def usss():
    netG.train(); netS.train()
    optG = torch.optim.Adam(netG.parameters(), lr=2e-4, betas=(0.9, 0.99), capturable=args.graph)
    optS = torch.optim.Adam(netS.parameters(), lr=2e-4, betas=(0.9, 0.99), capturable=args.graph)
    crit = fb.CNetLoss(channel=C)

    def step():
        y_fake = netG(x)
        cmap = netS(x, y)
        gl, l1, perc, sl = crit(y, y_fake, cmap)
        Loss = gl + 0.3 * sl
        optG.zero_grad(set_to_none=True)
        Loss.backward(retain_graph=True)
        NetLoss = gl + 0.65 * l1 + 0.3 * sl
        optS.zero_grad(set_to_none=True)
        NetLoss.backward()
        optG.step(); optS.step()
        return NetLoss
    return step


def rsss():
    netG.eval(); netS.train(); netD.train()
    optS = torch.optim.RMSprop(netS.parameters(), lr=5e-5, capturable=args.graph)
    optD = torch.optim.RMSprop(netD.parameters(), lr=5e-5, capturable=args.graph)
    gcrit = fb.CGeneratorLoss(channel=C)

    def step():
        cmap = netS(x, y)
        x_mask, y_mask = fb.soft_mask(x, cmap), fb.soft_mask(y, cmap)
        c_out = netD(x_mask, y_mask)
        y_unc = fb.soft_mask(y, cmap, other=x, region=region)
        nc_out = netD(x_mask, y_unc)
        optD.zero_grad(set_to_none=True)
        d_loss = 1 + fb.mean(nc_out) - fb.mean(c_out)
        d_loss.backward(retain_graph=True)
        optD.step()
        c_out2 = netD(x_mask, y_mask)
        y_fake = netG(x)
        gl, sl, _ = gcrit(y, y_fake, cmap)
        s_loss = fb.mean(c_out2) + 0.02 * fb.region_loss(cmap, region, nn.L1Loss()) + 0.5 * (gl + 0.0 * sl) \
            + 2.0 * fb.region_loss(cmap, 1 - region, nn.MSELoss())
        optS.zero_grad(set_to_none=True)
        s_loss.backward()
        optS.step()
        return s_loss
    return step


for name in args.which.split(","):
    step = eager = {"usss": usss, "rsss": rsss}[name]()
    launch = "eager"
    if args.graph:
        try:
            from fcdgan_b200.graph import GraphedStep
            g = GraphedStep(lambda: eager(), [], warmup=3)
            step = lambda: g()
            launch = "one CUDA graph"
        except Exception as e:
            launch = f"eager (capture failed: {type(e).__name__}: {str(e)[:100]})"
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    E.PROFILE = []
    eager(); torch.cuda.synchronize()
    agg = {}
    for n, tag, fl, nb, a, b in E.PROFILE:
        d = agg.setdefault(tag, [0.0, 0, 0.0]); d[0] += a.elapsed_time(b); d[1] += 1; d[2] += fl
    E.PROFILE = None
    tot = sum(v[0] for v in agg.values())
    res = {"step": name, "batch": B, "size": args.size, "precision": args.precision, "launch": launch, "ms_per_step": round(ms, 2),
           "tile_pairs_per_s": round(B / ms * 1e3, 2), "algorithmic_tflops": round(GF[name] * scale * B / ms, 1),
           "kernel_ms_sum": round(tot, 2), "loss": float(loss), "peak_mem_GiB": round(torch.cuda.max_memory_allocated() / 2**30, 1)}
    print(json.dumps(res), flush=True)
    with open(os.path.join(args.out, f"table_{name}_{args.precision}_b{B}_{args.size}.txt"), "w") as fh:
        fh.write("# " + json.dumps(res) + "\n# tag | ms/step | launches | share | TFLOP/s (algorithmic)\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            tf = v[2] / (v[0] * 1e-3) / 1e12 if v[0] > 0 and v[2] > 0 else 0.0
            fh.write(f"{k:46s} {v[0]:9.3f} {v[1]:5d} {v[0] / tot:7.3f} {tf:8.1f}\n")
