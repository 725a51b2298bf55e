"""North-star target configuration (BASELINE.json): Generator forward + backward (+ Adam step) at batch 32 x (13 x 256 x 256),
one GPU, whole iteration replayed from a CUDA graph.  Prints one JSON line per precision: tiles/s and the fraction of the
measured sustained bf16 tensor peak at 203.6 algorithmic GFLOP per tile."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import fcdgan_b200 as fb
from fcdgan_b200.graph import GraphedStep

dev = torch.device("cuda:0")
B, C = int(os.environ.get("B", 32)), 13
GF = 203.6
for prec in sys.argv[1:] or ["parity", "fast"]:
    fb.set_precision(prec)
    torch.manual_seed(0)
    netG = fb.Generator(C).to(dev).train()
    opt = torch.optim.Adam(netG.parameters(), lr=2e-4, betas=(0.9, 0.99), capturable=True)
    x, y, _, _ = bench.synth(16, 3, device=dev)
    x, y = x.repeat(B // 16, 1, 1, 1), y.repeat(B // 16, 1, 1, 1)
    zero = torch.zeros(B, 1, 256, 256, device=dev)

    def step(x, y):
        y_fake = netG(x)
        loss, _, _, _ = fb.losses._MaskedRecon.apply(y, y_fake, zero, fb.losses.LOSS_L1, False)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss

    g = GraphedStep(step, [x, y], warmup=3)
    for _ in range(3):
        g()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n):
        loss = g()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    pk = bench.peaks()
    tf = GF * 1e9 * B / (ms * 1e-3) / 1e12
    print(json.dumps({"workload": f"Generator fwd+bwd+Adam, batch {B} x 13x256x256", "precision": prec, "ms_per_iter": round(ms, 3),
                      "tiles_per_s": round(B / ms * 1e3, 1), "algorithmic_tflops": round(tf, 1),
                      "frac_of_sustained_bf16_peak": round(tf / pk["tflops_sustained"], 4), "peak": pk["tflops_sustained"],
                      "loss": round(float(loss), 5), "peak_mem_GiB": round(torch.cuda.max_memory_allocated() / 2**30, 1)}), flush=True)
    del g, netG, opt
    torch.cuda.empty_cache()
