#!/usr/bin/env python
"""bench.py — benchmarks of the FCD-GAN hot path (BASELINE.json metric and configurations).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config 2|g32|3|4|5] [--precision parity|fast]
                    [--impl ours|reference|cudnn] [--batch B]

`--config` (default 2 = BASELINE.json configs[1], the configuration the metric is quoted on):
    2    Generator + Discriminator forward/backward (+ Adam / RMSprop steps), 16 tile pairs of 13x256x256 per GPU, 227.3 GF/pair
             y_fake = G(x); masked L1(y, y_fake, cmap = 0) -> backward                                    (Demo_USSS.py:142-159)
             c_out = D(x*(1-cmap), y*(1-cmap)); nc_out = D(x*(1-cmap), (y*(1-r)+x*r)*(1-cmap)); d_loss     (Demo_RSSS.py:285-307)
    g32  north-star configuration: Generator forward + backward + Adam, 32 tiles of 13x256x256, 203.6 GF/tile
    3    USSS joint iteration G + S + CNetLoss with a live MS-SSIM gradient, 32 pairs of 13x512x512 (Demo_USSS.py:305-341), 3961 GF/pair
    4    RSSS adversarial iteration G + S + D, 16 pairs of 13x256x256 per GPU (Demo_RSSS.py:270-332), 889 GF/pair
    5    WSSS adversarial iteration, 32 (changed, unchanged) items of 3x256x256 per GPU (Demo_WSSS.py:240-323), 1654 GF/item
Configs 3 / 4 / 5 run the step bodies of fcdgan_b200.steps in their `lean` form by default (`--form faithful` replays the reference's
call sequence including the backward sweeps whose results it discards; same losses, same gradients at every optimizer step).
One "step" = one whole iteration, optimizer steps included.  N > 1 (torchrun): every rank runs its own batch (weak scaling)
and the gradients of every trained network are all-reduced over NCCL (fcdgan_b200.parallel.GradSync) at the step body's
exchange points (fcdgan_b200.steps / graph.YieldingStep).

Prints ONE JSON line (rank 0).  `value` is device-timed with inputs resident in HBM; `e2e` repeats the measurement through the
same public API with the inputs in pinned host memory (H2D inside the timed region, losses read back).  `roofline` is the
dominant kernel of the step (by summed CUDA-event time over an instrumented pass): algorithmic FLOPs per launch / mean launch
time against MEASURED_PEAKS.json's sustained bf16 peak.  `gpu_baseline` (N = 1) is the UNMODIFIED reference's classes moved to
the same B200 (`.cuda()`: torch + cuDNN — what the reference runs on a GPU), same step, same batch, with PyTorch's default TF32
convolutions and with TF32 off.  `cpu_baseline` / `--impl reference` time the UNMODIFIED reference classes (staged copy
oracle/_ref, see oracle/build_ref.py) on the host CPU, all host threads, on a bounded sample of the same workload;
`--impl cudnn` prints the gpu_baseline measurement as its own line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    "2": dict(metric="tile_pairs_per_sec_gen_disc_fwd_bwd_256x256x13", unit="tile-pairs/s", C=13, H=256, W=256, B=16, gf=227.3,
              cpu_B=16, nets="GD",
              workload="configs[1]: Generator+Discriminator fwd/bwd (+Adam/RMSprop step), batch {B}/GPU of 256x256x13 synthetic tile pairs"),
    "g32": dict(metric="tiles_per_sec_generator_fwd_bwd_256x256x13", unit="tiles/s", C=13, H=256, W=256, B=32, gf=203.6, cpu_B=8,
                nets="G", workload="north star: Generator fwd+bwd (+Adam step), batch {B}/GPU of 13x256x256 synthetic tiles"),
    "3": dict(metric="tile_pairs_per_sec_usss_joint_step_512x512x13", unit="tile-pairs/s", C=13, H=512, W=512, B=32, gf=3961.0, cpu_B=1,
              nets="GS",
              workload="configs[2]: USSS joint iteration G+S+CNetLoss (MS-SSIM weight 0.3, perception 0; USSS has no D, SURVEY.md 8(d)), "
                       "batch {B}/GPU of 512x512x13 synthetic tile pairs (146 GiB of the 178 GiB at batch 32)"),
    "4": dict(metric="tile_pairs_per_sec_rsss_step_256x256x13", unit="tile-pairs/s", C=13, H=256, W=256, B=16, gf=889.0, cpu_B=2,
              nets="GSD", workload="configs[3]: RSSS adversarial iteration G+S+D, batch {B}/GPU of 256x256x13 synthetic OSCD-shape pairs"),
    "5": dict(metric="items_per_sec_wsss_step_256x256x3", unit="items/s", C=3, H=256, W=256, B=32, gf=1654.0, cpu_B=2, nets="GSD",
              workload="configs[4]: WSSS adversarial iteration, batch {B}/GPU of (changed, unchanged) 256x256x3 synthetic WHU-shape pairs"),
}
METRIC = CONFIGS["2"]["metric"]
UNIT = CONFIGS["2"]["unit"]
C, H, W = 13, 256, 256          # defaults of synth() (config 2); scripts may override
BATCH_PER_GPU = 16
GF_PER_PAIR = 227.3


# DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the kernels that can dominate the step, taken from the
# committed `ncu --set full` captures named here (B = 16, 256 x 256, 64 -> 64, parity) — NOT measured by this run: `roofline.traffic_source`
# says so in the JSON line.
NCU_TRAFFIC_BYTES = {
    "conv_wgrad_tc 3x3s1 64->64": (542.5e6, "profiles/r02_ncu_full_kernels_summary.txt"),      # 537.0 MB read + 5.4 MB written
    "conv_fwd_tc 3x3s1 64->64": (495.5e6, "profiles/r02_ncu_full_kernels_summary.txt"),          # 268.7 + 226.9 (conv_halo_kernel)
    "conv_dgrad_tc 3x3s1 64->64": (493.9e6, "profiles/r01_ncu_full_conv_engines_final.txt"),
    "fcd_bn_act_bwd_apply": (774.9e6, "profiles/r02_ncu_full_kernels_summary.txt"),             # 536.9 + 238.0
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"tflops_sustained": d.get("bf16_tflops_sustained", 1373.7), "tflops_burst": d.get("bf16_tflops", 1623.3),
                "hbm_gbs": d.get("hbm_gbs", 6534.5), "source": "measured (MEASURED_PEAKS.json)"}
    return {"tflops_sustained": 1400.0, "tflops_burst": 1590.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


def synth(B, seed, device=None, pin=False, C=None, H=None, W=None):
    """z-scored T1, T2 = T1 + noise with one changed rectangle, region = dilated rectangle, a smooth synthetic
    change-density map (SURVEY.md §8(d) synthetic inputs).  -> [x, y, region, cmap]"""
    import torch

    C = C or globals()["C"]; H = H or globals()["H"]; W = W or globals()["W"]
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, C, H, W, generator=g)
    y = x + 0.3 * torch.randn(B, C, H, W, generator=g)
    region = torch.zeros(B, 1, H, W)
    cmap = 0.1 * torch.rand(B, 1, H, W, generator=g)
    for i in range(B):
        h0, w0 = 40 + 7 * (i % 8), 60 + 5 * (i % 8)
        y[i, :, h0:h0 + 64, w0:w0 + 64] = torch.randn(C, 64, 64, generator=g)
        region[i, :, h0 - 10:h0 + 74, w0 - 10:w0 + 74] = 1
        cmap[i, :, h0:h0 + 64, w0:w0 + 64] = 0.9
    ts = [x, y, region, cmap]
    if pin:
        ts = [t.pin_memory() for t in ts]
    if device is not None:
        ts = [t.to(device) for t in ts]
    return ts


def synth_for(config, B, seed, device=None, pin=False):
    """The input tensors of one step of `config`, in the order its step function takes them."""
    import torch

    cfg = CONFIGS[config]
    x, y, region, cmap = synth(B, seed, C=cfg["C"], H=cfg["H"], W=cfg["W"])
    if config == "2":
        ts = [x, y, region, cmap]
    elif config in ("g32", "3"):
        ts = [x, y]
    elif config == "4":
        ts = [x, y, region]
    else:       # WSSS: a changed pair and an unchanged pair (Demo_WSSS.py:247-270)
        g = torch.Generator().manual_seed(seed + 7)
        x_nc = torch.randn(x.shape, generator=g)
        ts = [x, y, x_nc, x_nc + 0.3 * torch.randn(x.shape, generator=g)]
    if pin:
        ts = [t.pin_memory() for t in ts]
    if device is not None:
        ts = [t.to(device) for t in ts]
    return ts


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def _loss_list(out):
    """Scalar losses of a step's return value (dict from fcdgan_b200.steps, or a tuple), in a fixed order."""
    import torch

    if isinstance(out, dict):
        keys = [k for k in ("Loss", "NetLoss", "d_loss", "s_loss") if k in out]      # the order oracle/ref_steps.py returns them in
        return [out[k] for k in keys]
    if torch.is_tensor(out):
        return [out]
    return list(out)


# ------------------------------------------------------------------------------------------------
def build_ours(config, dev, B, capturable, lean=True):
    """-> (genfn(*data) -> step generator yielding (network, wait) at its exchange points, trained networks, all networks)."""
    import torch

    import fcdgan_b200 as fb
    from fcdgan_b200 import steps as S

    cfg = CONFIGS[config]
    Cc, Hh, Ww = cfg["C"], cfg["H"], cfg["W"]
    nets = {}
    for k, make in (("G", lambda: fb.Generator(Cc)), ("S", lambda: fb.Segmentor(Cc, 1, True)),
                    ("D", lambda: fb.Discriminator_SRGAN_simple(Cc))):
        if k in cfg["nets"]:
            torch.manual_seed(0)          # the reference arm seeds the same way: identical default initialisation
            nets[k] = make().to(dev).train()
    netG, netS, netD = nets.get("G"), nets.get("S"), nets.get("D")
    # Demo_USSS.py:121-122; torch's single-kernel ("fused") implementation of the same update on both GPU arms (ours and the cuDNN
    # baseline): ~15 multi-tensor launches per step otherwise
    adam = lambda n: torch.optim.Adam(n.parameters(), lr=2e-4, betas=(0.9, 0.99), capturable=capturable, fused=True)
    rms = lambda n: torch.optim.RMSprop(n.parameters(), lr=5e-5, capturable=capturable)                      # Demo_RSSS.py:155-158
    recon = fb.losses._MaskedRecon

    if config in ("2", "g32"):
        optG = adam(netG)
        optD = rms(netD) if netD is not None else None
        zero_cmap = torch.zeros(B, 1, Hh, Ww, device=dev)

        def genfn(x, y, region=None, cmap=None):
            # generator iteration (Demo_USSS.py:142-159; perception / ssim weight 0)
            y_fake = netG(x)
            gen_loss, _, _, _ = recon.apply(y, y_fake, zero_cmap, fb.losses.LOSS_L1, False)
            optG.zero_grad(set_to_none=True)
            gen_loss.backward()
            if netD is None:
                yield netG, True
                optG.step()
                return (gen_loss,)
            yield netG, False               # G's bucket is in flight while the discriminator pass runs
            # discriminator update (Demo_RSSS.py:285-307)
            x_mask, y_mask = fb.soft_mask(x, cmap), fb.soft_mask(y, cmap)
            c_out = netD(x_mask, y_mask)
            nc_out = netD(x_mask, fb.soft_mask(y, cmap, other=x, region=region))
            d_loss = 1 + fb.mean(nc_out) - fb.mean(c_out)
            optD.zero_grad(set_to_none=True)
            d_loss.backward()
            yield netD, True
            optG.step()
            optD.step()
            return gen_loss, d_loss

        return genfn, [n for n in (netG, netD) if n is not None], list(nets.values())
    if config == "3":
        optG, optS = adam(netG), adam(netS)
        crit = fb.CNetLoss(channel=Cc)
        return (lambda x, y: S.usss_gen(netG, netS, x, y, crit, optG, optS, ssim_weight=0.3, lean=lean)), [netG, netS], list(nets.values())
    gcrit = fb.CGeneratorLoss(channel=Cc, perception_perBand=(config == "4"))
    netG.eval()                                                                   # Demo_RSSS.py:240, Demo_WSSS.py:207
    optS, optD = rms(netS), rms(netD)
    if config == "4":
        return (lambda x, y, region: S.rsss_gen(netG, netS, netD, x, y, region, gcrit, optS, optD, lean=lean)), [netS, netD], list(nets.values())
    return (lambda x, y, x_nc, y_nc: S.wsss_gen(netG, netS, netD, x, y, x_nc, y_nc, gcrit, optS, optD, lean=lean)), [netS, netD], list(nets.values())


def build_reference(config, dev, B, staged=True):
    """The same step over the UNMODIFIED reference classes on `dev` (cpu: the reference arm; cuda: the cuDNN baseline)."""
    import torch

    from oracle import ref_import, ref_steps as R

    M, L, _ = ref_import.load(prefer_staged=staged)
    cfg = CONFIGS[config]
    Cc, Hh, Ww = cfg["C"], cfg["H"], cfg["W"]
    nets = {}
    for k, make in (("G", lambda: M.Generator(Cc)), ("S", lambda: M.Segmentor(Cc, 1, True)),
                    ("D", lambda: M.Discriminator_SRGAN_simple(Cc))):
        if k in cfg["nets"]:
            torch.manual_seed(0)
            nets[k] = make().to(dev).train()
    netG, netS, netD = nets.get("G"), nets.get("S"), nets.get("D")
    adam = lambda n: torch.optim.Adam(n.parameters(), lr=2e-4, betas=(0.9, 0.99), fused=(torch.device(dev).type == "cuda") or None)
    rms = lambda n: torch.optim.RMSprop(n.parameters(), lr=5e-5)
    if config == "g32":
        optG = adam(netG)
        zero = torch.zeros(B, 1, Hh, Ww, device=dev)
        return lambda x, y: (R.g_step(netG, optG, x, y, zero),)
    if config == "2":
        optG, optD = adam(netG), rms(netD)
        zero = torch.zeros(B, 1, Hh, Ww, device=dev)
        return lambda x, y, region, cmap: R.gd_step(netG, netD, optG, optD, x, y, region, cmap, zero)
    if config == "3":
        optG, optS = adam(netG), adam(netS)
        crit = R.stub_perception(L.CNetLoss(channel=Cc)).to(dev)
        return lambda x, y: R.usss_step(netG, netS, crit, optG, optS, x, y, ssim_weight=0.3)
    gcrit = R.stub_perception(L.CGeneratorLoss(channel=Cc, perception_perBand=(config == "4"))).to(dev)
    netG.eval()
    optS, optD = rms(netS), rms(netD)
    if config == "4":
        return lambda x, y, region: R.rsss_step(netG, netS, netD, gcrit, L.region_loss, optS, optD, x, y, region)
    return lambda x, y, x_nc, y_nc: R.wsss_step(netG, netS, netD, gcrit, optS, optD, x, y, x_nc, y_nc)


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import fcdgan_b200 as fb
    from fcdgan_b200 import engine as E
    from fcdgan_b200 import parallel as P
    from fcdgan_b200.steps import drive

    cfg = CONFIGS[args.config]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    max_seconds = args.max_seconds if args.max_seconds > 0 else (900.0 if world > 1 else 0.0)
    if max_seconds > 0:            # self-destruct: a hung collective must not hold N GPUs until an outer timeout
        t_kill = threading.Timer(max_seconds, lambda: os._exit(3))
        t_kill.daemon = True
        t_kill.start()
    if world > 1:
        local = P.init_from_env("nccl")
    else:
        local = 0
        torch.cuda.set_device(0)
    dev = torch.device("cuda", local)
    fb.set_precision(args.precision)
    fb.set_streams(args.streams)
    B = args.batch or cfg["B"]
    use_graph = args.graph in ("on", "auto", "segmented")

    genfn, nets, all_nets = build_ours(args.config, dev, B, use_graph, lean=(args.form == "lean"))
    P.broadcast_parameters(nets)
    sync = P.GradSync()
    data = synth_for(args.config, B, 1234 + rank, device=dev)

    def eager_step(*inputs):
        return _loss_list(drive(genfn(*inputs), sync.on_grads))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the very first iteration from the initial weights, eagerly: its losses are what the CPU leg (unmodified reference, same
    # seeds, same batch) must reproduce — checked below when the CPU sample runs the full batch
    first_losses = [float(v.detach()) for v in eager_step(*data)]
    torch.cuda.empty_cache()        # the eager iteration's blocks must not sit beside the graph's private pool (config 3 is memory-bound)

    step = eager_step
    gstep = None
    graph_note = "eager"
    if use_graph:
        try:
            if world == 1 and args.graph != "segmented":
                from fcdgan_b200.graph import GraphedStep
                gstep = GraphedStep(lambda *a: _loss_list(drive(genfn(*a))), data, warmup=3, modules=all_nets)
                graph_note = "whole iteration captured in one CUDA graph (fcdgan_b200.graph.GraphedStep)"
            elif args.collectives == "captured" and world > 1:
                # experiment: the NCCL all-reduces captured INSIDE the one graph (round 1: hung on this stack)
                from fcdgan_b200.graph import GraphedStep
                gstep = GraphedStep(lambda *a: _loss_list(drive(genfn(*a), sync.on_grads)), data, warmup=3,
                                    capture_error_mode="thread_local", modules=all_nets)
                graph_note = "whole iteration INCLUDING the NCCL all-reduces captured in one CUDA graph"
            else:
                from fcdgan_b200.graph import YieldingStep
                gstep = YieldingStep(genfn, sync, data, warmup=3, modules=all_nets)
                graph_note = (f"{len(gstep.graphs)} CUDA graphs cut at the step's gradient-exchange points, NCCL all-reduces issued "
                              "eagerly between them (fcdgan_b200.graph.YieldingStep)")

            def step(*inputs):                      # noqa: F811 — replay; inputs are copied into the static buffers
                if inputs and inputs[0] is not data[0]:
                    gstep.copy_inputs(*inputs)
                return _loss_list(gstep())
        except Exception as e:   # capture is an optimisation, never a requirement
            step = eager_step
            graph_note = f"eager (graph capture failed: {type(e).__name__}: {str(e)[:120]})"
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()           # started BEFORE the warm-up: nvidia-smi's start-up stalls the driver for ~0.5 s
    # eager mode: torch's caching allocator needs ~8 iterations to reach a steady block layout (cudaMalloc bursts until
    # then); the extra untimed iterations only apply when the step is not replayed from a graph.
    n_warm = args.warmup if step is not eager_step else max(args.warmup, 8)
    for _ in range(n_warm):
        step(*data)
    barrier()
    sampler.rows.clear()          # keep only samples taken during the timed region
    l0 = E.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    e0.record()
    marks[0].record()
    t_host = time.perf_counter()
    for i in range(args.steps):
        step(*data)
        marks[i + 1].record()
    e1.record()
    host_ms = (time.perf_counter() - t_host) * 1e3 / args.steps     # CPU time to ISSUE one step (no sync inside)
    barrier()
    ms = e0.elapsed_time(e1)
    per_step = sorted(marks[i].elapsed_time(marks[i + 1]) for i in range(args.steps))
    launches = E.launch_count - l0
    if step is not eager_step:
        launches = gstep.launches_per_replay * args.steps     # replayed from the captured graph
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = t.item()
    value = world * B * args.steps / (ms_total / 1e3)
    peak_mem = torch.cuda.max_memory_allocated() / 2 ** 30

    # ---- end to end: pinned host inputs -> H2D -> step -> losses read back, every step.
    # The inputs of step i+1 are uploaded on a copy stream while step i computes (what a DataLoader with pinned memory and a
    # prefetch queue does); every step's H2D copy and D2H loss read-back happen inside the timed region.
    host = synth_for(args.config, B, 1234 + rank, pin=True)
    h2d = sum(t_.numel() * 4 for t_ in host)
    copy_stream = torch.cuda.Stream(device=dev)
    staging = [[torch.empty_like(t_, device=dev) for t_ in host] for _ in range(2)]
    uploaded = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    n_loss = len(first_losses)
    host_losses = torch.empty(n_loss, dtype=torch.float32).pin_memory()

    def upload(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])          # the step that read this slot has been issued and has run
            for d_, h_ in zip(staging[slot], host):
                d_.copy_(h_, non_blocking=True)
            uploaded[slot].record(copy_stream)

    def e2e_loop(n):
        main = torch.cuda.current_stream()
        for ev in consumed:
            ev.record(main)
        upload(0)
        out_l = None
        for i in range(n):
            slot = i & 1
            if i + 1 < n:
                upload(slot ^ 1)
            main.wait_event(uploaded[slot])
            ls = step(*staging[slot])                        # graphed: device-to-device into the graph's static inputs
            consumed[slot].record(main)
            host_losses.copy_(torch.stack([v.detach() for v in ls]), non_blocking=True)
            main.synchronize()                               # the step's result is on the host before the next step is issued
            out_l = host_losses.clone()
        return out_l

    e2e_loop(2)
    barrier()
    e0.record()
    losses = e2e_loop(args.steps)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * B * args.steps / (t.item() / 1e3)

    # ---- instrumented pass: per-call CUDA events -> dominant kernel roofline (not part of the timed numbers).
    # Every rank runs it (the step contains collectives when world > 1); only rank 0 records events.
    gstep = None                  # release the graphs' private pool first: the eager pass needs a whole iteration's memory again
    step = eager_step
    del staging
    torch.cuda.empty_cache()
    nprof = min(args.steps, 3)
    fb.set_streams(1)             # one stream, so that the events bracket each kernel alone
    if rank == 0:
        E.PROFILE = []
    for _ in range(nprof):
        eager_step(*data)
    barrier()
    out = None
    if rank == 0:
        agg = {}
        for name, tag, flops, nbytes, a, b in E.PROFILE:
            d = agg.setdefault(tag, [0.0, 0, 0.0, name])
            d[0] += a.elapsed_time(b); d[1] += 1; d[2] += flops
        E.PROFILE = None
        tot = sum(v[0] for v in agg.values())
        top = sorted(agg.items(), key=lambda kv: -kv[1][0])
        if args.profile_out:
            with open(args.profile_out, "w") as fh:
                fh.write(f"# per-step CUDA-event time by C-ABI call (instrumented pass, {nprof} steps, config {args.config}, precision "
                         f"{args.precision}); total {tot / nprof:.3f} ms/step\n# tag | ms/step | launches/step | share | TFLOP/s (algorithmic)\n")
                for k, v in top:
                    tf = v[2] / (v[0] * 1e-3) / 1e12 if v[0] > 0 and v[2] > 0 else 0.0
                    fh.write(f"{k:44s} {v[0] / nprof:9.3f} {v[1] // nprof:5d} {v[0] / tot:7.3f} {tf:8.1f}\n")
        pk = peaks()
        # the roofline entry is the dominant TENSOR kernel (the conv stack is the tensor-bound part of the path)
        conv_top = [kv for kv in top if kv[1][2] > 0] or top
        dom_tag, (dom_ms, dom_n, dom_flops, _) = conv_top[0]
        tc_ms = sum(v[0] for k, v in agg.items() if "_tc" in k)
        tc_fl = sum(v[2] for k, v in agg.items() if "_tc" in k)
        achieved = dom_flops / (dom_ms * 1e-3) / 1e12 if dom_ms > 0 else 0.0
        traffic = NCU_TRAFFIC_BYTES.get(dom_tag) if (args.precision == "parity" and args.config in ("2", "g32") and B == 16) else None
        roof = {"bound": "tensor", "kernel": dom_tag, "achieved": round(achieved, 1), "peak": pk["tflops_sustained"],
                "unit": "TFLOP/s", "frac": round(achieved / pk["tflops_sustained"], 4),
                "traffic": traffic[0] if traffic else None,
                "traffic_source": (f"{traffic[1]} (ncu --set full capture of this kernel at this shape; not measured by this run)"
                                   if traffic else None),
                "launches_per_step": dom_n // nprof, "avg_launch_ms": round(dom_ms / dom_n, 4),
                "share_of_step": round(dom_ms / tot, 3), "peak_source": pk["source"] + ", sustained bf16",
                "tc_conv_all": {"tflops": round(tc_fl / (tc_ms * 1e-3) / 1e12, 1) if tc_ms else None,
                                "share_of_step": round(tc_ms / tot, 3)},
                "step_algorithmic_tflops": round(cfg["gf"] * 1e9 * world * B * args.steps / (ms_total * 1e-3) / 1e12, 1),
                "step_frac_of_peak": round(cfg["gf"] * 1e9 * B * args.steps / (ms_total * 1e-3) / 1e12 / pk["tflops_sustained"], 4),
                "top5": [{"kernel": k, "ms_per_step": round(v[0] / nprof, 3), "launches": v[1] // nprof} for k, v in top[:5]]}
        gpu_base = cpu = None
        loss_check = None
        if world == 1 and not args.no_gpu_baseline:
            Bg = B
            while True:      # the reference's autograd keeps far more per sample than our tape: halve its batch until it fits
                try:
                    torch.cuda.empty_cache()
                    gpu_base = gpu_baseline(args.config, Bg, dev)
                    break
                except torch.OutOfMemoryError:
                    if Bg == 1:
                        gpu_base = {"unavailable": "out of memory at batch 1"}
                        break
                    Bg //= 2
                except Exception as e:
                    gpu_base = {"unavailable": f"{type(e).__name__}: {str(e)[:160]}"}
                    break
        if world == 1 and not args.no_cpu_baseline:
            cpu = cpu_baseline(args.config)
            if cpu.get("first_losses") and cpu.get("batch") == B:
                errs = [abs(a - b) / max(abs(b), 1e-6) for a, b in zip(first_losses, cpu["first_losses"])]
                loss_check = {"max_rel_err": round(max(errs), 7), "tolerance": 1e-3, "ok": max(errs) < 1e-3,
                              "what": "first-iteration losses, ours (GPU) vs the unmodified reference (CPU), same seeds / weights / batch"}
                if not loss_check["ok"]:      # reported in the line (tests/test_bench_gpu.py asserts it); stderr says it loudly
                    print(f"bench.py: LOSS CHECK FAILED: ours {first_losses} vs reference {cpu['first_losses']}", file=sys.stderr)
        out = {"metric": cfg["metric"], "value": round(value, 2), "unit": cfg["unit"], "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": round(ms_total / args.steps, 3),
               "ms_per_step_min_med_max": [round(per_step[0], 3), round(per_step[len(per_step) // 2], 3), round(per_step[-1], 3)],
               "host_issue_ms_per_step": round(host_ms, 3), "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "bf16x3-split (fp32-class)" if args.precision == "parity" else "bf16",
               "data": "synthetic",
               "config": {"workload": cfg["workload"].format(B=B), "precision": args.precision, "batch_per_gpu": B,
                          "parallelism": f"dp{world}", "launch": graph_note, "streams": args.streams,
                          "step_form": (None if args.config in ("2", "g32") else
                                        "lean: the backward sweeps whose results the reference's loop body discards are skipped (identical "
                                        "losses / map / gradients at every optimizer step: tests/test_steps_gpu.py::test_lean_steps_*; the "
                                        "algorithmic minimum SURVEY.md 8(d) counts)" if args.form == "lean" else
                                        "faithful: the reference's call sequence, call for call (its dead backward sweeps included)"),
                          "l2": "per-step working set (>20 GB) >> 126 MB L2; no flush needed",
                          "algorithmic_gflop_per_unit": cfg["gf"], "peak_mem_GiB": round(peak_mem, 1),
                          "engine": {k: fb.engine._cfg[k] for k in ("rowpack", "batch_branches", "im2col", "fuse_stats")},
                          "optimizer": "torch.optim.Adam(fused=True) / RMSprop (foreach), capturable"},
               "clocks": clocks,
               "e2e": {"value": round(e2e_value, 2), "unit": cfg["unit"], "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4 * n_loss},
               "gpu_launches": launches, "gpu_launches_note": "libfcd_b200 C-ABI calls in the timed region (each launches >= 1 kernel)",
               "roofline": roof, "cpu_baseline": cpu, "gpu_baseline": gpu_base, "first_losses": [round(v, 6) for v in first_losses],
               "loss_check": loss_check, "final_losses": [round(float(v), 5) for v in losses.tolist()]}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return out


# ------------------------------------------------------------------------------------------------
def gpu_baseline(config, B, dev, steps=5, warmup=3):
    """torch + cuDNN on the same B200: the UNMODIFIED reference classes moved to the GPU, same step / batch / seeds, eager (the
    way the demos run), CUDA-event timed; `tf32` = PyTorch's defaults (cuDNN convolutions may use TF32 — what Module.py actually
    does on a GPU), `fp32` = torch.backends.cudnn.allow_tf32 = False."""
    import torch

    cfg = CONFIGS[config]
    out = {"unit": cfg["unit"], "batch": B, "what": "unmodified reference classes (.cuda()), torch " + torch.__version__ +
           f" + cuDNN {torch.backends.cudnn.version()}, eager, {steps} steps after {warmup} warm-up"}
    data = synth_for(config, B, 1234, device=dev)
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    try:
        for name, tf32 in (("tf32", True), ("fp32", False)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = False     # PyTorch default
            step = build_reference(config, dev, B)
            first = None
            for i in range(warmup):
                ls = step(*data)
                if i == 0:
                    first = [float(v) for v in ls]
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                step(*data)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[name] = {"value": round(B / ms * 1e3, 2), "ms_per_step": round(ms, 3), "first_losses": [round(v, 6) for v in first],
                         "algorithmic_tflops": round(cfg["gf"] * B / ms, 1)}
            del step
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = saved
    return out


def cpu_baseline(config, steps=6, warmup=1, budget_s=30.0, batch=None):
    """The UNMODIFIED reference classes on all host cores: up to `steps` iterations of the same step (after `warmup`), stopping
    once `budget_s` seconds of timed work are spent.  Batch = the workload's own where one CPU step takes seconds (configs 2,
    g32 -> the first iteration doubles as the loss check of the GPU run), a smaller sample of it otherwise."""
    import torch

    from oracle import ref_import

    cfg = CONFIGS[config]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    Bc = batch or cfg["cpu_B"]
    if ref_import.importable():
        step = build_reference(config, torch.device("cpu"), Bc)
        kind, what = "reference", "unmodified reference classes (oracle/_ref staged copy of Module.py / Loss.py / ssim.py)"
    else:       # no staged reference on this box: the oracle port (config 2 only)
        if config != "2":
            return {"unavailable": "oracle/_ref not staged (run __graft_entry__.build() where /root/reference is mounted)"}
        step = _oracle_port_step(Bc)
        kind, what = "port", "torch CPU fp32 oracle port of the PyTorch reference (oracle/_ref not staged)"
    data = synth_for(config, Bc, 1234)
    first = None
    t_w = time.perf_counter()
    for i in range(warmup):
        ls = step(*data)
        if i == 0:
            first = [float(v) for v in ls]
    t_w = time.perf_counter() - t_w
    t0 = time.perf_counter()
    done = 0
    for _ in range(max(1, steps)):
        ls = step(*data)
        if first is None:
            first = [float(v) for v in ls]
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = (time.perf_counter() - t0) / done
    return {"value": round(Bc / dt, 4), "unit": cfg["unit"], "cores": cores, "kind": kind, "steps": done, "batch": Bc,
            "first_losses": [round(v, 6) for v in first],
            "sample": f"{done} step(s) (+{warmup} warm-up) of the same iteration at batch {Bc} (of {cfg['B']}) on {cores} host threads, "
                      f"{what}, perception term stubbed to 0 (weight 0 in this workload); {dt:.2f} s/step"}


def _oracle_port_step(Bc):
    """Config 2 restated with the oracle port (oracle/fcd_oracle.py) — only used when oracle/_ref is not staged."""
    import torch

    from oracle import fcd_oracle as O

    sdG = O.clone_sd(O.make_state_dict(O.generator_spec(13), 11), requires_grad=True)
    sdD = O.clone_sd(O.make_state_dict(O.discriminator_spec(13), 13), requires_grad=True)
    zero = torch.zeros(Bc, 1, 256, 256)

    def step(x, y, region, cmap):
        for sd in (sdG, sdD):
            for v in sd.values():
                v.grad = None
        y_fake = O.generator(sdG, x, train=True)
        gl = O.masked_recon_loss(y, y_fake, zero, "l1", skip_empty=False)
        gl.backward()
        m = 1 - cmap
        c_out = O.discriminator(sdD, x * m, y * m, train=True)
        nc_out = O.discriminator(sdD, x * m, (y * (1 - region) + x * region) * m, train=True)
        dl = 1 + nc_out.mean() - c_out.mean()
        dl.backward()
        return gl, dl

    return step


def run_reference(args):
    """Reference arm: the reference's own implementation of the path on the host CPU — the UNMODIFIED Module.py / Loss.py /
    ssim.py classes (staged copy oracle/_ref; the reference is a set of flat scripts that cannot be pip-installed),
    all host threads, same workload, each step a bounded sample of it."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    cfg = CONFIGS[args.config]
    steps, warmup = max(1, args.steps), max(0, min(args.warmup, 1))
    cb = cpu_baseline(args.config, steps=steps, warmup=warmup, budget_s=150.0, batch=args.batch)
    if "unavailable" in cb:
        return {"impl": "reference", "unavailable": cb["unavailable"]}
    return {"impl": "reference", "metric": cfg["metric"], "value": cb["value"], "unit": cfg["unit"],
            "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": cb["steps"], "warmup": warmup,
            "ms_per_step": round(cb["batch"] / cb["value"] * 1e3, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": cfg["workload"].format(B=cfg["B"]) + f"; CPU bounded sample (batch {cb['batch']} per step)"},
            "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": cfg["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def run_cudnn(args):
    """`--impl cudnn`: the gpu_baseline measurement as its own line (one GPU; other ranks exit)."""
    import torch

    if int(os.environ.get("RANK", "0")) != 0:
        return None
    cfg = CONFIGS[args.config]
    B = args.batch or cfg["B"]
    torch.cuda.set_device(0)
    gb = gpu_baseline(args.config, B, torch.device("cuda", 0), steps=max(1, args.steps), warmup=max(1, args.warmup))
    return {"impl": "cudnn", "metric": cfg["metric"], "value": gb["tf32"]["value"], "unit": cfg["unit"], "n_gpus": 1,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": gb["tf32"]["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "tf32 (PyTorch default for cuDNN convolutions)", "data": "synthetic",
            "config": {"workload": cfg["workload"].format(B=B)}, "gpu_baseline": gb}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", choices=sorted(CONFIGS), default="2")
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: the configuration's)")
    ap.add_argument("--precision", choices=["parity", "fast"], default="parity")
    ap.add_argument("--impl", choices=["ours", "reference", "cudnn"], default="ours")
    ap.add_argument("--graph", choices=["auto", "on", "off", "segmented"], default="auto",
                    help="capture the iteration in a CUDA graph (segmented: the N > 1 form — graphs cut at the exchange points — also at N = 1)")
    ap.add_argument("--form", choices=["lean", "faithful"], default="lean",
                    help="configs 3/4/5: skip the backward sweeps the reference's loop body throws away (lean) or replay its sequence call for call")
    ap.add_argument("--collectives", choices=["eager", "captured"], default="eager",
                    help="N > 1: NCCL calls issued eagerly between CUDA graphs (default) or captured inside one graph (experiment)")
    ap.add_argument("--streams", type=int, default=1, help="engine streams (independent branches / weight gradients run concurrently)")
    ap.add_argument("--max-seconds", type=float, default=0.0, help="hard-exit the process after this many seconds (0 = off at N = 1, 900 at N > 1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true")
    ap.add_argument("--profile-out", default=None, help="write the per-call CUDA-event table of the instrumented pass here")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    # stdout carries exactly ONE line, the JSON record: anything libraries print while running (NCCL's version banner, ...)
    # is routed to stderr by pointing fd 1 at fd 2 until the record is ready
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        out = {"reference": run_reference, "cudnn": run_cudnn, "ours": run_ours}[args.impl](args)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    if out is not None:
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
