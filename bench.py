#!/usr/bin/env python
"""bench.py — headline benchmark of the FCD-GAN hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--precision parity|fast] [--impl ours|reference]

Workload (BASELINE.json configs[1]): Generator + Discriminator forward/backward on a
batch of 16 synthetic 13-band 256x256 bi-temporal tile pairs per GPU:
    y_fake = G(x);  generator_loss = CNetLoss-style masked L1(y, y_fake, cmap=0)  -> backward -> Adam step  (Demo_USSS.py:142-159)
    c_out = D(x*(1-cmap), y*(1-cmap));  nc_out = D(x*(1-cmap), (y*(1-r)+x*r)*(1-cmap))
    d_loss = 1 + mean(nc_out) - mean(c_out) -> backward -> RMSprop step                                      (Demo_RSSS.py:285-307)
Algorithmic conv FLOPs per tile pair: G 203.6 GF + D 2 x 11.9 GF = 227.3 GF.
One "step" = that whole iteration, optimizer steps included.  N > 1 (torchrun): every rank runs its own 16 pairs
(weak scaling) and the gradients of G and D are all-reduced over NCCL (fcdgan_b200.parallel.GradSync).

Prints ONE JSON line (rank 0).  `value` is device-timed with inputs resident in HBM; `e2e` repeats the measurement
through the same public API with the inputs in pinned host memory (H2D inside the timed region, losses read back).
`roofline` is the dominant kernel of the step (by summed CUDA-event time over an instrumented pass): algorithmic
FLOPs per launch / mean launch time against MEASURED_PEAKS.json's sustained bf16 peak.
`--impl reference` times the reference's own algorithm on the host CPU (the oracle port of the PyTorch reference,
all host threads) on a bounded sample of the same workload.
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

METRIC = "tile_pairs_per_sec_gen_disc_fwd_bwd_256x256x13"
UNIT = "tile-pairs/s"
C, H, W = 13, 256, 256
BATCH_PER_GPU = 16
GF_PER_PAIR = 227.3


# DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the kernels that can dominate the step, from the
# committed `ncu --set full` capture profiles/r01_ncu_full_conv_engines_final.txt (B = 16, 256 x 256, 64 -> 64, parity).
NCU_TRAFFIC_BYTES = {
    "conv_wgrad_tc 3x3s1 64->64": 541.1e6,   # profiles/r01_ncu_wgrad_halo_roundrobin.txt (537.1 MB read + 4.0 MB written)
    "conv_fwd_tc 3x3s1 64->64": 494.2e6,
    "conv_dgrad_tc 3x3s1 64->64": 493.9e6,
    "fcd_bn_act_bwd_apply": 778.4e6,
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"tflops_sustained": d.get("bf16_tflops_sustained", 1373.7), "tflops_burst": d.get("bf16_tflops", 1623.3),
                "hbm_gbs": d.get("hbm_gbs", 6534.5), "source": "measured (MEASURED_PEAKS.json)"}
    return {"tflops_sustained": 1400.0, "tflops_burst": 1590.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


def synth(B, seed, device=None, pin=False):
    """z-scored T1, T2 = T1 + noise with one changed rectangle, region = dilated rectangle, a smooth synthetic
    change-density map (SURVEY.md §8(d) synthetic inputs)."""
    import torch

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


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    import fcdgan_b200 as fb
    from fcdgan_b200 import engine as E
    from fcdgan_b200 import parallel as P

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
    B = BATCH_PER_GPU

    torch.manual_seed(0)
    netG, netD = fb.Generator(C).to(dev).train(), fb.Discriminator_SRGAN_simple(C).to(dev).train()
    P.broadcast_parameters([netG, netD])
    # world > 1: capturing the NCCL all-reduces INSIDE a graph hung on this stack (torch 2.11 / NCCL 2.28.9), so
    # the iteration is captured as three graphs with the collectives issued eagerly between them.
    use_graph = args.graph in ("on", "auto")
    optG = torch.optim.Adam(netG.parameters(), lr=2e-4, betas=(0.9, 0.99), capturable=use_graph)   # Demo_USSS.py:121
    optD = torch.optim.RMSprop(netD.parameters(), lr=5e-5, capturable=use_graph)                   # Demo_RSSS.py:157
    crit = fb.losses._MaskedRecon
    sync = P.GradSync()
    zero_cmap = torch.zeros(B, 1, H, W, device=dev)

    def step(x, y, region, cmap):
        # generator iteration (Demo_USSS.py:142-159, perception weight 0 — out of scope)
        y_fake = netG(x)
        gen_loss, _, _, _ = crit.apply(y, y_fake, zero_cmap, fb.losses.LOSS_L1, False)
        optG.zero_grad(set_to_none=True)
        gen_loss.backward()
        sync.start(netG)
        # discriminator iteration (Demo_RSSS.py:285-307)
        x_mask, y_mask = fb.soft_mask(x, cmap), fb.soft_mask(y, cmap)
        c_out = netD(x_mask, y_mask)
        y_unc = fb.soft_mask(y, cmap, other=x, region=region)
        nc_out = netD(x_mask, y_unc)
        d_loss = 1 + fb.mean(nc_out) - fb.mean(c_out)
        optD.zero_grad(set_to_none=True)
        d_loss.backward()
        sync.start(netD)
        sync.finish()
        optG.step()
        optD.step()
        return gen_loss, d_loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # N > 1: the same iteration cut at its two exchange points; the pieces are CUDA graphs, the all-reduces run eagerly
    # between them (fcdgan_b200.graph.SegmentedStep)
    def seg_g(x, y, region, cmap):
        y_fake = netG(x)
        gen_loss, _, _, _ = crit.apply(y, y_fake, zero_cmap, fb.losses.LOSS_L1, False)
        optG.zero_grad(set_to_none=True)
        gen_loss.backward()
        sync.pack(netG)
        return gen_loss

    def seg_d(x, y, region, cmap):
        x_mask, y_mask = fb.soft_mask(x, cmap), fb.soft_mask(y, cmap)
        c_out = netD(x_mask, y_mask)
        nc_out = netD(x_mask, fb.soft_mask(y, cmap, other=x, region=region))
        d_loss = 1 + fb.mean(nc_out) - fb.mean(c_out)
        optD.zero_grad(set_to_none=True)
        d_loss.backward()
        sync.pack(netD)
        return d_loss

    def seg_opt(x, y, region, cmap):
        sync.unpack()
        optG.step()
        optD.step()

    def exchange_d():
        sync.launch(netD)
        sync.wait()

    data = synth(B, 1234 + rank, device=dev)
    eager_step = step
    graph_note = "eager"
    if use_graph:
        try:
            if world == 1:
                from fcdgan_b200.graph import GraphedStep
                gstep = GraphedStep(eager_step, data, warmup=3)
                graph_note = "whole iteration captured in one CUDA graph (fcdgan_b200.graph.GraphedStep)"
            else:
                from fcdgan_b200.graph import SegmentedStep
                gstep = SegmentedStep([seg_g, seg_d, seg_opt], [lambda: sync.launch(netG), exchange_d], data, warmup=3)
                graph_note = ("three CUDA graphs (G pass / D pass / optimizer steps) with the two NCCL all-reduces issued eagerly "
                              "between them (fcdgan_b200.graph.SegmentedStep)")

            def step(*inputs):                      # noqa: F811 — replay; inputs are copied into the static buffers
                if inputs and inputs[0] is not data[0]:
                    gstep.copy_inputs(*inputs)
                out_ = gstep()
                return (out_[0], out_[1]) if world > 1 else out_
        except Exception as e:   # capture is an optimisation, never a requirement
            step = eager_step
            graph_note = f"eager (graph capture failed: {type(e).__name__}: {str(e)[:120]})"
    # L2 note: one step streams > 20 GB of activations through a 126 MB L2, so nothing survives between steps.
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
        gl, dl = step(*data)
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

    # ---- end to end: pinned host inputs -> H2D -> step -> losses read back, every step.
    # The inputs of step i+1 are uploaded on a copy stream while step i computes (what a DataLoader with pinned memory and a
    # prefetch queue does); every step's H2D copy and D2H loss read-back happen inside the timed region.
    host = synth(B, 1234 + rank, pin=True)
    h2d = sum(t_.numel() * 4 for t_ in host)
    graphed = step is not eager_step
    copy_stream = torch.cuda.Stream(device=dev)
    staging = [[torch.empty_like(t_, device=dev) for t_ in host] for _ in range(2)]
    uploaded = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    host_losses = torch.empty(2, dtype=torch.float32).pin_memory()

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
            gl_, dl_ = step(*staging[slot])                  # graphed: device-to-device into the graph's static inputs
            consumed[slot].record(main)
            host_losses.copy_(torch.stack([gl_.detach(), dl_.detach()]), non_blocking=True)
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
    nprof = min(args.steps, 3)
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
                fh.write(f"# per-step CUDA-event time by C-ABI call (instrumented pass, {nprof} steps, precision {args.precision}); "
                         f"total {tot / nprof:.3f} ms/step\n# tag | ms/step | launches/step | share | TFLOP/s (algorithmic)\n")
                for k, v in top:
                    tf = v[2] / (v[0] * 1e-3) / 1e12 if v[0] > 0 and v[2] > 0 else 0.0
                    fh.write(f"{k:44s} {v[0] / nprof:9.3f} {v[1] // nprof:5d} {v[0] / tot:7.3f} {tf:8.1f}\n")
        pk = peaks()
        dom_tag, (dom_ms, dom_n, dom_flops, _) = top[0]
        # all tcgen05 conv launches together (the conv stack is the tensor-bound part of the step)
        tc_ms = sum(v[0] for k, v in agg.items() if "_tc " in k)
        tc_fl = sum(v[2] for k, v in agg.items() if "_tc " in k)
        achieved = dom_flops / (dom_ms * 1e-3) / 1e12 if dom_ms > 0 else 0.0
        roof = {"bound": "tensor", "kernel": dom_tag, "achieved": round(achieved, 1), "peak": pk["tflops_sustained"],
                "unit": "TFLOP/s", "frac": round(achieved / pk["tflops_sustained"], 4),
                "traffic": NCU_TRAFFIC_BYTES.get(dom_tag) if args.precision == "parity" else None,
                "launches_per_step": dom_n // nprof, "avg_launch_ms": round(dom_ms / dom_n, 4),
                "share_of_step": round(dom_ms / tot, 3), "peak_source": pk["source"] + ", sustained bf16",
                "tc_conv_all": {"tflops": round(tc_fl / (tc_ms * 1e-3) / 1e12, 1) if tc_ms else None,
                                "share_of_step": round(tc_ms / tot, 3)},
                "step_algorithmic_tflops": round(GF_PER_PAIR * 1e9 * B * args.steps / (ms_total * 1e-3) / 1e12 * (1 if world == 1 else 1), 1),
                "top5": [{"kernel": k, "ms_per_step": round(v[0] / nprof, 3), "launches": v[1] // nprof} for k, v in top[:5]]}
        cpu = cpu_baseline(bounded=True) if world == 1 and not args.no_cpu_baseline else None
        out = {"metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": round(ms_total / args.steps, 3),
               "ms_per_step_min_med_max": [round(per_step[0], 3), round(per_step[len(per_step) // 2], 3), round(per_step[-1], 3)],
               "host_issue_ms_per_step": round(host_ms, 3), "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "bf16x3-split (fp32-class)" if args.precision == "parity" else "bf16",
               "data": "synthetic",
               "config": {"workload": "configs[1]: Generator+Discriminator fwd/bwd (+Adam/RMSprop step), batch 16/GPU of 256x256x13 "
                                      "synthetic tile pairs", "precision": args.precision, "batch_per_gpu": B,
                          "parallelism": f"dp{world}", "launch": graph_note, "l2": "per-step working set (>20 GB) >> 126 MB L2; no flush needed",
                          "algorithmic_gflop_per_pair": GF_PER_PAIR},
               "clocks": clocks,
               "e2e": {"value": round(e2e_value, 2), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 8},
               "gpu_launches": launches, "gpu_launches_note": "libfcd_b200 C-ABI calls in the timed region (each launches >= 1 kernel)",
               "roofline": roof, "cpu_baseline": cpu, "final_losses": [round(float(v), 5) for v in losses.tolist()]}
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return out


# ------------------------------------------------------------------------------------------------
def cpu_step_fn(Bc):
    """The same iteration restated on the CPU with the oracle port of the reference (oracle/fcd_oracle.py)."""
    import torch

    from oracle import fcd_oracle as O

    sdG = O.clone_sd(O.make_state_dict(O.generator_spec(C), 11), requires_grad=True)
    sdD = O.clone_sd(O.make_state_dict(O.discriminator_spec(C), 13), requires_grad=True)
    x, y, region, cmap = [t[:Bc] for t in synth(max(Bc, 1), 1234)]
    zero = torch.zeros(Bc, 1, H, W)

    def step():
        for sd in (sdG, sdD):
            for v in sd.values():
                v.grad = None
        y_fake = O.generator(sdG, x, train=True)
        O.masked_recon_loss(y, y_fake, zero, "l1", skip_empty=False).backward()
        m = 1 - cmap
        c_out = O.discriminator(sdD, x * m, y * m, train=True)
        nc_out = O.discriminator(sdD, x * m, (y * (1 - region) + x * region) * m, train=True)
        (1 + nc_out.mean() - c_out.mean()).backward()

    return step


def cpu_baseline(bounded=True, steps=6, warmup=1, Bc=4, budget_s=150.0):
    """Times the oracle port on all host cores: `steps` iterations of the same G+D step at batch `Bc` (a bounded sample of
    the batch-16 workload, about 10-30 s of CPU work on the GPU box), stopping early once `budget_s` is spent."""
    import torch

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step = cpu_step_fn(Bc)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        step()
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = (time.perf_counter() - t0) / done
    return {"value": round(Bc / dt, 4), "unit": UNIT, "cores": cores, "kind": "port", "steps": done,
            "sample": f"{done} step(s) (+{warmup} warm-up) of the same G+D iteration at batch {Bc} (of 16) on {cores} host threads, "
                      f"torch CPU fp32 oracle port of the PyTorch reference; {dt:.2f} s/step"}


def run_reference(args):
    """Reference arm: the reference's own algorithm for this path on the host CPU.  The reference is a set of flat Python
    scripts that cannot be installed or shipped to the GPU box, so this times the oracle port
    (oracle/fcd_oracle.py, pinned against the unmodified reference by tests/test_oracle_golden.py) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    steps, warmup = max(1, args.steps), max(0, min(args.warmup, 1))
    cb = cpu_baseline(bounded=True, steps=steps, warmup=warmup, Bc=4, budget_s=150.0)
    return {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
            "steps": cb["steps"], "warmup": warmup, "ms_per_step": round(4 / cb["value"] * 1e3, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "configs[1]: Generator+Discriminator fwd/bwd, 256x256x13 synthetic tile pairs; CPU bounded sample "
                                   "(batch 4 per step)"},
            "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--precision", choices=["parity", "fast"], default="parity")
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--graph", choices=["auto", "on", "off"], default="auto", help="capture the iteration in a CUDA graph")
    ap.add_argument("--max-seconds", type=float, default=0.0, help="hard-exit the process after this many seconds (0 = off at N = 1, 900 at N > 1)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-out", default=None, help="write the per-call CUDA-event table of the instrumented pass here")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    # stdout carries exactly ONE line, the JSON record: anything libraries print while running (NCCL's version banner, ...)
    # is routed to stderr by pointing fd 1 at fd 2 until the record is ready
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        out = run_reference(args) if args.impl == "reference" else run_ours(args)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    if out is not None:
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
