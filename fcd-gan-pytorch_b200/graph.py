"""CUDA-graph capture of a whole training iteration.

A network pass here is a few hundred short C-ABI launches issued from Python; at B200 speeds the host needs longer
to ISSUE an iteration (~30 ms for Generator + Discriminator) than the GPU needs to run it.  Capturing the iteration
(forward, hand-written backward, optimizer steps — everything is stream-ordered and allocation goes through torch's
graph-private pool) into one CUDA graph removes that bound: a replay is a single launch.

    step = GraphedStep(fn, static_inputs)      # fn(*static_inputs) -> tensors; runs warm-up iterations, then captures
    out = step()                               # replay; `out` are the static output tensors
    step.copy_inputs(*new_inputs)              # refill the static input buffers (e.g. from pinned host memory)

Requirements (the usual CUDA-graph rules): shapes fixed; optimizers constructed with `capturable=True`;
`zero_grad(set_to_none=True)` inside `fn`; no host synchronisation inside `fn`.

`modules=` — every network the step touches.  Their `.grad` tensors are dropped after the warm-up, BEFORE the capture: a
gradient tensor left over from the warm-up lives in the ordinary allocator pool, and a step that accumulates into it before
zeroing it (the wasted Segmentor gradients of `d_loss.backward()` in Demo_RSSS.py:305, the first of the two backward sweeps of
Demo_USSS.py:327) would bake its address into the graph and then free it — every replay would write through a dangling
pointer (found in round 2: replays of the RSSS / USSS / WSSS steps faulted as soon as the allocator returned that block to the
driver).  With the gradients dropped, the capture allocates them from the graph's private pool, where the address stays valid.

The warm-up iterations (and the capture pass does not execute anything) are REAL iterations: optimizers step, BatchNorm
running statistics and `num_batches_tracked` advance.  Pass `optimizers=` and `restore_after_warmup=True` to have parameters,
buffers and optimizer state put back, in place, to their values from before the warm-up once the capture is done.

Do not keep autograd-connected tensors of EARLIER iterations (a loss, a change-density map) alive across the capture: they pin
that iteration's autograd graph, whose AccumulateGrad nodes — created on the stream that iteration ran on — PyTorch then
re-uses inside the capture, and the cross-stream wait it inserts ("legacy stream depends on a capturing stream") invalidates
the capture.  Keep `loss.detach()` / `loss.item()` instead.
"""
from __future__ import annotations

import gc
import os
from typing import Callable, Sequence

import torch

from . import engine as E


def _snapshot(modules, optimizers):
    """Values of every parameter / buffer and of every optimizer-state tensor (warm-up iterations are REAL iterations)."""
    mods = [{k: v.detach().clone() for k, v in m.state_dict().items()} for m in modules]
    opts = [{p: {k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in st.items()} for p, st in o.state.items()}
            for o in optimizers]
    return mods, opts


def _restore(modules, optimizers, snap) -> None:
    """Write the snapshot back IN PLACE (the graphs hold the addresses); optimizer state that did not exist before the
    warm-up is zeroed, which is what a freshly constructed Adam / RMSprop state holds."""
    mods, opts = snap
    with torch.no_grad():
        for m, sd in zip(modules, mods):
            own = m.state_dict()
            for k, v in sd.items():
                own[k].copy_(v)
        for o, old in zip(optimizers, opts):
            for p, st in o.state.items():
                for k, v in st.items():
                    if torch.is_tensor(v):
                        v.copy_(old[p][k]) if p in old and k in old[p] else v.zero_()


def _drop_grads(modules) -> None:
    for m in modules:
        for p in m.parameters():
            p.grad = None


class GraphedStep:
    def __init__(self, fn: Callable, static_inputs: Sequence[torch.Tensor], warmup: int = 3,
                 capture_error_mode: str = "global", modules: Sequence[torch.nn.Module] = (),
                 optimizers: Sequence[torch.optim.Optimizer] = (), restore_after_warmup: bool = False):
        self.fn = fn
        self.static_inputs = list(static_inputs)
        snap = _snapshot(modules, optimizers) if restore_after_warmup else None
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn(*self.static_inputs)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        _drop_grads(modules)
        torch.cuda.empty_cache()        # the warm-up's blocks sit in the side stream's pool: hand them back before the capture
        gc.collect()                    # see YieldingStep: no autograd node of an earlier iteration may survive into the capture
        self.graph = torch.cuda.CUDAGraph()
        n0 = E.launch_count
        # "thread_local" lets other threads (e.g. the NCCL watchdog) keep making CUDA calls during the capture
        with torch.cuda.graph(self.graph, capture_error_mode=capture_error_mode):
            self.outputs = fn(*self.static_inputs)
        self.launches_per_replay = E.launch_count - n0      # libfcd_b200 C-ABI calls recorded in the graph
        if snap is not None:
            _restore(modules, optimizers, snap)
        E.bump_weight_epoch()

    def copy_inputs(self, *tensors: torch.Tensor) -> None:
        for dst, src in zip(self.static_inputs, tensors):
            dst.copy_(src, non_blocking=True)

    def __call__(self):
        self.graph.replay()
        E.bump_weight_epoch()      # parameters changed on the device without their version counters moving
        return self.outputs


class YieldingStep:
    """The multi-GPU launch path for the step bodies of steps.py: `genfn(*static_inputs)` is a generator that yields
    `(network, wait)` at its exchange points (steps.usss_gen / rsss_gen / wsss_gen ...).  The iteration is captured as one
    CUDA graph per stretch between exchange points (sharing one memory pool); at each point the gradient bucket of
    `network` is all-reduced by an eager NCCL call (`sync.launch`), and when `wait` is set the stream waits for every bucket
    in flight and the next graph starts by writing the averaged gradients back (`sync.unpack`).  Buckets launched with
    `wait=False` stay in flight while the following graph runs.

        step = YieldingStep(lambda x, y: steps.usss_gen(netG, netS, x, y, crit, optG, optS), sync, [x, y])
        out = step()                       # replay
    """

    def __init__(self, genfn: Callable, sync, static_inputs: Sequence[torch.Tensor], warmup: int = 3,
                 capture_error_mode: str = "thread_local", modules: Sequence[torch.nn.Module] = (),
                 optimizers: Sequence[torch.optim.Optimizer] = (), restore_after_warmup: bool = False):
        from .steps import drive

        self.sync = sync
        self.static_inputs = list(static_inputs)
        snap = _snapshot(modules, optimizers) if restore_after_warmup else None
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                drive(genfn(*self.static_inputs), sync.on_grads)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        _drop_grads(modules)            # see the module docstring: no warm-up gradient tensor may be written by the graphs
        torch.cuda.empty_cache()
        self.graphs, self.actions = [], []
        n0 = E.launch_count
        gen = genfn(*self.static_inputs)
        pool, unpack_next, done = None, False, False
        while not done:
            # autograd nodes of earlier (eager / warm-up) iterations that are still alive would be re-used by this capture with
            # the stream they were created on; a dead reference cycle is enough to keep them: collect before every capture
            gc.collect()
            g = torch.cuda.CUDAGraph()
            err = None
            try:
                with torch.cuda.graph(g, pool=pool, capture_error_mode=capture_error_mode):
                    try:
                        if unpack_next:
                            sync.unpack()
                            unpack_next = False
                        try:
                            net, wait = next(gen)
                            sync.pack(net)
                        except StopIteration as e:
                            self.outputs = e.value
                            done = True
                    except BaseException as e:      # keep the ROOT cause: capture_end() would mask it with "invalidated"
                        err = e
            except BaseException as e:
                err = err or e
            if err is not None:
                raise RuntimeError(f"YieldingStep: capture of graph {len(self.graphs)} failed: {type(err).__name__}: {err}") from err
            pool = g.pool()
            self.graphs.append(g)
            if not done:                    # keeps the backend's call sequence identical on every rank during set-up
                sync.launch(net)
                if wait:
                    sync.wait()
                    unpack_next = True
                self.actions.append((net, wait))
        self.launches_per_replay = E.launch_count - n0
        torch.cuda.synchronize()
        if snap is not None:
            _restore(modules, optimizers, snap)
        E.bump_weight_epoch()

    def copy_inputs(self, *tensors: torch.Tensor) -> None:
        for dst, src in zip(self.static_inputs, tensors):
            dst.copy_(src, non_blocking=True)

    def __call__(self):
        debug = bool(os.environ.get("FCD_GRAPH_DEBUG"))
        for i, g in enumerate(self.graphs):
            g.replay()
            if debug:                       # localise a faulting segment (FCD_GRAPH_DEBUG=1)
                try:
                    torch.cuda.synchronize()
                except Exception as e:
                    raise RuntimeError(f"YieldingStep: replay of graph {i} of {len(self.graphs)} faulted: {e}") from e
            if i < len(self.actions):
                net, wait = self.actions[i]
                self.sync.launch(net)
                if wait:
                    self.sync.wait()
        E.bump_weight_epoch()
        return self.outputs
