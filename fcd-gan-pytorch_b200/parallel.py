"""Data-parallel layer (no reference counterpart — the reference is single-device).

Tile pairs are independent samples, so the batch is sharded over ranks (one process per GPU, `torchrun`) and the
only exchange step of an iteration is the all-reduce (mean) of each network's gradients over NCCL / NVLink.
Each network backward is ONE autograd node here, so its gradients become available together; `GradSync` packs
them into one flat fp32 bucket per network (G 2.2 MB, D 8.3 MB, S 163 MB at 13 bands) and launches the all-reduce
asynchronously, so it overlaps with whatever the step does next (e.g. the discriminator pass while the
generator's bucket is in flight); `finish()` waits, scales by 1/world and scatters the result back into `.grad`.
BatchNorm statistics stay per rank (standard DDP semantics).

Works on any torch.distributed backend: NCCL on the GPU box, gloo in the CPU unit tests (tests/test_parallel.py).
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None) -> int:
    """Initialise the default process group from torchrun's environment; returns the local rank."""
    import os

    if not dist.is_initialized():
        import datetime

        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        # a mismatched collective should fail in minutes, not hold a GPU box for the default 10
        dist.init_process_group(backend=backend, timeout=datetime.timedelta(seconds=int(os.environ.get("FCD_DIST_TIMEOUT_S", "180"))))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if torch.cuda.is_available():
        torch.cuda.set_device(local)
    return local


def shard_batch(n_items: int, rank: int, world: int) -> slice:
    """Contiguous, near-equal shard of a global batch of tile pairs for `rank`."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return slice(start, start + base + (1 if rank < rem else 0))


def broadcast_parameters(modules: Iterable[torch.nn.Module], src: int = 0, group=None) -> None:
    """Make every rank start from rank `src`'s parameters and buffers (BatchNorm running stats included)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    for m in modules:
        for t in list(m.parameters()) + list(m.buffers()):
            dist.broadcast(t.data, src=src, group=group)
    # the broadcast wrote through `.data`: version counters did not move, so packed conv weights made by an earlier forward
    # would be stale (engine._packed)
    from . import engine

    engine.invalidate_weight_cache()


class GradSync:
    """Bucketed asynchronous gradient all-reduce, one bucket per network."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self._pending: List = []
        self._buckets: Dict[int, torch.Tensor] = {}
        self._packed: Dict[int, tuple] = {}

    # ---- the four phases; pack / unpack are plain device work (CUDA-graph capturable), launch / wait talk to the backend ----
    def pack(self, module: torch.nn.Module) -> None:
        """Gather `module`'s gradients into its flat bucket (call right after its backward)."""
        if self.world == 1:
            return
        params = [p for p in module.parameters() if p.grad is not None]
        if not params:
            return
        n = sum(p.grad.numel() for p in params)
        flat = self._buckets.get(id(module))
        if flat is None or flat.numel() != n or flat.device != params[0].grad.device:
            flat = torch.empty(n, dtype=torch.float32, device=params[0].grad.device)
            self._buckets[id(module)] = flat
        torch.cat([p.grad.reshape(-1) for p in params], out=flat)
        self._packed[id(module)] = (flat, params)

    def launch(self, module: torch.nn.Module) -> None:
        """Start the asynchronous all-reduce of the packed bucket of `module`."""
        flat = self._buckets.get(id(module))      # the bucket outlives pack(): under CUDA-graph replay pack() is not re-run
        if self.world == 1 or flat is None:
            return
        self._pending.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def wait(self) -> None:
        """Make the current stream wait for every all-reduce in flight."""
        for work in self._pending:
            work.wait()
        self._pending.clear()

    def unpack(self) -> None:
        """Scale the reduced buckets by 1/world and write them back into `.grad`."""
        for flat, params in self._packed.values():
            flat.mul_(1.0 / self.world)
            views, off = [], 0
            for p in params:
                k = p.grad.numel()
                views.append(flat[off:off + k].view_as(p.grad))
                off += k
            torch._foreach_copy_([p.grad for p in params], views)      # one multi-tensor launch instead of one per parameter
        self._packed.clear()

    # ---- eager composition ------------------------------------------------------------------------------------------
    def start(self, module: torch.nn.Module) -> None:
        """Launch the all-reduce of `module`'s gradients (call right after its backward)."""
        self.pack(module)
        self.launch(module)

    def finish(self) -> None:
        """Wait for every bucket in flight and write the averaged gradients back."""
        self.wait()
        self.unpack()

    def on_grads(self, module: torch.nn.Module, wait: bool) -> None:
        """Exchange-point callback of the step bodies (steps.py): `module`'s gradients are complete; `wait` = the next
        statement is an optimizer step that needs every bucket in flight."""
        self.start(module)
        if wait:
            self.finish()
