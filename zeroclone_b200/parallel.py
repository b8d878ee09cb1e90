"""Multi-GPU plumbing: one process per GPU (torchrun), trees sharded by contiguous id ranges.

The search needs NO collective: every tree touches only its own arena (the reference treats games
independently too, engine.py:131-138).  torch.distributed is used for (a) agreeing on timings and
counters across ranks (max / sum), (b) gathering per-rank results, and (c) the one real exchange
step of the whole system: the gradient all-reduce of the training step (the insertion point is
between loss.backward() and optimizer.step(), reference models/chess_value/network.py:93-94).
Works with backend "nccl" on GPUs and "gloo" on CPU (tests).
"""
from __future__ import annotations

import os
from typing import Iterable, List, Tuple

import torch
import torch.distributed as dist


def env_rank() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment; (0, 0, 1) when not launched by it."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init(backend: str | None = None, device: torch.device | None = None) -> Tuple[int, int]:
    """Join the process group named by the environment (MASTER_ADDR/PORT, RANK, WORLD_SIZE)."""
    rank, local, world = env_rank()
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, **kw)
    return rank, world


def shard_range(n_total: int, rank: int, world: int) -> range:
    """Contiguous block of tree/game ids owned by `rank`; sizes differ by at most one."""
    base, extra = divmod(n_total, world)
    start = rank * base + min(rank, extra)
    return range(start, start + base + (1 if rank < extra else 0))


def _active() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def _reduce_device(device):
    """NCCL reduces device tensors only; gloo reduces host tensors."""
    if device is not None:
        return device
    if _active() and dist.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return "cpu"


def reduce_max(values: Iterable[float], device=None) -> List[float]:
    t = torch.tensor(list(values), dtype=torch.float64, device=_reduce_device(device))
    if _active():
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def reduce_sum(values: Iterable[float], device=None) -> List[float]:
    t = torch.tensor(list(values), dtype=torch.float64, device=_reduce_device(device))
    if _active():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.tolist()


def gather_objects(obj):
    """Every rank's object, in rank order, on every rank."""
    if not _active():
        return [obj]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out


class GradSync:
    """Data-parallel gradient averaging for the value-network training step: one flat bucket
    (2,383,361 fp32 gradients = 9.53 MB for the 128x8 tower), one all-reduce (NCCL over
    NVLink/NVSwitch on GPUs), then divide by the world size.  Optionally weights each rank's
    gradient by its sample count so that uneven self-play shards average correctly."""

    def __init__(self, model: torch.nn.Module):
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.flat = None

    def broadcast_parameters(self, src: int = 0) -> None:
        if _active():
            for p in self.params:
                dist.broadcast(p.data, src)

    def __call__(self, model=None, n_samples: int | None = None) -> None:
        if not _active():
            return
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in self.params]
        flat = torch._utils._flatten_dense_tensors(grads)
        if n_samples is None:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
            flat.div_(dist.get_world_size())
        else:
            w = torch.tensor([float(n_samples)], device=flat.device, dtype=flat.dtype)
            flat.mul_(w)
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
            dist.all_reduce(w, op=dist.ReduceOp.SUM)
            flat.div_(w.clamp_min(1.0))
        for p, g in zip(self.params, torch._utils._unflatten_dense_tensors(flat, grads)):
            p.grad = g.contiguous() if p.grad is None else p.grad.copy_(g)
