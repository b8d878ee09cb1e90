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
    """Data-parallel gradient averaging for the value-network training step: ONE flat bucket, ONE all-reduce per
    step (NCCL over NVLink/NVSwitch on GPUs) -- 2,383,361 fp32 gradients + 1 sample count = 9.53 MB for the
    128x8 tower.  Every parameter's `.grad` is a view into the bucket, so nothing is copied in or out; each rank's
    gradient is weighted by its sample count (the count rides in the bucket's last element), so uneven self-play
    shards -- including empty ones -- average correctly.  Keep the views alive: `optimizer.zero_grad(set_to_none=False)`
    (or `sync.zero_()`).  `timed=True` brackets the collective with CUDA events; `allreduce_ms()` reads them back."""

    def __init__(self, model: torch.nn.Module, timed: bool = False):
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.buffers = [b for b in model.buffers() if b.dtype.is_floating_point]
        dev, dt = self.params[0].device, self.params[0].dtype
        self.numel = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(self.numel + 1, device=dev, dtype=dt)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()
        self.timed = timed and dev.type == "cuda"
        self._events: list = []
        self.calls = 0

    @property
    def bucket_bytes(self) -> int:
        return self.flat.numel() * self.flat.element_size()

    def zero_(self) -> None:
        self.flat.zero_()

    def broadcast_parameters(self, src: int = 0) -> None:
        """rank `src`'s parameters and BatchNorm statistics to every rank, one flat broadcast"""
        if not _active():
            return
        tensors = [p.data for p in self.params] + [b.data for b in self.buffers]
        flat = torch._utils._flatten_dense_tensors(tensors)
        dist.broadcast(flat, src)
        for t, f in zip(tensors, torch._utils._unflatten_dense_tensors(flat, tensors)):
            t.copy_(f)

    def average_buffers(self) -> None:
        """BatchNorm running_mean / running_var are updated from each rank's own shard during training; average
        them so every rank folds the SAME network into its self-play evaluator (one flat all-reduce)."""
        if not _active() or not self.buffers:
            return
        flat = torch._utils._flatten_dense_tensors([b.data for b in self.buffers])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(dist.get_world_size())
        for b, f in zip(self.buffers, torch._utils._unflatten_dense_tensors(flat, self.buffers)):
            b.data.copy_(f)

    def __call__(self, model=None, n_samples: int | None = None) -> None:
        """between loss.backward() and optimizer.step() (network.py:93-94)"""
        if not _active():
            return
        for p in self.params:                      # a grad re-created by zero_grad(set_to_none=True) is folded back
            if p.grad is None or p.grad.data_ptr() < self.flat.data_ptr() or p.grad.data_ptr() >= self.flat.data_ptr() + self.bucket_bytes:
                self._rebind()
                break
        w = 1.0 if n_samples is None else float(n_samples)      # unweighted: every rank counts once
        if n_samples is not None:
            self.flat[:-1].mul_(w)
        self.flat[-1] = w
        if self.timed:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
        if self.timed:
            b.record()
            self._events.append((a, b))
        self.flat[:-1].div_(self.flat[-1].clamp_min(1.0))
        self.calls += 1

    def reduce_bucket(self) -> None:
        """The collective alone, for a caller that fills the bucket itself (training.py's CUDA-graph step): gradients already
        weighted by the sample count, the count in the last element; the caller divides afterwards."""
        if not _active():
            return
        if self.timed:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
        if self.timed:
            b.record()
            self._events.append((a, b))
        self.calls += 1

    def rebind(self) -> None:
        self._rebind()

    def _rebind(self) -> None:
        off = 0
        for p in self.params:
            view = self.flat[off:off + p.numel()].view_as(p)
            if p.grad is None:
                view.zero_()
            elif p.grad.data_ptr() != view.data_ptr():
                view.copy_(p.grad)
            p.grad = view
            off += p.numel()

    def allreduce_ms(self) -> List[float]:
        """device time of every timed all-reduce so far (synchronises)"""
        if not self._events:
            return []
        self._events[-1][1].synchronize()
        out = [a.elapsed_time(b) for a, b in self._events]
        self._events = []
        return out
