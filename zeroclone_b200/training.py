"""The value network's training step (models/chess_value/network.py:75-101 of the reference: Adam on the MSE
between the tower's output and the self-play label) as the data-parallel step of scripts/train.py.

Multi-GPU: every rank holds its own shard of positions on its device and runs the same number of steps; between
`loss.backward()` and `optimizer.step()` -- the place the reference leaves for it, network.py:93-94 -- the gradients
are averaged with ONE NCCL all-reduce of one flat bucket (parallel.GradSync).  Nothing inside the step synchronises
the host with the device: the loss is accumulated on the device and read once per epoch.
"""
from __future__ import annotations

import time
from typing import Optional

import numpy as np
import torch

from . import parallel


class _StepGraphs:
    """forward_backward: zero the bucket, gather rows `idx`, forward, MSE, backward, weight the gradients by the batch size and
    put the count into the bucket's last element;  update: divide by the all-reduced count, Adam step.  Warm-up runs the same
    operations eagerly on a side stream first (PyTorch's capture recipe) and is then undone."""

    def __init__(self, model, opt, sync, xs, ys, batch_size, device):
        self.idx = torch.arange(batch_size, device=device) % len(xs)
        self.loss = torch.zeros((), device=device)

        def fwd_bwd():
            sync.flat.zero_()
            loss = torch.nn.functional.mse_loss(model(xs[self.idx]), ys[self.idx])
            loss.backward()
            self.loss.copy_(loss.detach() * batch_size)
            sync.flat[:-1].mul_(float(batch_size))
            sync.flat[-1:].fill_(float(batch_size))

        def update():
            sync.flat[:-1].div_(sync.flat[-1].clamp_min(1.0))
            opt.step()

        tensors = list(model.parameters()) + list(model.buffers())
        snapshot = [t.detach().clone() for t in tensors]
        side = torch.cuda.Stream(device=device)
        side.wait_stream(torch.cuda.current_stream(device))
        with torch.cuda.stream(side):
            for _ in range(3):
                fwd_bwd()
                sync.reduce_bucket()
                update()
        torch.cuda.current_stream(device).wait_stream(side)
        # the warm-up steps must not count as training: parameters, BatchNorm statistics and Adam's moments go back
        with torch.no_grad():
            for t, t0 in zip(tensors, snapshot):
                t.copy_(t0)
            for st in opt.state.values():
                for v in st.values():
                    if torch.is_tensor(v):
                        v.zero_()
        sync.rebind()          # backward may have replaced a .grad during warm-up: make them views of the bucket again
        self.forward_backward, self.update = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.forward_backward):
            fwd_bwd()
        with torch.cuda.graph(self.update):
            update()


def train_epochs(model: torch.nn.Module, states: np.ndarray, values: np.ndarray, *, epochs: int, lr: float, batch_size: int,
                 device: torch.device, rank: int = 0, timed: bool = False, verbose: bool = True,
                 max_steps: Optional[int] = None, use_graphs: Optional[bool] = None) -> dict:
    """Train in place; returns {"loss": mean over epochs of the per-sample loss, "steps", "step_ms", and with
    timed=True "allreduce_ms" (mean device time of the collective), "allreduce_bytes"}."""
    model.to(device).train()
    sync = parallel.GradSync(model, timed=timed)
    sync.broadcast_parameters(0)
    opt = torch.optim.Adam(model.parameters(), lr=lr)
    xs = torch.from_numpy(np.ascontiguousarray(states)).float().to(device)
    ys = torch.from_numpy(np.ascontiguousarray(values)).float().to(device).unsqueeze(1)
    n = len(xs)
    steps = int(parallel.reduce_max([-(-n // batch_size)])[0])     # the longest shard's; shorter shards wrap around
    if max_steps is not None:
        steps = min(steps, max_steps)
    lane = torch.arange(batch_size, device=device)
    total, done_steps = 0.0, 0
    graphs = None
    if use_graphs is None:
        use_graphs = device.type == "cuda" and n > 0
    if use_graphs:
        # The step is launch-bound in eager mode (~150 small kernels for forward + backward, ~12 ms for 150 GFLOP): capture
        # [zero grads, gather the batch, forward, backward, weight the gradients] and [average, Adam update] as two CUDA
        # graphs around the one eager all-reduce.  Static inputs: the batch's row indices.
        opt = torch.optim.Adam(model.parameters(), lr=lr, capturable=True)
        graphs = _StepGraphs(model, opt, sync, xs, ys, batch_size, device)
    if device.type == "cuda":
        torch.cuda.synchronize(device)
    t0 = time.perf_counter()
    for epoch in range(1, epochs + 1):
        perm = torch.randperm(n, device=device) if n else None
        running = torch.zeros((), device=device)
        for s in range(steps):
            if graphs is not None:
                graphs.idx.copy_(perm[(lane + s * batch_size) % n])
                graphs.forward_backward.replay()
                running += graphs.loss
                sync.reduce_bucket()
                graphs.update.replay()
                done_steps += 1
                continue
            opt.zero_grad(set_to_none=False)            # gradients are views into the all-reduce bucket
            if n:
                idx = perm[(lane + s * batch_size) % n]
                loss = torch.nn.functional.mse_loss(model(xs[idx]), ys[idx])
                loss.backward()
                running += loss.detach() * batch_size
                sync(model, n_samples=batch_size)
            else:                                       # a rank without positions still takes part in the collective
                sync(model, n_samples=0)
            opt.step()
            done_steps += 1
        avg = float(running.item()) / max(1, steps * batch_size)      # the only host<->device sync of the epoch
        total += avg
        if rank == 0 and verbose:
            print(f"Epoch {epoch}/{epochs} — Loss: {avg:.4f}")
    if device.type == "cuda":
        torch.cuda.synchronize(device)
    dt = time.perf_counter() - t0
    sync.average_buffers()
    out = {"loss": total / max(1, epochs), "steps": done_steps, "cuda_graphs": graphs is not None, "step_ms": dt * 1e3 / max(1, done_steps), "positions": n,
           "allreduce_bytes": sync.bucket_bytes}
    if timed:
        ms = sync.allreduce_ms()
        if ms:
            ms_sorted = sorted(ms)
            out["allreduce_ms_list"] = ms
            out["allreduce_ms"] = sum(ms) / len(ms)
            out["allreduce_ms_median"] = ms_sorted[len(ms) // 2]
            out["allreduce_calls"] = len(ms)
    for p in model.parameters():          # drop the views: the bucket dies with `sync`
        p.grad = None
    return out
