"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: shard ranges, timing/counter
reductions, result gathering and the gradient all-reduce of the training step."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from zeroclone_b200 import parallel


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank),
                      WORLD_SIZE=str(world))
    r, w = parallel.init("gloo")
    assert (r, w) == (rank, world)
    # shards tile the id space
    mine = parallel.shard_range(4099, r, w)
    all_shards = parallel.gather_objects((mine.start, mine.stop))
    # timings: max over ranks; counters: sum over ranks
    tmax = parallel.reduce_max([10.0 + r, 5.0 - r])
    tsum = parallel.reduce_sum([len(mine)])
    # gradient all-reduce == the gradient of the concatenated batch (equal shard sizes)
    from zeroclone_b200.models.connect4_value.network import ValueNetwork
    torch.manual_seed(0)
    model = ValueNetwork(channels=8, blocks=1)
    sync = parallel.GradSync(model)
    sync.broadcast_parameters(0)
    g = torch.Generator().manual_seed(123)
    x = torch.randn(8, 2, 6, 7, generator=g)
    y = torch.randn(8, 1, generator=g)
    model.train()
    xs, ys = x[r * 4:(r + 1) * 4], y[r * 4:(r + 1) * 4]
    torch.nn.functional.mse_loss(model(xs), ys).backward()
    sync(model)
    flat = torch.cat([p.grad.view(-1) for p in model.parameters()])
    # weighted variant with uneven sample counts
    model.zero_grad()
    n_r = 2 if r == 0 else 6
    lo = 0 if r == 0 else 2
    torch.nn.functional.mse_loss(model(x[lo:lo + n_r]), y[lo:lo + n_r]).backward()
    sync(model, n_samples=n_r)
    flat_w = torch.cat([p.grad.view(-1) for p in model.parameters()])
    assert sync.calls == 2                     # one collective per step
    # the training step of scripts/train.py with uneven shards, one of them EMPTY: same number of collectives on
    # both ranks, identical parameters and BatchNorm statistics afterwards
    from zeroclone_b200.training import train_epochs
    n_mine = 0 if r == 0 else 40
    xs_t = torch.randn(n_mine, 2, 6, 7, generator=g).numpy()
    ys_t = torch.randn(n_mine, generator=g).numpy()
    st = train_epochs(model, xs_t, ys_t, epochs=2, lr=1e-3, batch_size=16, device=torch.device("cpu"), rank=r, verbose=False)
    assert st["steps"] == 2 * 3 and st["positions"] == n_mine
    state = torch.cat([t.detach().float().view(-1) for t in list(model.parameters()) + [b for b in model.buffers() if b.dtype.is_floating_point]])
    q.put((rank, all_shards, tmax, tsum, flat.numpy(), flat_w.numpy(), state.numpy()))
    dist.destroy_process_group()


def test_world_size_2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=180) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, shards0, tmax0, tsum0, g0, gw0, st0), (_, shards1, tmax1, tsum1, g1, gw1, st1) = got
    assert np.array_equal(st0, st1) and np.isfinite(st0).all()      # ranks end the training step with the same network
    assert shards0 == shards1 == [(0, 2050), (2050, 4099)]
    assert tmax0 == tmax1 == [11.0, 5.0] and tsum0 == tsum1 == [4099.0]
    assert np.array_equal(g0, g1) and np.array_equal(gw0, gw1)
    # single-process reference: BatchNorm in train mode uses per-rank batch statistics, so compare against
    # the average of the two per-shard gradients computed locally
    from zeroclone_b200.models.connect4_value.network import ValueNetwork
    torch.manual_seed(0)
    model = ValueNetwork(channels=8, blocks=1)
    g = torch.Generator().manual_seed(123)
    x = torch.randn(8, 2, 6, 7, generator=g)
    y = torch.randn(8, 1, generator=g)
    model.train()

    def grad_of(lo, hi):
        model.zero_grad()
        torch.nn.functional.mse_loss(model(x[lo:hi]), y[lo:hi]).backward()
        return torch.cat([p.grad.view(-1) for p in model.parameters()]).numpy().copy()

    want = (grad_of(0, 4) + grad_of(4, 8)) / 2
    assert np.allclose(g0, want, atol=1e-6)
    want_w = (2 * grad_of(0, 2) + 6 * grad_of(2, 8)) / 8
    assert np.allclose(gw0, want_w, atol=1e-6)


def test_shard_range_properties():
    for n in (1, 7, 4096, 32768, 16385):
        for world in (1, 2, 4, 8):
            parts = [parallel.shard_range(n, r, world) for r in range(world)]
            assert parts[0].start == 0 and parts[-1].stop == n
            assert all(a.stop == b.start for a, b in zip(parts, parts[1:]))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
