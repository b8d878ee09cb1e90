#!/usr/bin/env python
"""Self-play + value-network training, the reference's scripts/train.py:191-281 over the batched GPU
engine: cycles of  self-play -> dataset -> replay mix (all new + 30 % of old) -> train -> rotate
checkpoint,  with the reference's hyper-parameter schedule and CSV stats rows.

Multi-GPU (torchrun, one process per GPU): every rank self-plays its shard of the games with its own
engine (no collective), trains on its own positions, and the gradients are averaged with ONE NCCL
all-reduce per step (zeroclone_b200.training / parallel.GradSync -- the insertion point the reference leaves
between loss.backward() and optimizer.step(), models/chess_value/network.py:93-94).  Rank 0 writes
`latest.pth` (whole-module pickle, train.py:143) and rotates the previous one into checkpoints/.

  python scripts/train.py -c connect4_value --cycles 2 --games-cap 64 --sims-cap 64
  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 scripts/train.py -c connect4_value
"""
from __future__ import annotations

import argparse
import contextlib
import csv
import os
import sys
import time
from datetime import datetime
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
from zeroclone_b200 import parallel  # noqa: E402
from zeroclone_b200.engine import Engine  # noqa: E402
from zeroclone_b200.models import core  # noqa: E402
from zeroclone_b200.selfplay import DeviceSelfPlay  # noqa: E402
from zeroclone_b200.training import train_epochs  # noqa: E402

_REPLAY_STATES: list = []
_REPLAY_VALUES: list = []
_CSV = csv.writer(sys.stdout, lineterminator="\n")


def emit_stats(**kv):
    """one CSV row: timestamp + values sorted by key (train.py:84-89; parsed positionally by train_manager.py)"""
    _CSV.writerow([int(time.time())] + [kv[k] for k in sorted(kv)])
    sys.stdout.flush()


@contextlib.contextmanager
def timer(label):
    t0 = time.time()
    yield
    print(f"{label:<20} : {time.time() - t0:6.2f} s")


def update_replay(states_new, values_new, frac_old=0.30):
    """all fresh positions + a random 30 % of everything seen before (train.py:27-50)"""
    if not _REPLAY_STATES:
        _REPLAY_STATES.append(states_new)
        _REPLAY_VALUES.append(values_new)
        return states_new, values_new
    old_s, old_v = np.concatenate(_REPLAY_STATES), np.concatenate(_REPLAY_VALUES)
    k = int(frac_old * len(old_s))
    pick = np.random.choice(len(old_s), k, replace=False) if k > 0 else np.zeros(0, dtype=np.int64)
    _REPLAY_STATES.append(states_new)
    _REPLAY_VALUES.append(values_new)
    return np.concatenate([old_s[pick], states_new]), np.concatenate([old_v[pick], values_new])


def schedule_hyperparams(cycle, *, games_cap=2000, sims_cap=800, init_lr=3e-4, lr_decay=0.95, lr_floor=1e-5):
    """train.py:173-188"""
    return {"games": min(games_cap, 500 * (cycle + 1)), "simulations": int(min(100 * (1.2 ** cycle), sims_cap)),
            "c_puct": max(1.25, 2.5 * (0.97 ** cycle)), "lr": max(init_lr * (lr_decay ** cycle), lr_floor)}


def train_and_save_latest(model_type, model, states, values, *, epochs, lr, batch_size, device, rank, world):
    """Adam + MSE on (states, values) (network.py:75-101), data-parallel (zeroclone_b200.training.train_epochs: one
    flat NCCL all-reduce per step, no host sync inside the step, BatchNorm statistics averaged over ranks at the end);
    rank 0 then rotates the previous checkpoint and saves the new `latest.pth` (train.py:136-146)."""
    module, latest_path = core.get_value_network(model_type)
    stats = train_epochs(model, states, values, epochs=epochs, lr=lr, batch_size=batch_size, device=device, rank=rank,
                         timed=world > 1 and device.type == "cuda")
    if rank == 0:
        if "allreduce_ms" in stats:
            gbs = stats["allreduce_bytes"] / (stats["allreduce_ms_median"] * 1e-3) / 1e9
            print(f"train step {stats['step_ms']:.2f} ms; gradient all-reduce {stats['allreduce_bytes'] / 1e6:.2f} MB in "
                  f"{stats['allreduce_ms_median'] * 1e3:.0f} us (median of {stats['allreduce_calls']}) = {gbs:.0f} GB/s algorithmic")
        ckpt = Path(latest_path).parent / "checkpoints"
        ckpt.mkdir(parents=True, exist_ok=True)
        if Path(latest_path).exists():                      # rotate (train.py:136-141)
            Path(latest_path).rename(ckpt / f"{datetime.now().strftime('%Y%m%d_%H%M%S')}.pth")
        torch.save(model.to("cpu"), latest_path)
        model.to(device)
        print("Saved new *latest* model to", latest_path)
    return stats["loss"]


def full_training_run(config_name, *, cycles=30, batch_size=256, epochs=4, games_cap=2000, sims_cap=800, init_lr=3e-4,
                      lr_decay=0.95, lr_floor=1e-5):
    rank, local, world = parallel.env_rank()
    device = torch.device("cuda", local) if torch.cuda.is_available() else torch.device("cpu")
    if device.type == "cuda":
        torch.cuda.set_device(device)
    parallel.init(device=device if device.type == "cuda" else None)
    cfg_path = config_name if os.path.exists(config_name) else os.path.join(os.path.dirname(__file__), "..", "configs", f"{config_name}.yaml")
    engine = Engine(cfg_path)
    value = engine.values[0]
    if value.model is None:
        raise SystemExit("train.py needs a network evaluator (value_function: network_latest)")
    model_type = engine.config["value"]["model_type"]
    if rank == 0:
        emit_stats(stage="pid", pid=os.getpid())
    loss_acc = 0.0
    for cycle in range(cycles):
        engine.reset_all_games()
        hp = schedule_hyperparams(cycle, games_cap=games_cap, sims_cap=sims_cap, init_lr=init_lr, lr_decay=lr_decay, lr_floor=lr_floor)
        engine.config["mcts"]["simulations"], engine.config["mcts"]["c_puct"] = hp["simulations"], hp["c_puct"]
        my_games = len(parallel.shard_range(hp["games"], rank, world))
        if rank == 0:
            emit_stats(stage="cycle_start", cycle=cycle + 1, total_cycles=cycles, games_target=hp["games"], sims=hp["simulations"])
        with timer("SELF-PLAY"):          # device-resident: roots stay in HBM between moves, engine.threads games in flight
            sp = DeviceSelfPlay(engine.backend, value, engine.policy, n_slots=engine.threads, device=local if device.type == "cuda" else None,
                                batch_size=engine.batch_size, mode=engine.select_mode).play(my_games, hp["simulations"], hp["c_puct"], seed=cycle * 1000003 + rank)
            if rank == 0:
                emit_stats(stage="game_done", cycle=cycle + 1, finished=len(sp["results"]), target=my_games)
        games_h = parallel.reduce_sum([sp["games_per_hour"]])[0]
        states_now, values_now = sp["dataset"]
        states, values = update_replay(states_now, values_now)
        if rank == 0:
            print(f"Replay buffer : {len(values):,} positions total (+{len(values_now)} this cycle); self-play {games_h:,.0f} games/hour")
        with timer("TRAIN"):
            loss_acc += train_and_save_latest(model_type, value.model, states, values, epochs=epochs, lr=hp["lr"],
                                              batch_size=batch_size, device=device, rank=rank, world=world)
        value.model.eval()
        value.refresh()            # the resident tower takes the freshly trained weights (zc_tower_update_weights)
    if rank == 0:
        emit_stats(stage="train_done", cycle=cycles, loss=loss_acc / max(1, cycles), epochs=epochs)
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser(description="Self-play + value-net trainer on the batched GPU engine")
    ap.add_argument("-c", "--config", required=True, help="config name under ./configs/ (without .yaml) or a path")
    ap.add_argument("--cycles", type=int, default=30)
    ap.add_argument("--batch-size", type=int, default=256)
    ap.add_argument("--epochs", type=int, default=4)
    ap.add_argument("--num-workers", type=int, default=8, help="accepted for CLI compatibility; data is already on the device")
    ap.add_argument("--games-cap", type=int, default=2000)
    ap.add_argument("--sims-cap", type=int, default=800)
    ap.add_argument("--init-lr", type=float, default=3e-4)
    ap.add_argument("--lr-decay", type=float, default=0.95)
    ap.add_argument("--lr-floor", type=float, default=1e-5)
    a = ap.parse_args()
    full_training_run(a.config, cycles=a.cycles, batch_size=a.batch_size, epochs=a.epochs, games_cap=a.games_cap,
                      sims_cap=a.sims_cap, init_lr=a.init_lr, lr_decay=a.lr_decay, lr_floor=a.lr_floor)
