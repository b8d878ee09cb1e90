"""Self-play driver: the reference's `simulate_games` loop (scripts/train.py:151-170) over the
batched engine -- every iteration is ONE device search over all unfinished games, one move per
game; finished games are replaced by fresh ones until `total_games` have been started.
Reports games/hour, the second half of BASELINE.json's metric."""
from __future__ import annotations

import time
from typing import Callable, List, Optional


def simulate_games(engine, total_games: int, simulations: Optional[int] = None, c: Optional[float] = None,
                   on_game_done: Optional[Callable[[int, int], None]] = None) -> dict:
    """Play `total_games` games with at most `engine.threads` in flight.
    Returns {"results": [...], "moves": plies played, "seconds": wall, "games_per_hour", "sims_per_sec"}."""
    sims = simulations if simulations is not None else engine.config["mcts"]["simulations"]
    c_val = c if c is not None else engine.config["mcts"]["c_puct"]
    first = min(total_games, engine.threads)
    unfinished = set(range(first))
    final: List[Optional[int]] = [None] * first
    plies = searched = 0
    t0 = time.time()
    while unfinished:
        batch = sorted(unfinished)
        results = engine.play_mcts_parallel(batch, simulations=sims, c=c_val)
        plies += len(batch)
        searched += sum(1 for i in batch if i in engine.last_search)
        for idx in batch:
            if results[idx] is None:
                continue
            unfinished.discard(idx)
            final[idx] = results[idx]
            if on_game_done is not None:
                on_game_done(sum(r is not None for r in final), total_games)
            if len(final) < total_games:            # refill: a new game takes a new index (train.py:166-168)
                final.append(None)
                unfinished.add(engine.add_game())
    dt = time.time() - t0
    return {"results": final, "moves": plies, "seconds": dt, "games_per_hour": len(final) / dt * 3600.0,
            "sims_per_sec": plies * sims / dt}


class DeviceSelfPlay:
    """Device-resident self-play: root states live in HBM between moves.

    Per ply: gather the running games' roots (device) -> one batched search -> `zc_search_advance`
    applies every tree's chosen move and evaluates the new position on the device (win / draw incl.
    chess 50-ply counter and KMP repetition) -> scatter back.  The host only receives (move, result)
    per game and a copy of the packed states for the dataset.  Refill semantics of
    scripts/train.py:151-170: `n_slots` games in flight until `total_games` have been started.
    Labels as Engine.get_dataset (engine.py:60-89): the final position -1, alternating backwards, draws 0.
    """

    def __init__(self, backend, value, policy, n_slots: int, device: int | None = None, batch_size: int = 32,
                 hist_cap: int = 512, mode: tuple | None = None):
        import torch

        from . import _ffi, mcts
        self.torch, self._ffi, self._mcts = torch, _ffi, mcts
        self.backend, self.value, self.policy = backend, value, policy
        self.game = backend.ZC_GAME
        self.n_slots, self.batch_size, self.hist_cap = n_slots, batch_size, hist_cap
        self.mode = mode or (_ffi.SELECT_UCB1, 1.0, 0)        # mcts.select_mode(config["mcts"]); default = the reference's UCB1
        self.device = torch.cuda.current_device() if device is None else device
        self.kind, self.ev = value.device_spec(self.game)
        self.pol = policy.device_policy
        self.pol_freedom = getattr(policy, "device_freedom", 0.0)
        self.init_rec = self._pack(backend.create_init_state())

    def _pack(self, state):
        import numpy as np
        rec = np.zeros(1, dtype=self.backend.STATE_DTYPE)
        rec[0] = self.backend.pack_state(state)
        return rec

    def play(self, total_games: int, simulations: int, c: float = 1.4, seed: int = 0, record: bool = True) -> dict:
        import ctypes as C

        import numpy as np
        torch, _ffi = self.torch, self._ffi
        dev = torch.device("cuda", self.device)
        if simulations < 1:
            raise ValueError("DeviceSelfPlay.play: simulations must be >= 1 (a search that never ran has no move to play)")
        item = self.init_rec.dtype.itemsize
        chess = self.game == _ffi.GAME_CHESS
        if total_games <= 0:              # a rank whose shard of the games is empty (games_cap < world size)
            out = {"results": [], "moves": 0, "seconds": 0.0, "games_per_hour": 0.0, "sims_per_sec": 0.0}
            if record:
                out["dataset"] = (np.zeros((0,) + tuple(self.backend.TENSOR_SHAPE), dtype=np.float32), np.zeros(0, dtype=np.float32))
                out["trajectories"] = []
            return out
        n_slots = min(self.n_slots, total_games)
        hist_cap = self.hist_cap
        with torch.cuda.device(dev), self._mcts.device_lock(self.game, self.device):
            stream = torch.cuda.current_stream().cuda_stream
            states_all = torch.from_numpy(np.repeat(self.init_rec, n_slots).view(np.uint8).reshape(n_slots, item).copy()).to(dev)
            hist_all = torch.zeros((n_slots, 2, hist_cap, 8), dtype=torch.uint8, device=dev) if chess else None
            slot_plies = np.zeros(n_slots, dtype=np.int64)      # plies of the game currently in each slot
            hlen_all = torch.zeros((n_slots, 2), dtype=torch.int32, device=dev) if chess else None
            ts = self._mcts.searcher(self.game, n_slots, simulations, self.device)
            ts.set_policy_freedom(self.pol_freedom)
            ts.set_mode(*self.mode)
            slot_game = list(range(n_slots))              # game id played in each slot
            started, finished = n_slots, 0
            results = [None] * total_games
            traj = [[self.init_rec[0].copy()] for _ in range(total_games)] if record else None
            active = np.arange(n_slots, dtype=np.int64)
            ones_all = torch.ones(n_slots, dtype=torch.uint8, device=dev)
            init_dev = torch.from_numpy(self.init_rec.view(np.uint8).reshape(1, item).copy()).to(dev)
            plies = 0
            t0 = time.time()
            while len(active):
                n = len(active)
                # every slot busy (the steady state while games are being refilled): work on the slot arrays in place;
                # otherwise gather the running games' rows, and scatter them back after the move
                whole = n == n_slots
                idx = None if whole else torch.from_numpy(active).to(dev)
                roots = states_all if whole else states_all.index_select(0, idx).contiguous()
                ts.set_roots_dev(roots.data_ptr(), n, stream)
                if self.kind == "builtin":
                    ts.run(simulations, c, self.batch_size, self.ev, self.pol, seed + plies, stream)
                else:
                    ts.run_network(self.ev, simulations, c, self.batch_size, self.pol, seed + plies)
                ones = ones_all[:n]
                res = np.zeros(n, dtype=np.int32)
                mv = np.zeros(n, dtype=_ffi.CHESS_MOVE_DTYPE)
                if chess:
                    if (int(slot_plies[active].max()) + 1) // 2 + 1 >= hist_cap:
                        # a side's move list is about to fill up: double the history (the reference's deques are
                        # unbounded, state.h:14); a full history is an error in zc_search_advance, never a silent drop
                        grown = torch.zeros((n_slots, 2, 2 * hist_cap, 8), dtype=torch.uint8, device=dev)
                        grown[:, :, :hist_cap] = hist_all
                        hist_all, hist_cap = grown, 2 * hist_cap
                    hist = hist_all if whole else hist_all.index_select(0, idx).contiguous()
                    hlen = hlen_all if whole else hlen_all.index_select(0, idx).contiguous()
                    _ffi.check(_ffi.lib().zc_search_advance(ts._h, roots.data_ptr(), ones.data_ptr(), hist.data_ptr(), hlen.data_ptr(),
                                                           hist_cap, res.ctypes.data_as(C.c_void_p), mv.ctypes.data_as(C.c_void_p), stream))
                    if not whole:
                        hist_all.index_copy_(0, idx, hist)
                        hlen_all.index_copy_(0, idx, hlen)
                else:
                    _ffi.check(_ffi.lib().zc_search_advance(ts._h, roots.data_ptr(), ones.data_ptr(), None, None, 0,
                                                           res.ctypes.data_as(C.c_void_p), mv.ctypes.data_as(C.c_void_p), stream))
                plies += n
                slot_plies[active] += 1
                new_host = roots.cpu().numpy().view(self.init_rec.dtype).reshape(n) if record else None
                # bookkeeping per ply, vectorised: which games ended, which slots get the next game (refill
                # semantics of simulate_games, scripts/train.py:151-170)
                act = active
                if record:
                    for j, slot in enumerate(active.tolist()):
                        traj[slot_game[slot]].append(new_host[j].copy())
                done = np.nonzero(res != _ffi.RESULT_ONGOING)[0]
                keep = np.ones(n, dtype=bool)
                refill = []
                for j in done:                       # a handful per ply
                    slot = int(act[j])
                    results[slot_game[slot]] = int(res[j])
                    finished += 1
                    if started < total_games:
                        slot_game[slot] = started
                        slot_plies[slot] = 0
                        started += 1
                        refill.append(int(j))
                    else:
                        keep[j] = False
                if not whole:
                    states_all.index_copy_(0, idx, roots)
                if refill:
                    ridx = torch.from_numpy(act[np.asarray(refill, dtype=np.int64)]).to(dev)
                    states_all[ridx] = init_dev
                    if chess:
                        hlen_all[ridx] = 0
                active = act[keep]
            torch.cuda.synchronize()
            dt = time.time() - t0
        out = {"results": results, "moves": plies, "seconds": dt, "games_per_hour": total_games / dt * 3600.0,
               "sims_per_sec": plies * simulations / dt}
        if record:
            lens = [len(t) for t in traj]
            flat = np.concatenate([np.stack(t) for t in traj]) if traj else np.zeros(0, dtype=self.init_rec.dtype)
            planes = np.zeros((len(flat),) + tuple(self.backend.TENSOR_SHAPE), dtype=np.float32)
            _ffi.check(_ffi.lib().zc_states_to_tensor(self.game, flat.ctypes.data_as(C.c_void_p), len(flat),
                                                      planes.ctypes.data_as(C.c_void_p)))
            labels = np.concatenate([(0.0 if r == 0 else -1.0) * (-1.0) ** (np.arange(L)[::-1]) for r, L in zip(results, lens)])
            out["dataset"] = (planes, labels.astype(np.float32))
            out["trajectories"] = traj
        return out
