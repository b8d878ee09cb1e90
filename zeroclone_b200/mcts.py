"""`get_move` with the reference's signature (engine/mcts/src/bindings_mcts.cpp:9-11,
engine/mcts/__init__.py) running on the GPU, and the shared pool of search handles.

The reference's get_move is generic over duck-typed Python callbacks.  This one recognises the
built-in backends / Value / Policy objects of this package and runs the CUDA search; anything else
raises TypeError -- there is deliberately no CPU search to fall back to.
"""
from __future__ import annotations

import threading
from typing import Dict, Sequence, Tuple

import numpy as np

from . import _ffi
from .search import TreeSearch

_POOL: Dict[Tuple[int, int], TreeSearch] = {}
_POOL_LOCK = threading.RLock()
_KEY_LOCKS: Dict[Tuple[int, int], threading.RLock] = {}


def _key(game: int, device: int | None) -> Tuple[int, int]:
    if device is None:
        import torch
        device = torch.cuda.current_device() if torch.cuda.is_available() else 0
    return game, device


def device_lock(game: int, device: int | None = None) -> threading.RLock:
    """One lock per pooled handle.  A search is set_roots -> run -> results on ONE shared handle (and one shared
    network evaluator); callers that may run concurrently -- the REST server's threadpool, ctypes calls release the
    GIL -- hold this lock for the whole sequence, or a game would receive a move computed for another game's root.
    (The reference built a private tree per request, mcts.cpp:104-109.)"""
    key = _key(game, device)
    with _POOL_LOCK:
        return _KEY_LOCKS.setdefault(key, threading.RLock())


def searcher(game: int, n_trees: int, sims: int, device: int | None = None) -> TreeSearch:
    """A TreeSearch handle with room for n_trees x sims on `device`; capacities only ever grow (trees
    geometrically).  A failed growth leaves the pool consistent: the old handle stays if the new one cannot be
    created next to it; if memory is the problem the old one is released first and, should the retry fail too,
    the pool entry is dropped so the next call starts clean."""
    key = _key(game, device)
    with _POOL_LOCK:
        ts = _POOL.get(key)
        if ts is not None and ts.max_trees >= n_trees and ts.max_sims >= sims:
            return ts
        if ts is None:
            cap_t = max(n_trees, 1)
        elif n_trees > ts.max_trees:
            cap_t = max(n_trees, 2 * ts.max_trees)
        else:
            cap_t = ts.max_trees                 # only `sims` grew: keep the tree capacity
        cap_s = max(sims, ts.max_sims if ts else 0, 32)
        try:
            new = TreeSearch(game, cap_t, cap_s, device=key[1])
        except _ffi.ZcError:
            if ts is None:
                raise
            ts.close()                       # make room and try once more, at the size actually asked for
            del _POOL[key]
            new = TreeSearch(game, max(n_trees, 1), max(sims, 32), device=key[1])
        else:
            if ts is not None:
                ts.close()
        _POOL[key] = new
        return new


def release_all() -> None:
    with _POOL_LOCK:
        for ts in _POOL.values():
            ts.close()
        _POOL.clear()


def _device_pieces(value, policy, backend):
    game = getattr(backend, "ZC_GAME", None)
    if game is None or not hasattr(backend, "pack_state"):
        raise TypeError("get_move: unknown backend; the CUDA search supports zeroclone_b200.games.* only "
                        "(no CPU search exists to run arbitrary Python callbacks)")
    if not hasattr(value, "device_spec"):
        raise TypeError("get_move: value must be a zeroclone_b200.value_functions.Value")
    if not hasattr(policy, "device_policy"):
        raise TypeError("get_move: policy must be a zeroclone_b200.policy_functions.Policy")
    return game, value.device_spec(game), policy.device_policy


def select_mode(cfg: dict | None) -> tuple:
    """(select, virtual_loss, prior_weight) from the optional `mcts:` keys `select: ucb1|puct`, `virtual_loss`, `prior_weight`
    (extra keys; the reference's YAMLs have none of them and get its UCB1)."""
    cfg = cfg or {}
    name = str(cfg.get("select", "ucb1")).lower()
    if name not in ("ucb1", "puct"):
        raise ValueError(f"mcts.select must be 'ucb1' or 'puct', not {name!r}")
    return (_ffi.SELECT_PUCT if name == "puct" else _ffi.SELECT_UCB1, float(cfg.get("virtual_loss", 1.0)), int(cfg.get("prior_weight", 0)))


def search_batch(states: Sequence, value, policy, backend, simulations: int, c: float, batch_size: int = 32,
                 seed: int | None = None, stats: bool = False, mode: tuple | None = None) -> dict:
    """Search all `states` at once; returns TreeSearch.results() plus 'moves_out' (backend move objects).
    mode: select_mode(...) tuple; None = the reference's UCB1."""
    game, (kind, ev), pol = _device_pieces(value, policy, backend)
    n = len(states)
    roots = np.zeros(n, dtype=backend.STATE_DTYPE)
    for i, s in enumerate(states):
        roots[i] = backend.pack_state(s)
    if seed is None:
        seed = int(np.random.SeedSequence().generate_state(1, dtype=np.uint64)[0])
    if simulations < 1:
        raise ValueError("simulations must be >= 1")
    with device_lock(game):
        ts = searcher(game, n, simulations)
        ts.set_mode(*(mode or (_ffi.SELECT_UCB1, 1.0, 0)))
        ts.set_roots(roots)
        ts.set_policy_freedom(getattr(policy, "device_freedom", 0.0))
        if kind == "builtin":
            ts.run(simulations, c, batch_size, ev, pol, seed)
        else:
            ts.run_network(ev, simulations, c, batch_size, pol, seed)
        out = ts.results(stats=stats)
    res = out["result"]
    if game == _ffi.GAME_C4:
        out["moves_out"] = [backend.move_from_result(r["best_move"]) if r["best"] >= 0 else None for r in res]
    else:
        out["moves_out"] = [backend.move_from_result(r["best_move"], r["best_move_value"]) if r["best"] >= 0 else None
                            for r in res]
    return out


def get_move(state, value, policy, backend, simulations=1000, c=1.4, batch_size=32):
    """One tree: returns the most-visited root move (mcts.cpp:150-159), lowest index on ties."""
    out = search_batch([state], value, policy, backend, int(simulations), float(c), int(batch_size))
    mv = out["moves_out"][0]
    if mv is None:
        raise ValueError("get_move: the root has no legal move or no simulation ran")   # UB in the reference (mcts.cpp:156)
    return mv
