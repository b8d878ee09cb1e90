"""`get_move` with the reference's signature (engine/mcts/src/bindings_mcts.cpp:9-11,
engine/mcts/__init__.py) running on the GPU, and the shared pool of search handles.

The reference's get_move is generic over duck-typed Python callbacks.  This one recognises the
built-in backends / Value / Policy objects of this package and runs the CUDA search; anything else
raises TypeError -- there is deliberately no CPU search to fall back to.
"""
from __future__ import annotations

from typing import Dict, Sequence, Tuple

import numpy as np

from . import _ffi
from .search import TreeSearch

_POOL: Dict[Tuple[int, int], TreeSearch] = {}


def searcher(game: int, n_trees: int, sims: int, device: int | None = None) -> TreeSearch:
    """A TreeSearch handle with room for n_trees x sims on `device`, grown geometrically."""
    if device is None:
        import torch
        device = torch.cuda.current_device() if torch.cuda.is_available() else 0
    key = (game, device)
    ts = _POOL.get(key)
    if ts is None or ts.max_trees < n_trees or ts.max_sims < sims:
        cap_t = max(n_trees, 2 * ts.max_trees if ts and ts.max_trees < n_trees else 0, 1)
        cap_s = max(sims, ts.max_sims if ts else 0, 32)
        if ts is not None:
            ts.close()
        ts = TreeSearch(game, cap_t, cap_s, device=device)
        _POOL[key] = ts
    return ts


def release_all() -> None:
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


def search_batch(states: Sequence, value, policy, backend, simulations: int, c: float, batch_size: int = 32,
                 seed: int | None = None, stats: bool = False) -> dict:
    """Search all `states` at once; returns TreeSearch.results() plus 'moves_out' (backend move objects)."""
    game, (kind, ev), pol = _device_pieces(value, policy, backend)
    n = len(states)
    roots = np.zeros(n, dtype=backend.STATE_DTYPE)
    for i, s in enumerate(states):
        roots[i] = backend.pack_state(s)
    if seed is None:
        seed = int(np.random.SeedSequence().generate_state(1, dtype=np.uint64)[0])
    ts = searcher(game, n, simulations)
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
