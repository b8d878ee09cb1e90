"""GPU: size-independent properties at BASELINE.json's full sizes, including the modes that have no
bit-exact oracle (random expansion policy, random rollouts, neural evaluator): structural invariants of
the reference's statistics on sampled trees, conservation laws over all trees, determinism."""
import numpy as np
import pytest
import torch

from zeroclone_b200 import _ffi
from zeroclone_b200.evaluator import NetEvaluator
from zeroclone_b200.search import TreeSearch
from zeroclone_b200.workloads import c4_roots_set_b, chess_roots_set_b

pytestmark = pytest.mark.gpu


def check_sampled(ts, n, sims, max_abs, sample=24):
    out = ts.results()
    res = out["result"]
    assert (res["root_visits"] == sims).all()
    assert (out["visits"].sum(axis=1) == sims).all()
    assert (res["nodes"] + res["reevaluated_leaves"] == sims + 1).all()       # one node per simulation unless a move-less node is re-evaluated
    best = res["best"]
    rows = np.arange(n)
    assert (out["visits"][rows, best] == out["visits"].max(axis=1)).all()     # most-visited child ...
    first_max = (out["visits"] == out["visits"].max(axis=1, keepdims=True)).argmax(axis=1)
    assert (best == first_max).all()                                          # ... lowest index on ties (mcts.cpp:150-155)
    for t in np.linspace(0, n - 1, sample).astype(int):
        tv = ts.read_tree(int(t))
        assert tv.check_invariants(max_abs) == int(res["nodes"][t])
    return out


@pytest.mark.parametrize("evaluator,policy,max_abs", [
    (_ffi.EVAL_C4_POSITIONAL, _ffi.POLICY_FIRST, 1.0), (_ffi.EVAL_C4_ROLLOUT, _ffi.POLICY_RANDOM, 1.0),
    (_ffi.EVAL_C4_TERMINAL, _ffi.POLICY_LAST, 1.0)])
def test_c4_full_size_4096_trees_800_sims(evaluator, policy, max_abs):
    n, sims = 4096, 800
    ts = TreeSearch(_ffi.GAME_C4, n, sims)
    ts.set_roots(c4_roots_set_b(n))
    ts.run(sims, 1.4, 32, evaluator, policy, seed=11)
    check_sampled(ts, n, sims, max_abs)
    # determinism: the same seed reproduces every tree bit for bit
    h1 = ts.tree_hash()
    ts.set_roots(c4_roots_set_b(n))
    ts.run(sims, 1.4, 32, evaluator, policy, seed=11)
    assert (ts.tree_hash() == h1).all()


def test_c4_full_size_value_net_split_phase():
    n, sims = 4096, 800
    torch.manual_seed(0)
    from zeroclone_b200.models.connect4_value.network import ValueNetwork
    ev = NetEvaluator(ValueNetwork().eval(), "cuda")
    ts = TreeSearch(_ffi.GAME_C4, n, sims)
    ts.set_roots(c4_roots_set_b(n))
    ts.run_network(ev, sims, 1.4, 32, _ffi.POLICY_FIRST)
    check_sampled(ts, n, sims, 1.0, sample=12)


def test_chess_full_size_value_net_split_phase():
    """BASELINE configs[3]: 2048 trees x 800 sims, stub leaves + the fused tower on 17x8x8 planes"""
    n, sims = 2048, 800
    torch.manual_seed(0)
    from zeroclone_b200.models.chess_value.network import ValueNetwork
    ev = NetEvaluator(ValueNetwork().eval(), "cuda")
    ts = TreeSearch(_ffi.GAME_CHESS, n, sims)
    ts.set_roots(chess_roots_set_b(n))
    ts.run_network(ev, sims, 1.4, 32, _ffi.POLICY_FIRST)
    check_sampled(ts, n, sims, 1.0, sample=12)
    c = ts.counters()
    assert c["simulations"] == n * sims
    assert c["arena_slots_used"] < 8 * c["nodes"], "leaves should be 3-slot stubs"


@pytest.mark.parametrize("policy", [_ffi.POLICY_FIRST, _ffi.POLICY_RANDOM])
def test_chess_full_size_2048_trees_1600_sims(policy):
    n, sims = 2048, 1600
    ts = TreeSearch(_ffi.GAME_CHESS, n, sims)
    ts.set_roots(chess_roots_set_b(n))
    ts.run(sims, 1.4, 32, _ffi.EVAL_CHESS_CRUDE, policy, seed=5)
    check_sampled(ts, n, sims, 1000.0, sample=12)


@pytest.mark.parametrize("freedom", [0.0, 3.0])
def test_chess_immediate_value_policy_expansion_order(freedom):
    """Policy.immediate_value (policy_functions.py:14-17): every expansion picks among the untried moves whose
    capture value is >= best untried - policy_freedom.  So after n < k expansions of the root, every expanded
    move is worth at least (best unexpanded - freedom); with freedom 0 the expanded set is a top-n by value."""
    import ctypes as C
    from zeroclone_b200.games.chess import chess_backend as cb
    fen = "r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1"      # Kiwipete: 8 captures of 3 kinds
    root = cb.state_from_fen(fen)
    n = 64
    roots = np.zeros(n, dtype=_ffi.CHESS_STATE_DTYPE)
    for i in range(n):
        roots[i] = cb.pack_state(root)
    seen_sets = set()
    for sims in (5, 12, 30):
        ts = TreeSearch(_ffi.GAME_CHESS, n, sims)
        ts.set_roots(roots)
        ts.set_policy_freedom(freedom)
        ts.run(sims, 1.4, 32, _ffi.EVAL_CHESS_CRUDE, _ffi.POLICY_IMMEDIATE_VALUE, seed=17)
        out = ts.results()
        for t in range(n):
            k = int(out["result"][t]["n_moves"])
            vals = out["moves"][t]["value"][:k]
            expanded = out["visits"][t][:k] > 0
            assert int(expanded.sum()) == min(sims, k)
            if expanded.all():
                continue
            assert vals[expanded].min() >= vals[~expanded].max() - freedom, (sims, t)
            seen_sets.add((sims, tuple(np.nonzero(expanded)[0].tolist())))
        if sims == 30:
            check_sampled(ts, n, sims, 1000.0, sample=4)
    assert len(seen_sets) > 10, "trees of one position must not all expand the same moves (device RNG per tree)"


@pytest.mark.parametrize("freedom,crit", [(0.0, 10.83), (2.0, 24.32), (3.0, 80.08)])
def test_chess_immediate_value_first_pick_is_uniform_over_the_candidates(freedom, crit):
    """random.choice([a for a in moves if a[1] >= best - policy_freedom]) (policy_functions.py:14-17): the FIRST expansion of a
    root picks uniformly among exactly those moves.  20 000 trees on one position (Kiwipete: 46 moves in the reference's rules, captures worth 1 and 3);
    chi-square at p = 0.001 for the candidate count of each freedom (2, 8 and 46 candidates: df 1, 7, 45)."""
    from zeroclone_b200.games.chess import chess_backend as cb
    root = cb.state_from_fen("r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1")
    moves = cb.get_legal_moves(root)
    vals = np.array([m[1] for m in moves])
    cand = np.nonzero(vals >= vals.max() - freedom)[0]
    n = 20000
    roots = np.zeros(n, dtype=_ffi.CHESS_STATE_DTYPE)
    roots[:] = cb.pack_state(root)
    ts = TreeSearch(_ffi.GAME_CHESS, n, 32)
    ts.set_roots(roots)
    ts.set_policy_freedom(freedom)
    ts.run(1, 1.4, 1, _ffi.EVAL_CHESS_CRUDE, _ffi.POLICY_IMMEDIATE_VALUE, seed=23)
    out = ts.results()
    first = out["visits"][:, :len(moves)].argmax(axis=1)
    assert (out["visits"].sum(axis=1) == 1).all()
    counts = np.bincount(first, minlength=len(moves)).astype(np.float64)
    assert counts[np.setdiff1d(np.arange(len(moves)), cand)].sum() == 0          # never a move outside the candidate set
    e = n / len(cand)
    chi2 = float(((counts[cand] - e) ** 2 / e).sum())
    print(f"freedom {freedom}: {len(cand)} candidates of {len(moves)} moves, chi2 {chi2:.1f} (critical {crit})")
    assert chi2 < crit
