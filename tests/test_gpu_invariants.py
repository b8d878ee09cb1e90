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


@pytest.mark.parametrize("policy", [_ffi.POLICY_FIRST, _ffi.POLICY_RANDOM])
def test_chess_full_size_2048_trees_1600_sims(policy):
    n, sims = 2048, 1600
    ts = TreeSearch(_ffi.GAME_CHESS, n, sims)
    ts.set_roots(chess_roots_set_b(n))
    ts.run(sims, 1.4, 32, _ffi.EVAL_CHESS_CRUDE, policy, seed=5)
    check_sampled(ts, n, sims, 1000.0, sample=12)
