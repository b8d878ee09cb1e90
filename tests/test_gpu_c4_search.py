"""GPU parity: the CUDA search against the oracle / golden vectors, bit-exact (C4)."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import zc_oracle as zo
from zeroclone_b200 import _ffi
from zeroclone_b200.search import TreeSearch, c4_pack_cols, c4_pack_rows

pytestmark = pytest.mark.gpu

EVAL = {"c4_terminal": _ffi.EVAL_C4_TERMINAL, "c4_positional": _ffi.EVAL_C4_POSITIONAL}
POL = {"first": _ffi.POLICY_FIRST, "last": _ffi.POLICY_LAST}


def roots_array(packed):
    a = np.zeros(len(packed), dtype=_ffi.C4_STATE_DTYPE)
    for i, (x, o, t) in enumerate(packed):
        a[i] = (x, o, t, 0)
    return a


def test_golden_c4_search_cases_bit_exact():
    cases = load_golden("c4_search.json")
    ts = TreeSearch(_ffi.GAME_C4, max_trees=4, max_sims=800)
    for cs in cases:
        ts.set_roots(roots_array([c4_pack_rows(cs["rows"], cs["turn"])]))
        ts.run(cs["sims"], cs["c"], cs["batch"], EVAL[cs["evaluator"]], POL[cs["policy"]])
        out = ts.results()
        r = out["result"][0]
        k = int(r["n_moves"])
        tag = (cs["cols"], cs["evaluator"], cs["policy"], cs["sims"], cs["c"], cs["batch"])
        assert [int(m) for m in out["moves"][0]["fr"][:k]] == [m[0] for m in cs["moves"]], tag
        assert out["visits"][0][:k].tolist() == cs["Na"], tag
        assert out["value_sums"][0][:k].tolist() == cs["Wa"], tag
        assert int(r["best"]) == cs["best"], tag
        assert int(r["nodes"]) == cs["nodes_created"], tag
        assert int(r["sum_leaf_depth"]) == cs["sum_leaf_depth"] and int(r["max_leaf_depth"]) == cs["max_leaf_depth"], tag


def random_roots(n, seed):
    rng = np.random.default_rng(seed)
    packed, states = [], []
    for i in range(n):
        while True:
            s, cols = zo.c4_init(), []
            for _ in range(int(rng.integers(0, 36))):
                legal = sorted(zo.c4_legal(s))
                if not legal:
                    break
                c = int(rng.choice(legal))
                cols.append(c)
                s = zo.c4_play(s, c)
            if zo.c4_legal(s):
                break
        packed.append(c4_pack_cols(cols))
        states.append(s)
    return packed, states


@pytest.mark.parametrize("evaluator,policy,sims,c,batch", [
    ("c4_terminal", "first", 800, 1.4, 32), ("c4_positional", "first", 800, 1.4, 32),
    ("c4_positional", "last", 500, 2.0, 32), ("c4_terminal", "first", 333, 1.25, 5),
    ("c4_positional", "first", 97, 0.7, 1), ("c4_positional", "last", 1600, 1.4, 32),
])
def test_whole_tree_hash_matches_oracle_on_random_roots(evaluator, policy, sims, c, batch):
    n = 96
    packed, states = random_roots(n, seed=sims + batch)
    ts = TreeSearch(_ffi.GAME_C4, max_trees=n, max_sims=sims)
    ts.set_roots(roots_array(packed))
    ts.run(sims, c, batch, EVAL[evaluator], POL[policy])
    out = ts.results()
    hashes = ts.tree_hash()
    oev = {"c4_terminal": zo.EVAL_C4_TERMINAL, "c4_positional": zo.EVAL_C4_POSITIONAL}[evaluator]
    opol = {"first": zo.POLICY_FIRST, "last": zo.POLICY_LAST}[policy]
    for i in range(n):
        o = zo.search(zo.GAME_C4, states[i], sims, c, batch, oev, opol)
        k = o.n_moves
        assert int(out["result"][i]["n_moves"]) == k
        assert out["visits"][i][:k].tolist() == o.Na, i
        assert out["value_sums"][i][:k].tolist() == o.Wa, i
        assert int(out["result"][i]["best"]) == o.best, i
        assert int(out["result"][i]["nodes"]) == o.nodes_created, i
        assert int(out["result"][i]["reevaluated_leaves"]) == o.reevaluated_leaves, i
        assert int(hashes[i]) == o.tree_hash, i


def test_many_identical_trees_agree_and_counters():
    n = 4096
    ts = TreeSearch(_ffi.GAME_C4, max_trees=n, max_sims=800)
    ts.set_roots(roots_array([c4_pack_cols([])] * n))
    ts.run(800, 1.4, 32, _ffi.EVAL_C4_POSITIONAL, _ffi.POLICY_FIRST)
    out = ts.results()
    assert (out["visits"] == out["visits"][0]).all()
    assert out["visits"][0].tolist() == [122, 97, 129, 129, 129, 97, 97]
    assert (out["result"]["best_move"][:, 0] == 2).all()
    assert len(set(ts.tree_hash().tolist())) == 1
    cnt = ts.counters()
    assert cnt["simulations"] == n * 800 and cnt["nodes"] == n * 801


def test_arena_overflow_is_reported_not_silent():
    ts = TreeSearch(_ffi.GAME_C4, max_trees=2, max_sims=800, arena_slots_per_tree=512)
    ts.set_roots(roots_array([c4_pack_cols([])] * 2))
    ts.run(800, 1.4, 32, _ffi.EVAL_C4_TERMINAL, _ffi.POLICY_FIRST)
    with pytest.raises(_ffi.ZcError) as e:
        ts.results()
    assert e.value.code == _ffi.ZC_ECAPACITY
