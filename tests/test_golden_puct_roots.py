"""CPU: the oracle's PUCT restatement reproduces its committed golden vectors (the library's own definition, frozen -- the
reference has no PUCT), and the benchmark's root sets are the committed ones (SHA-256)."""
import hashlib

import pytest

from conftest import load_golden
from oracle import zc_oracle as zo
from zeroclone_b200.workloads import c4_roots_set_b, chess_roots_set_b

G = load_golden("puct_and_roots.json")


def oracle_case(cs):
    if cs["game"] == "c4":
        return zo.search_puct(zo.GAME_C4, zo.c4_from_moves(cs["cols"]), cs["sims"], cs["c"], cs["batch"], getattr(zo, "EVAL_" + cs["evaluator"].upper()),
                              cs["virtual_loss"], cs["prior_weight"])
    return zo.search_puct(zo.GAME_CHESS, zo.ch_from_fen(cs["fen"]), cs["sims"], cs["c"], cs["batch"], zo.EVAL_CHESS_CRUDE, cs["virtual_loss"], cs["prior_weight"])


@pytest.mark.parametrize("i", range(len(G["puct"])))
def test_oracle_puct_golden(i):
    cs = G["puct"][i]
    o = oracle_case(cs)
    assert o.Na == cs["Na"] and o.Wa == cs["Wa"] and o.best == cs["best"]
    assert o.nodes_created == cs["nodes_created"] and o.max_leaf_depth == cs["max_leaf_depth"] and str(o.tree_hash) == cs["tree_hash"]
    assert sum(o.Na) == cs["sims"]


def test_benchmark_root_sets_are_the_committed_ones():
    d = G["root_sets_sha256"]
    assert hashlib.sha256(c4_roots_set_b(4096).tobytes()).hexdigest() == d["c4_set_b_4096"]
    assert hashlib.sha256(c4_roots_set_b(64, first_tree_id=30000).tobytes()).hexdigest() == d["c4_set_b_first_id_30000_x64"]
    assert hashlib.sha256(chess_roots_set_b(2048).tobytes()).hexdigest() == d["chess_set_b_2048"]


@pytest.mark.gpu
@pytest.mark.parametrize("i", range(len(G["puct"])))
def test_gpu_puct_golden(i):
    import numpy as np
    from zeroclone_b200 import _ffi
    from zeroclone_b200.search import TreeSearch, c4_pack_cols
    cs = G["puct"][i]
    c4 = cs["game"] == "c4"
    if c4:
        roots = np.zeros(1, dtype=_ffi.C4_STATE_DTYPE)
        roots[0] = c4_pack_cols(cs["cols"]) + (0,)
        ev = {"c4_positional": _ffi.EVAL_C4_POSITIONAL, "c4_terminal": _ffi.EVAL_C4_TERMINAL}[cs["evaluator"]]
    else:
        import ctypes as C
        s = _ffi.ChessState()
        _ffi.check(_ffi.lib().zc_chess_from_fen(cs["fen"].encode(), C.byref(s)))
        roots = np.frombuffer(bytes(s), dtype=_ffi.CHESS_STATE_DTYPE).copy()
        ev = _ffi.EVAL_CHESS_CRUDE
    ts = TreeSearch(_ffi.GAME_C4 if c4 else _ffi.GAME_CHESS, 1, cs["sims"])
    ts.set_mode(_ffi.SELECT_PUCT, cs["virtual_loss"], cs["prior_weight"])
    ts.set_roots(roots)
    ts.run(cs["sims"], cs["c"], cs["batch"], ev, _ffi.POLICY_FIRST)
    out, h = ts.results(), ts.tree_hash()
    k = len(cs["Na"])
    assert out["visits"][0][:k].tolist() == cs["Na"] and out["value_sums"][0][:k].tolist() == cs["Wa"]
    assert int(out["result"][0]["best"]) == cs["best"] and str(int(h[0])) == cs["tree_hash"]
