"""GPU: the C-ABI entry points added in round 2 -- pipelined root readout, in-place tower weight update, and the
error paths of the self-play step and of the pooled search handles (nothing may fail silently)."""
import ctypes as C

import numpy as np
import pytest
import torch

from zeroclone_b200 import _ffi, mcts
from zeroclone_b200.evaluator import NetEvaluator
from zeroclone_b200.search import TreeSearch
from zeroclone_b200.workloads import c4_roots_set_a, c4_roots_set_b, chess_roots_set_b

pytestmark = pytest.mark.gpu


def test_results_begin_end_is_the_same_readout_and_pipelines():
    n, sims = 512, 200
    a, b = c4_roots_set_b(n), c4_roots_set_b(n, first_tree_id=5000)
    ts = TreeSearch(_ffi.GAME_C4, n, sims)
    want = []
    for roots in (a, b):
        ts.set_roots(roots)
        ts.run(sims, 1.4, 32, _ffi.EVAL_C4_POSITIONAL, _ffi.POLICY_FIRST)
        want.append(ts.results(stats=False)["result"].copy())
    # two searches enqueued back to back, no host wait in between; _end collects the latest
    dev = [torch.from_numpy(r.view(np.uint8).reshape(n, -1).copy()).cuda() for r in (a, b)]
    for k, d in enumerate(dev):
        ts.set_roots_dev(d.data_ptr(), n)
        ts.run(sims, 1.4, 32, _ffi.EVAL_C4_POSITIONAL, _ffi.POLICY_FIRST)
        ts.results_begin()
        if k == 0:
            first = ts.results_end()["result"].copy()
            ts.results_begin()                     # begin twice: the second readout of the same trees is identical
    last = ts.results_end()["result"]
    fields = [f for f in first.dtype.names if f != "_pad"]
    assert all(np.array_equal(first[f], want[0][f]) for f in fields)
    assert all(np.array_equal(last[f], want[1][f]) for f in fields)
    with pytest.raises(_ffi.ZcError) as e:
        ts.results_end()
    assert e.value.code == _ffi.ZC_ESTATE


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_tower_update_weights_equals_a_fresh_tower(dtype):
    from zeroclone_b200.models.connect4_value.network import ValueNetwork
    torch.manual_seed(1)
    m1 = ValueNetwork().eval()
    torch.manual_seed(2)
    m2 = ValueNetwork().eval()
    with torch.no_grad():                      # non-trivial BatchNorm statistics, as after training
        for mod in m2.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.running_mean.normal_(0, 0.1)
                mod.running_var.uniform_(0.5, 1.5)
    x = (torch.rand(3000, 2, 6, 7) < 0.3).to("cuda", dtype)
    ev = NetEvaluator(m1, "cuda", dtype)
    out1 = ev(x).clone()
    ev.update_weights(m2)
    out2 = ev(x).clone()
    fresh = NetEvaluator(m2, "cuda", dtype)(x)
    assert torch.equal(out2, fresh) and not torch.equal(out1, out2)
    ev.update_weights(m1)
    assert torch.equal(ev(x), out1)
    assert _ffi.lib().zc_tower_fault(ev._h) == 0


def _advance(ts, roots_dev, n, chess=False, hist_cap=8):
    ones = torch.ones(n, dtype=torch.uint8, device="cuda")
    res = np.zeros(n, dtype=np.int32)
    mv = np.zeros(n, dtype=_ffi.CHESS_MOVE_DTYPE)
    hist = torch.zeros((n, 2, hist_cap, 8), dtype=torch.uint8, device="cuda") if chess else None
    hlen = torch.zeros((n, 2), dtype=torch.int32, device="cuda") if chess else None
    rc = _ffi.lib().zc_search_advance(ts._h, roots_dev.data_ptr(), ones.data_ptr(), hist.data_ptr() if chess else None,
                                      hlen.data_ptr() if chess else None, hist_cap if chess else 0, res.ctypes.data_as(C.c_void_p),
                                      mv.ctypes.data_as(C.c_void_p), None)
    return rc, res, hlen


def test_advance_refuses_trees_that_outgrew_their_arena():
    n, sims = 16, 400
    roots = c4_roots_set_a(n)
    dev = torch.from_numpy(roots.view(np.uint8).reshape(n, -1).copy()).cuda()
    before = dev.clone()
    ts = TreeSearch(_ffi.GAME_C4, n, sims, arena_slots_per_tree=400)       # far too small for 400 simulations
    ts.set_roots_dev(dev.data_ptr(), n)
    ts.run(sims, 1.4, 32, _ffi.EVAL_C4_POSITIONAL, _ffi.POLICY_FIRST)
    rc, res, _ = _advance(ts, dev, n)
    assert rc == _ffi.ZC_ECAPACITY and b"outgrew its arena" in _ffi.lib().zc_last_error()
    assert (res == _ffi.RESULT_ONGOING).all() and torch.equal(dev, before)      # no move from a truncated tree was played


def test_advance_refuses_when_no_simulation_ran():
    n = 8
    roots = c4_roots_set_a(n)
    dev = torch.from_numpy(roots.view(np.uint8).reshape(n, -1).copy()).cuda()
    ts = TreeSearch(_ffi.GAME_C4, n, 64)
    ts.set_roots_dev(dev.data_ptr(), n)
    rc, res, _ = _advance(ts, dev, n)
    assert rc == _ffi.ZC_ESTATE and b"no visited root move" in _ffi.lib().zc_last_error()
    assert (res == _ffi.RESULT_ONGOING).all()


def test_advance_refuses_a_full_chess_history():
    n, sims, cap = 4, 64, 8
    roots = chess_roots_set_b(n)
    dev = torch.from_numpy(roots.view(np.uint8).reshape(n, -1).copy()).cuda()
    ts = TreeSearch(_ffi.GAME_CHESS, n, sims)
    ones = torch.ones(n, dtype=torch.uint8, device="cuda")
    hist = torch.zeros((n, 2, cap, 8), dtype=torch.uint8, device="cuda")
    hlen = torch.zeros((n, 2), dtype=torch.int32, device="cuda")
    res = np.zeros(n, dtype=np.int32)
    mv = np.zeros(n, dtype=_ffi.CHESS_MOVE_DTYPE)
    rc = 0
    for ply in range(2 * cap + 2):
        ts.set_roots_dev(dev.data_ptr(), n)
        ts.run(sims, 1.4, 32, _ffi.EVAL_CHESS_CRUDE, _ffi.POLICY_FIRST)
        rc = _ffi.lib().zc_search_advance(ts._h, dev.data_ptr(), ones.data_ptr(), hist.data_ptr(), hlen.data_ptr(), cap,
                                          res.ctypes.data_as(C.c_void_p), mv.ctypes.data_as(C.c_void_p), None)
        if rc != 0:
            break
        ones = torch.from_numpy((res == _ffi.RESULT_ONGOING).astype(np.uint8)).cuda()
        if not ones.any():
            break
    assert int(hlen.max()) <= cap
    if int(hlen.max()) == cap and ones.any():
        assert rc == _ffi.ZC_ECAPACITY and b"history" in _ffi.lib().zc_last_error()


def test_searcher_pool_survives_a_failed_growth():
    mcts.release_all()
    ts = mcts.searcher(_ffi.GAME_C4, 64, 100)
    assert ts.max_trees >= 64
    with pytest.raises(_ffi.ZcError):
        mcts.searcher(_ffi.GAME_C4, 64, 200_000_000)          # ~1.8 TB of node arenas: cudaMalloc refuses
    ts2 = mcts.searcher(_ffi.GAME_C4, 32, 100)                 # the pool is usable and consistent afterwards
    roots = c4_roots_set_b(32)
    ts2.set_roots(roots)
    ts2.run(100, 1.4, 32, _ffi.EVAL_C4_POSITIONAL, _ffi.POLICY_FIRST)
    assert (ts2.results(stats=False)["result"]["root_visits"] == 100).all()
    # growing only `sims` keeps the tree capacity
    cap = ts2.max_trees
    ts3 = mcts.searcher(_ffi.GAME_C4, 8, 400)
    assert ts3.max_trees >= cap and ts3.max_sims >= 400
    mcts.release_all()
