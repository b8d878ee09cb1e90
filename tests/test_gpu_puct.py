"""GPU: the opt-in PUCT mode (stored priors, virtual loss; zeroclone_b200/csrc/puct.cuh) against its CPU restatement
oracle/zc_oracle.c:zo_search_puct -- visits, value sums, chosen move and the whole-tree hash, bit for bit.  The reference
has no PUCT (mcts.cpp:41-63 is UCB1), so this pins the kernels on the library's own definition; the UCB1 path is untouched
(its tests run next to these on the same handles)."""
import numpy as np
import pytest
import torch

from oracle import zc_oracle as zo
from test_gpu_parity_bench_sets import c4_oracle_state, chess_oracle_state, c4_planes
from zeroclone_b200 import _ffi
from zeroclone_b200.search import TreeSearch
from zeroclone_b200.workloads import c4_roots_set_b, chess_roots_set_b

pytestmark = pytest.mark.gpu
EV = {"c4_terminal": (_ffi.EVAL_C4_TERMINAL, zo.EVAL_C4_TERMINAL), "c4_positional": (_ffi.EVAL_C4_POSITIONAL, zo.EVAL_C4_POSITIONAL)}


def check(out, hashes, i, o):
    assert out["visits"][i][:o.n_moves].tolist() == o.Na, (i, out["visits"][i][:o.n_moves].tolist(), o.Na)
    assert out["value_sums"][i][:o.n_moves].tolist() == o.Wa, i
    assert int(out["result"][i]["best"]) == o.best and int(hashes[i]) == o.tree_hash, i
    assert int(out["result"][i]["nodes"]) == o.nodes_created and int(out["result"][i]["max_leaf_depth"]) == o.max_leaf_depth, i


@pytest.mark.parametrize("evaluator,sims,c,batch,vloss", [
    ("c4_positional", 800, 1.4, 32, 1.0), ("c4_terminal", 800, 2.5, 32, 1.0), ("c4_positional", 333, 0.7, 8, 0.5),
    ("c4_positional", 97, 1.4, 1, 1.0), ("c4_terminal", 500, 1.25, 32, 0.0)])
def test_c4_puct_matches_oracle(evaluator, sims, c, batch, vloss):
    roots = np.concatenate([c4_roots_set_b(48), c4_roots_set_b(16, first_tree_id=30000)])
    ts = TreeSearch(_ffi.GAME_C4, len(roots), sims)
    ts.set_mode(_ffi.SELECT_PUCT, vloss, 0)
    ts.set_roots(roots)
    ts.run(sims, c, batch, EV[evaluator][0], _ffi.POLICY_FIRST)
    out, hashes = ts.results(), ts.tree_hash()
    for i, rec in enumerate(roots):
        check(out, hashes, i, zo.search_puct(zo.GAME_C4, c4_oracle_state(rec), sims, c, batch, EV[evaluator][1], vloss, 0))
    # the same handle back in the reference's mode: UCB1 results are the reference's again
    ts.set_mode(_ffi.SELECT_UCB1)
    ts.set_roots(roots)
    ts.run(sims, c, batch, EV[evaluator][0], _ffi.POLICY_FIRST)
    out, hashes = ts.results(), ts.tree_hash()
    for i in (0, 17, 63):
        o = zo.search(zo.GAME_C4, c4_oracle_state(roots[i]), sims, c, batch, EV[evaluator][1], zo.POLICY_FIRST)
        assert out["visits"][i][:o.n_moves].tolist() == o.Na and int(hashes[i]) == o.tree_hash


@pytest.mark.parametrize("sims,c,batch,prior_weight", [(800, 1.4, 32, 0), (400, 30.0, 32, 2), (250, 5.0, 7, 1)])
def test_chess_puct_matches_oracle(sims, c, batch, prior_weight):
    roots = np.concatenate([chess_roots_set_b(26), chess_roots_set_b(6, first_tree_id=9000)])
    ts = TreeSearch(_ffi.GAME_CHESS, len(roots), sims)
    ts.set_mode(_ffi.SELECT_PUCT, 1.0, prior_weight)
    ts.set_roots(roots)
    ts.run(sims, c, batch, _ffi.EVAL_CHESS_CRUDE, _ffi.POLICY_FIRST)
    out, hashes = ts.results(), ts.tree_hash()
    for i, rec in enumerate(roots):
        check(out, hashes, i, zo.search_puct(zo.GAME_CHESS, chess_oracle_state(rec), sims, c, batch, zo.EVAL_CHESS_CRUDE, 1.0, prior_weight))


def test_c4_puct_split_phase_and_root_priors():
    """external evaluator (select / evaluate / back up kernels) with an exact function of the packed planes, and priors
    supplied for the roots"""
    w = torch.tensor([1, 2, 3, 4, 3, 2, 1], dtype=torch.float32)

    class Dyadic:
        dtype = torch.bfloat16

        def __call__(self, planes, out):
            out.copy_(((planes[:, 0].float() - planes[:, 1].float()) * w.to(planes.device).view(1, 1, 7)).sum(dim=(1, 2)) / 64.0)
            return out

    def ext(states_u8):
        p = c4_planes(states_u8)
        return ((p[:, 0] - p[:, 1]) * w.numpy().reshape(1, 1, 7)).sum(axis=(1, 2)) / 64.0

    roots = c4_roots_set_b(40)
    sims = 400
    rng = np.random.default_rng(5)
    pri = rng.dirichlet(np.ones(7), size=len(roots)).astype(np.float32)
    ts = TreeSearch(_ffi.GAME_C4, len(roots), sims)
    ts.set_mode(_ffi.SELECT_PUCT, 1.0, 0)
    ts.set_roots(roots)
    ts.set_root_priors(pri)
    ts.run_network(Dyadic(), sims, 2.0, 32, _ffi.POLICY_FIRST)
    out, hashes = ts.results(), ts.tree_hash()
    for i, rec in enumerate(roots):
        k = int(out["result"][i]["n_moves"])
        o = zo.search_puct(zo.GAME_C4, c4_oracle_state(rec), sims, 2.0, 32, zo.EVAL_EXTERNAL, 1.0, 0, root_priors=pri[i][:k].tolist(), external=ext)
        check(out, hashes, i, o)
    # priors matter: the most visited root move follows a one-hot prior when values are flat
    ts.set_roots(roots[:1])
    onehot = np.full((1, 7), 1e-4, dtype=np.float32)
    onehot[0, 5] = 1.0
    ts.set_root_priors(onehot)
    ts.run(200, 4.0, 32, _ffi.EVAL_C4_TERMINAL, _ffi.POLICY_FIRST)
    assert int(ts.results()["result"][0]["best"]) == 5


def test_engine_config_selects_puct():
    from zeroclone_b200.engine import Engine
    cfg = {"game": "connect4", "backend": "c4_backend", "value_function": "c4_positional", "threads": 3,
           "mcts": {"simulations": 200, "c_puct": 1.4, "select": "puct", "virtual_loss": 1.0}}
    eng = Engine(cfg)
    res = eng.play_mcts_parallel([0, 1, 2], simulations=200, c=1.4)
    assert set(res) == {0, 1, 2} and all(r is None for r in res.values())
    st = eng.last_search_stats(0)
    o = zo.search_puct(zo.GAME_C4, zo.c4_init(), 200, 1.4, 32, zo.EVAL_C4_POSITIONAL, 1.0, 0)
    assert st["visits"].tolist() == o.Na and st["best"] == o.best
    with pytest.raises(ValueError):
        Engine(dict(cfg, mcts={"select": "alphabeta"}))


def test_chess_puct_split_phase_matches_oracle():
    """external-evaluator kernels in PUCT mode on chess, with material read back from the packed planes (exact in fp32)"""
    vals = torch.tensor([1, 3, 3, 5, 9, 0, -1, -3, -3, -5, -9, 0], dtype=torch.float32, device="cuda")

    class Material:
        dtype = torch.float16

        def __call__(self, planes, out):
            mat = (planes[:, :12].float().sum(dim=(2, 3)) * vals).sum(dim=1)
            out.copy_(mat * (2 * planes[:, 12, 0, 0].float() - 1))
            return out

    pv = {ord(k): v for k, v in {'P': 1, 'N': 3, 'B': 3, 'R': 5, 'Q': 9, 'p': -1, 'n': -3, 'b': -3, 'r': -5, 'q': -9}.items()}

    def ext(states_u8):
        out = np.zeros(len(states_u8))
        for i, s in enumerate(states_u8):
            out[i] = (1 - 2 * int(s[64])) * sum(pv.get(int(b), 0) for b in s[:64])
        return out

    roots = chess_roots_set_b(16)
    sims = 300
    ts = TreeSearch(_ffi.GAME_CHESS, len(roots), sims)
    ts.set_mode(_ffi.SELECT_PUCT, 1.0, 1)
    ts.set_roots(roots)
    ts.run_network(Material(), sims, 10.0, 32, _ffi.POLICY_FIRST)
    out, hashes = ts.results(), ts.tree_hash()
    for i, rec in enumerate(roots):
        check(out, hashes, i, zo.search_puct(zo.GAME_CHESS, chess_oracle_state(rec), sims, 10.0, 32, zo.EVAL_EXTERNAL, 1.0, 1, external=ext))


def test_device_selfplay_in_puct_mode_finishes_games():
    from zeroclone_b200 import mcts
    from zeroclone_b200.games.connect4 import c4_backend
    from zeroclone_b200.policy_functions import Policy
    from zeroclone_b200.selfplay import DeviceSelfPlay
    from zeroclone_b200.value_functions import Value
    sp = DeviceSelfPlay(c4_backend, Value("c4_positional"), Policy("random"), n_slots=64, device=0,
                        mode=mcts.select_mode({"select": "puct", "virtual_loss": 1.0}))
    out = sp.play(96, 200, 1.4, seed=3, record=True)
    assert len(out["results"]) == 96 and all(r in (-1, 0, 1) for r in out["results"])
    planes, labels = out["dataset"]
    assert planes.shape[1:] == (2, 6, 7) and len(planes) == len(labels) == out["moves"] + 96
