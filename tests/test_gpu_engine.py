"""GPU: the Engine facade and get_move over the CUDA search (reference tests/test_engine_configs.py
and the API contract of SURVEY.md §8b)."""
import glob
import os

import numpy as np
import pytest

from conftest import REPO, load_golden

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cfg", sorted(glob.glob(os.path.join(REPO, "configs", "*.yaml"))))
def test_engine_config_basic_move_like_reference(cfg):
    # reference tests/test_engine_configs.py:17-22: one 1-simulation search per config must not raise
    from engine.engine import Engine
    import engine.mcts as mcts
    eng = Engine(cfg)
    assert len(eng.legal_moves()) > 0
    mv = mcts.get_move(eng.get_state(), eng.values[0], eng.policy, eng.backend, 1, eng.config['mcts']['c_puct'], 1)
    assert mv[0] in [m[0] for m in eng.legal_moves()]


def test_get_move_matches_reference_golden_with_first_policy():
    import engine.mcts as mcts
    from engine.games.connect4 import c4_backend as c4
    from engine.games.chess import chess_backend as ch
    from engine.policy_functions import Policy
    from engine.value_functions import Value
    for cs in load_golden("c4_search.json")[:12]:
        s = c4.State([list(r) for r in cs["rows"]], cs["turn"])
        mv = mcts.get_move(s, Value(cs["evaluator"]), Policy(cs["policy"]), c4, cs["sims"], cs["c"], cs["batch"])
        assert mv == (cs["moves"][cs["best"]][0], 0)
    for cs in load_golden("chess_search.json")[:6]:
        s = ch.state_from_fen(cs["fen"])
        mv = mcts.get_move(s, Value("crude_chess_score"), Policy(cs["policy"]), ch, cs["sims"], cs["c"], cs["batch"])
        want = cs["moves"][cs["best"]]
        assert mv == (tuple(want[:4]), want[4])


def test_play_mcts_parallel_plays_full_games_and_dataset():
    from engine.engine import Engine
    eng = Engine({"game": "connect4", "backend": "c4_backend", "value_function": "random_rollout", "threads": 24})
    unfinished = set(range(24))
    for ply in range(43):
        if not unfinished:
            break
        res = eng.play_mcts_parallel(sorted(unfinished), simulations=64, c=1.4)
        assert set(res) == unfinished
        unfinished -= {i for i, r in res.items() if r is not None}
    assert not unfinished
    assert all(h.result in (-1, 0, 1) for h in eng.history)
    x, y = eng.get_dataset()
    assert x.shape[1:] == (2, 6, 7) and len(x) == len(y) == sum(len(h.states) for h in eng.history)
    st = eng.last_search_stats(0)
    assert st["visits"].sum() == 64
    # a finished game returns its result without searching (engine.py:122-125)
    assert eng.play_mcts(0, 8) == eng.history[0].result


def test_random_policy_and_rollout_are_valid_searches():
    from zeroclone_b200 import _ffi
    from zeroclone_b200.search import TreeSearch
    n, sims = 512, 320
    ts = TreeSearch(_ffi.GAME_C4, n, sims)
    ts.set_roots(np.zeros(n, dtype=_ffi.C4_STATE_DTYPE))
    ts.run(sims, 1.4, 32, _ffi.EVAL_C4_ROLLOUT, _ffi.POLICY_RANDOM, seed=7)
    out = ts.results()
    assert (out["visits"].sum(axis=1) == sims).all() and (out["result"]["nodes"] == sims + 1).all()
    assert (np.abs(out["value_sums"]) <= out["visits"] + 1e-9).all()          # rollout values are in {-1,0,1}
    # first expansions are spread over the seven columns (uniform random choice), trees differ
    assert len(set(map(tuple, out["visits"].tolist()))) > 20     # visit counts are quantised by the batch of 32
    # centre columns win more rollouts for the first player than edge columns (sanity of the evaluator)
    q = (out["value_sums"] / np.maximum(1, out["visits"])).mean(axis=0)
    order = out["moves"][0]["fr"][:7].tolist()
    assert q[order.index(3)] > q[order.index(0)] and q[order.index(3)] > q[order.index(6)]
    # same seed -> same trees; different seed -> different trees
    h1 = ts.tree_hash()
    ts.set_roots(np.zeros(n, dtype=_ffi.C4_STATE_DTYPE))
    ts.run(sims, 1.4, 32, _ffi.EVAL_C4_ROLLOUT, _ffi.POLICY_RANDOM, seed=7)
    assert (ts.tree_hash() == h1).all()
    ts.set_roots(np.zeros(n, dtype=_ffi.C4_STATE_DTYPE))
    ts.run(sims, 1.4, 32, _ffi.EVAL_C4_ROLLOUT, _ffi.POLICY_RANDOM, seed=8)
    assert (ts.tree_hash() != h1).mean() > 0.9


def test_chess_engine_with_network_evaluator_runs():
    from engine.engine import Engine
    eng = Engine({"game": "chess", "backend": "chess_backend", "value_function": "network_latest", "threads": 6,
                  "value": {"model_type": "chess_value", "batch_size": 256}})
    res = eng.play_mcts_parallel(range(6), simulations=96, c=1.4)
    assert set(res) == set(range(6)) and all(r is None for r in res.values())
    assert all(len(h.states) == 2 for h in eng.history)


def test_selfplay_refill_and_training_cycle(tmp_path, monkeypatch):
    """simulate_games refill semantics (train.py:151-170) + one training cycle writing latest.pth"""
    import importlib.util
    import torch
    from zeroclone_b200.engine import Engine
    from zeroclone_b200.models import core
    from zeroclone_b200.selfplay import simulate_games
    eng = Engine({"game": "connect4", "backend": "c4_backend", "value_function": "network_latest", "threads": 8,
                  "mcts": {"simulations": 32, "c_puct": 1.4}, "value": {"model_type": "connect4_value", "batch_size": 256}})
    sp = simulate_games(eng, 20)
    assert len(sp["results"]) == 20 and all(r in (-1, 0, 1) for r in sp["results"])
    assert len(eng.states) == 20 and sp["games_per_hour"] > 0
    x, y = eng.get_dataset()
    assert len(x) == sum(len(h.states) for h in eng.history)
    spec = importlib.util.spec_from_file_location("zc_train", os.path.join(REPO, "scripts", "train.py"))
    tr = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tr)
    monkeypatch.setattr(core, "_ROOT", tmp_path)
    (tmp_path / "connect4_value").mkdir()
    loss = tr.train_and_save_latest("connect4_value", eng.values[0].model, x, y, epochs=1, lr=1e-3, batch_size=64,
                                    device=torch.device("cuda", 0), rank=0, world=1)
    assert np.isfinite(loss) and (tmp_path / "connect4_value" / "latest.pth").exists()
    assert tr.schedule_hyperparams(0)["simulations"] == 100 and tr.schedule_hyperparams(30)["simulations"] == 800
