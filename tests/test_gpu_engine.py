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


@pytest.mark.parametrize("game", ["connect4", "chess"])
def test_device_selfplay_matches_engine_host_loop(game):
    """Device-resident self-play (advance kernel: apply move + win/draw incl. 50-ply and repetition) must
    play exactly the games the Engine host loop plays (deterministic evaluator + policy)."""
    from zeroclone_b200.engine import Engine
    from zeroclone_b200.policy_functions import Policy
    from zeroclone_b200.selfplay import DeviceSelfPlay, simulate_games
    from zeroclone_b200.value_functions import Value
    n, sims = 6, 40
    cfg = {"game": game, "backend": "c4_backend" if game == "connect4" else "chess_backend", "threads": n,
           "value_function": "c4_positional" if game == "connect4" else "crude_chess_score", "policy_functions": "first",
           "mcts": {"simulations": sims, "c_puct": 1.4}}
    eng = Engine(cfg)
    assert eng.policy.name == "first"
    max_plies = 60 if game == "chess" else 43
    for ply in range(max_plies):                 # chess: bounded number of plies, games need not finish
        unfinished = [i for i in range(n) if eng.history[i].result is None]
        if not unfinished:
            break
        eng.play_mcts_parallel(unfinished, simulations=sims, c=1.4)
    sp = DeviceSelfPlay(eng.backend, Value(cfg["value_function"]), Policy("first"), n_slots=n)
    if game == "connect4":
        out = sp.play(n, sims, 1.4)
        assert out["results"] == [h.result for h in eng.history]
        for g in range(n):
            want = np.array([eng.backend.pack_state(s) for s in eng.history[g].states], dtype=eng.backend.STATE_DTYPE)
            assert np.array_equal(np.stack(out["trajectories"][g]), want), g
        x, y = eng.get_dataset()
        assert np.array_equal(out["dataset"][0], x) and np.array_equal(out["dataset"][1], y)
    else:
        # identical games => every slot plays the same line; run the device loop until the first game ends or 60 plies
        ref_states = eng.history[0].states
        import ctypes as C
        import torch
        from zeroclone_b200 import _ffi, mcts
        dev = torch.device("cuda", 0)
        rec = np.zeros(1, dtype=eng.backend.STATE_DTYPE)
        rec[0] = eng.backend.pack_state(ref_states[0])
        roots = torch.from_numpy(np.repeat(rec, n).view(np.uint8).reshape(n, -1).copy()).to(dev)
        hist = torch.zeros((n, 2, 512, 8), dtype=torch.uint8, device=dev)
        hlen = torch.zeros((n, 2), dtype=torch.int32, device=dev)
        ones = torch.ones(n, dtype=torch.uint8, device=dev)
        ts = mcts.searcher(_ffi.GAME_CHESS, n, sims)
        for ply in range(1, len(ref_states)):
            ts.set_roots_dev(roots.data_ptr(), n)
            ts.run(sims, 1.4, 32, _ffi.EVAL_CHESS_CRUDE, _ffi.POLICY_FIRST)
            res = np.zeros(n, dtype=np.int32)
            mv = np.zeros(n, dtype=_ffi.CHESS_MOVE_DTYPE)
            _ffi.check(_ffi.lib().zc_search_advance(ts._h, roots.data_ptr(), ones.data_ptr(), hist.data_ptr(), hlen.data_ptr(), 512,
                                                   res.ctypes.data_as(C.c_void_p), mv.ctypes.data_as(C.c_void_p), None))
            got = roots.cpu().numpy().view(eng.backend.STATE_DTYPE).reshape(n)[0]
            want = eng.backend.pack_state(ref_states[ply])
            assert got.tobytes() == want.tobytes(), ply
            host_result = eng._evaluate(ref_states[ply])
            assert int(res[0]) == (_ffi.RESULT_ONGOING if host_result is None else host_result), ply
        assert int(hlen.sum().item()) == n * (len(ref_states) - 1)


def test_device_selfplay_refill_and_network():
    from zeroclone_b200.games.connect4 import c4_backend
    from zeroclone_b200.policy_functions import Policy
    from zeroclone_b200.selfplay import DeviceSelfPlay
    from zeroclone_b200.value_functions import Value
    sp = DeviceSelfPlay(c4_backend, Value("network_latest", model_type="connect4_value"), Policy(), n_slots=16)
    out = sp.play(40, 32, 1.4, seed=3)
    assert len(out["results"]) == 40 and all(r in (-1, 0, 1) for r in out["results"])
    x, y = out["dataset"]
    assert x.shape[1:] == (2, 6, 7) and len(x) == len(y) == sum(len(t) for t in out["trajectories"])
    assert out["games_per_hour"] > 0


def test_engine_with_immediate_value_policy_on_device():
    """`policy_functions: immediate_value` (the key engine.py:27 actually reads) + policy.policy_freedom run on the GPU"""
    from engine.engine import Engine
    eng = Engine({"game": "chess", "backend": "chess_backend", "value_function": "crude_chess_score", "threads": 3,
                  "policy_functions": "immediate_value", "policy": {"policy_freedom": 3}, "mcts": {"simulations": 64, "c_puct": 1.4}})
    assert eng.policy.name == "immediate_value" and eng.policy.device_freedom == 3.0
    res = eng.play_mcts_parallel([0, 1, 2], 64, 1.4)
    assert set(res) == {0, 1, 2} and all(r is None for r in res.values())
    assert all(eng.get_state(i).turn == 1 for i in range(3))
    c4 = Engine({"game": "connect4", "backend": "c4_backend", "value_function": "c4_terminal", "threads": 2,
                 "policy_functions": "immediate_value"})
    assert c4.play_mcts(0, 64, 1.4) is None      # all Connect Four moves carry value 0: same as random


def test_shipped_chess_value_config_at_its_own_size():
    """configs/chess_value.yaml as shipped by the reference: 10 games, 50 000 simulations per move, network evaluator, random
    expansion policy.  One move for every game through the Engine API: no arena overflow, every root fully searched."""
    import os
    from zeroclone_b200.engine import Engine
    eng = Engine(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "configs", "chess_value.yaml"))
    sims = eng.config["mcts"]["simulations"]
    assert eng.threads == 10 and sims == 50000
    res = eng.play_mcts_parallel(range(eng.threads), simulations=sims, c=eng.config["mcts"]["c_puct"])
    assert set(res) == set(range(10)) and all(r is None for r in res.values())         # ten opening moves, games go on
    for i in range(10):
        st = eng.last_search_stats(i)
        assert int(st["visits"].sum()) == sims and st["best"] == int(st["visits"].argmax())
        assert len(eng.history[i].states) == 2


def test_timing_check_script_like_the_reference(tmp_path):
    """scripts/timing_check.py (reference: scripts/timing_check.py:15-58): same flags, same five printed lines"""
    import subprocess
    import sys
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for cfg, extra in (("configs/chess_value.yaml", []), ("configs/crude_chess.yaml", ["--trees", "64"])):
        out = subprocess.run([sys.executable, "scripts/timing_check.py", "-c", cfg, "--sims", "64", "--batch", "32", "--loops", "3", *extra],
                             cwd=repo, env=dict(os.environ, PYTHONPATH=repo), capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stderr[-2000:]
        lines = out.stdout.strip().splitlines()
        assert lines[0] == "--- get_move timing (3 runs) ---" and lines[1] == "simulations : 64" and lines[2] == "batch size  : 32"
        assert lines[3].startswith("mean  time  : ") and lines[4].startswith("median time : ")
        if extra:
            assert lines[5].startswith("batched     : 64 trees in ")
