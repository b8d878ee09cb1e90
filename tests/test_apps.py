"""Callers either side of the hot path: REST wire format and the arena."""
import collections
import importlib.util
import os
import sys

import pytest

from conftest import REPO


def load(path, name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REPO, path))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def test_serialize_state_wire_format():
    srv = load("server/main.py", "zc_server_main")
    from zeroclone_b200.games.chess import chess_backend as cb
    from zeroclone_b200.games.connect4 import c4_backend as c4
    s = cb.create_init_state()
    s = cb.play_move(s, next(m for m in cb.get_legal_moves(s) if m[0] == (6, 4, 4, 4)))
    d = srv.serialize_state(s)
    assert len(d["board"]) == 64 and d["turn"] == 1 and d["fifty_move_rule_counter"] == 0
    assert d["w_ck"] is True or d["w_ck"] == 1
    # pybind11 hands std::deque over as a Python list, which the reference's serializer skips: the move
    # histories are not part of the wire format
    assert "hist_white" not in d and "hist_black" not in d
    c = c4.play_move(c4.create_init_state(), (3, 0))
    d = srv.serialize_state(c)
    assert d["turn"] == 1 and len(d["board"]) == 6 and d["board"][5][3] == "X"

    class Odd:
        board = 7
        name = "x"
        q = collections.deque([1, 2])
    assert srv.serialize_state(Odd()) == {"board": 7, "name": "x", "q": [1, 2]}


def test_rest_endpoints_host_side(tmp_path):
    from starlette.testclient import TestClient
    srv = load("server/main.py", "zc_server_main2")
    srv.config_path = os.path.join(REPO, "configs", "crude_chess.yaml")
    with TestClient(srv.app) as client:
        r = client.get("/legal_moves/0").json()
        assert r["idx"] == 0 and len(r["moves"]) == 20 and [6, 0, 5, 0] in r["moves"]
        r = client.post("/play_move", json={"idx": 0, "move": [6, 4, 4, 4]})
        assert r.status_code == 200
        body = r.json()
        assert body["result"] is None and body["turn"] == 1 and body["board"][4 * 8 + 4] == ord("P")
        bad = client.post("/play_move", json={"idx": 0, "move": [0, 0, 5, 5]})
        assert bad.status_code == 400 and "Illegal move" in bad.json()["detail"]
        idx = client.post("/add_game").json()["idx"]
        st = client.get(f"/state/{idx}").json()
        assert st["idx"] == idx and st["turn"] == 0
        assert client.get("/state/99").status_code == 400


def test_arena_win_rate_and_refill():
    ev = load("scripts/evaluate.py", "zc_evaluate")
    assert ev.win_rate([1, 1, -1, 0], True) == 0.625
    assert ev.win_rate([1, 1, -1, 0], False) == 0.375

    class FakeEngine:           # finishes every game on its third ply, result alternating
        threads = 2
        config = {"mcts": {"simulations": 1, "c_puct": 1.0}}

        def __init__(self):
            self.plies = [0, 0]

        def add_game(self):
            self.plies.append(0)
            return len(self.plies) - 1

        def play_mcts_parallel(self, idxs, sims, c):
            out = {}
            for i in idxs:
                self.plies[i] += 1
                out[i] = (1 if i % 2 == 0 else -1) if self.plies[i] == 3 else None
            return out
    res = ev.simulate(FakeEngine(), 5)
    assert res == [1, -1, 1, -1, 1]


@pytest.mark.gpu
def test_rest_play_mcts_and_arena_on_gpu():
    import torch
    from starlette.testclient import TestClient
    srv = load("server/main.py", "zc_server_main3")
    srv.config_path = os.path.join(REPO, "configs", "crude_chess.yaml")
    with TestClient(srv.app) as client:
        r = client.post("/play_mcts", json={"idx": 0, "simulations": 64, "c": 1.4})
        assert r.status_code == 200 and r.json()["turn"] == 1
    ev = load("scripts/evaluate.py", "zc_evaluate2")
    from zeroclone_b200.value_functions import Value
    torch.manual_seed(1)
    a = Value("network_latest", model_type="connect4_value")
    torch.manual_seed(2)
    b = Value("network_latest", model_type="connect4_value")
    cfg = {"game": "connect4", "backend": "c4_backend", "value_function": "network_latest", "threads": 4,
           "mcts": {"simulations": 32, "c_puct": 1.4}, "value": {"model_type": "connect4_value"}}
    wr = ev.evaluate_pair(cfg, a, b, 6)
    assert 0.0 <= wr <= 1.0
    assert a.evaluator() is not b.evaluator()      # two networks resident, one per side
