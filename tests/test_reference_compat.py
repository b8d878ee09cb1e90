"""The reference's own test-suite behaviours (tests/test_cb.py, tests/test_engine_configs.py in the
reference) exercised through the reference's import paths (`engine.*`, `models.*`), plus the Engine
API contract of SURVEY.md §8b.  CPU part here; the searches run under -m gpu."""
import glob
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import REPO


def test_setup_py_build_ext_inplace_is_accepted():
    for d in ("engine/games/chess", "engine/mcts"):
        subprocess.run([sys.executable, "setup.py", "build_ext", "--inplace"], cwd=os.path.join(REPO, d), check=True,
                       stdout=subprocess.DEVNULL)


def test_chess_backend_surface_like_test_cb(monkeypatch):
    from engine.games.chess import chess_backend as backend
    state = backend.create_init_state()
    moves = backend.get_legal_moves(state)
    assert isinstance(moves, list) and len(moves) == 20
    assert backend.check_win(state) is False and backend.check_draw(state) is False
    assert backend.play_move(state, moves[0]).turn == 1 - state.turn
    t = backend.state_to_tensor(state)
    assert t.ndim == 3 and t.shape[1:] == (8, 8) and t.dtype == np.float32
    # the module is a plain module with rebindable attributes (test_cb.py:55-59)
    monkeypatch.setattr(backend, "check_draw", lambda s: getattr(s, "fifty_move_rule_counter", 0) >= 50)
    white, black = {(7, 6, 5, 5), (5, 5, 7, 6)}, {(0, 6, 2, 5), (2, 5, 0, 6)}
    for ply in range(50):
        move = next((m for m in backend.get_legal_moves(state) if m[0] in (white if state.turn == 0 else black)), None)
        assert move is not None
        state = backend.play_move(state, move)
        assert backend.check_draw(state) == (ply == 49)


@pytest.mark.parametrize("fen,is_win,is_draw", [
    ("rnb1kbnr/pppp1ppp/8/4p3/6Pq/5P2/PPPPP2P/RNBQKBNR w KQkq - 0 1", True, False),
    ("r1bqkbnr/ppp2Qpp/n2p4/4p3/2B1P3/8/PPPP1PPP/RNB1K1NR b KQkq - 0 1", True, False),
    ("7k/5Q2/6K1/8/8/8/8/8 b - - 0 1", False, True),
    ("8/8/8/8/8/8/2n5/2K4k w - - 0 1", False, True),
    ("8/8/8/1k6/8/8/4K3/5B2 w - - 0 1", False, True)])
def test_fen_endings_like_test_cb(fen, is_win, is_draw):
    from engine.games.chess import chess_backend as backend
    s = backend.state_from_fen(fen)
    assert backend.check_win(s) is is_win and backend.check_draw(s) is is_draw


def test_chess_state_fields_are_read_write_and_repetition_draw():
    from engine.games.chess import chess_backend as backend
    s = backend.create_init_state()
    for f in ("board", "turn", "fifty_move_rule_counter", "w_ck", "w_cq", "b_ck", "b_cq", "hist_white", "hist_black"):
        assert hasattr(s, f)
    s.fifty_move_rule_counter = 50
    assert backend.check_draw(s)
    s = backend.State(board=s.board, turn=0, fifty_move_rule_counter=0, w_ck=True, w_cq=True, b_ck=True, b_cq=True,
                      hist_white=[], hist_black=[])
    # the g1f3/g8f6 knight shuffle is a repetition draw from ply 12
    white, black = [(7, 6, 5, 5), (5, 5, 7, 6)], [(0, 6, 2, 5), (2, 5, 0, 6)]
    first_draw = None
    for ply in range(14):
        want = (white if s.turn == 0 else black)[(ply // 2) % 2]
        s = backend.play_move(s, next(m for m in backend.get_legal_moves(s) if m[0] == want))
        if first_draw is None and backend.check_draw(s):
            first_draw = ply + 1
    assert first_draw == 12
    assert s.hist_white[0][0] in white and len(s.hist_white) == 7


def test_c4_backend_surface():
    from engine.games.connect4 import c4_backend as b
    s = b.create_init_state()
    assert s.turn == 0 and len(s.board) == 6 and len(s.board[0]) == 7
    mv = b.get_legal_moves(s)
    assert isinstance(mv, set) and mv == {(c, 0) for c in range(7)}
    s2 = b.play_move(s, (3, 0))
    assert s2.board[5][3] == 'X' and s2.turn == 1 and s.board[5][3] == ' '
    assert b.state_to_tensor(s2).shape == (2, 6, 7) and b.state_to_tensor(s2)[1, 5, 3] == 1.0
    for c in (3, 3, 3):                       # X X X X vertical would need alternation; build a row instead
        pass
    s = b.create_init_state()
    for c in (0, 0, 1, 1, 2, 2, 3):
        s = b.play_move(s, (c, 0))
    assert b.check_win(s) and not b.check_draw(s)


@pytest.mark.parametrize("cfg", sorted(glob.glob(os.path.join(REPO, "configs", "*.yaml"))))
def test_engine_constructs_from_every_config(cfg):
    from engine.engine import Engine
    eng = Engine(cfg)
    assert len(eng.legal_moves()) > 0
    assert eng.policy.name == "random"        # the reference's key typo (engine.py:27) is preserved
    assert len(eng.states) == eng.threads == len(eng.history)
    assert eng.values[0] is eng.values[1]


def test_engine_api_contract():
    from engine.engine import Engine
    eng = Engine({"game": "chess", "backend": "chess_backend", "value_function": "crude_chess_score", "threads": 2})
    assert eng.get_state(1).turn == 0
    idx = eng.add_game()
    assert idx == 2
    with pytest.raises(ValueError, match="Illegal move"):
        eng.play_move(((0, 0, 4, 4), 0.0), 0)
    assert eng.play_move(eng.legal_moves(0)[0], 0) is None
    assert eng.play_moves_parallel({1: eng.legal_moves(1)[3]}) == {1: None}
    assert len(eng.get_hist(0)) == 2
    # fool's mate: black wins => result -1, labels alternate backwards from -1 (engine.py:60-89)
    eng.reset_all_games()
    for want in [(6, 5, 5, 5), (1, 4, 3, 4), (6, 6, 4, 6), (0, 3, 4, 7)]:
        res = eng.play_move(next(m for m in eng.legal_moves(0) if m[0] == want), 0)
    assert res == -1
    x, y = eng.get_dataset()
    assert x.shape == (5, 17, 8, 8) and y.tolist() == [-1, 1, -1, 1, -1]
    with pytest.raises(ValueError):
        Engine({"game": "chess", "backend": "chess_backend", "value_function": "crude_chess_score"}, value_functions=[None])
    # Connect Four moves pass _is_legal (the reference raises TypeError here, engine.py:156)
    c4 = Engine(os.path.join(REPO, "configs", "connect4.yaml"))
    assert c4.play_move((3, 0)) is None


def test_models_contract_and_checkpoint_roundtrip(tmp_path):
    import torch
    import models.core as core
    module, latest = core.get_value_network("chess_value")
    assert latest.name == "latest.pth" and latest.parent.name == "chess_value"
    for name in ("ValueNetwork", "ValueNetDataset", "add_safe_globals", "train"):
        assert hasattr(module, name)
    net = module.ValueNetwork()
    assert sum(p.numel() for p in net.parameters()) == 2383361
    module.add_safe_globals()
    path = tmp_path / "latest.pth"
    torch.save(net, path)                                                # whole-module pickle, train.py:143
    back = torch.load(path, map_location="cpu")
    x = torch.randn(2, 17, 8, 8)
    net.eval(), back.eval()
    assert torch.equal(net(x), back(x))
    ds = module.ValueNetDataset(np.zeros((3, 17, 8, 8), np.float32), np.zeros(3, np.float32))
    assert len(ds) == 3 and ds[0][0].shape == (17, 8, 8)
    assert core.list_checkpoints("chess_value") == [] or all(p.suffix == ".pth" for p in core.list_checkpoints("chess_value"))


def test_get_move_rejects_foreign_callbacks_loudly():
    import engine.mcts as mcts
    from engine.games.connect4 import c4_backend as b
    with pytest.raises(TypeError, match="no CPU search"):
        mcts.get_move(b.create_init_state(), object(), lambda m: m[0], object(), 10, 1.4, 4)
