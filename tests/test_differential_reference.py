"""Differential tests of this package's backend modules against the UNMODIFIED reference modules (oracle/_ref: the compiled
chess_backend and the stock c4_backend.py from pyref.zip) on inputs far outside normal play: arbitrary moves fed to play_move
(two-file king moves, promotions, captures of anything), random histories for the repetition rule, hand-made Connect Four boards
with gaps and several fours, odd FEN strings.  Whatever the reference computes for such inputs is the contract."""
import importlib.util
import os

import numpy as np
import pytest

from oracle import ref_harness as rh
from test_rules_odd_boards import odd_boards

pytestmark = pytest.mark.skipif(not rh.ref_available(), reason="oracle/_ref not built (needs /root/reference)")


@pytest.fixture(scope="module")
def ref_chess():
    return rh.ref_modules()[1]


@pytest.fixture(scope="module")
def ref_c4():
    path = os.path.join(rh.pyref_dir(), "engine", "games", "connect4", "c4_backend.py")
    spec = importlib.util.spec_from_file_location("_ref_c4_backend_under_test", path)      # no clash with this repo's `engine` aliases
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def mine_state(cb, board, turn, fifty, flags, hw=(), hb=()):
    return cb.State(list(board), turn, fifty, *flags, list(hw), list(hb))


def same_state(a, r):
    return (list(a.board) == list(r.board) and a.turn == r.turn and a.fifty_move_rule_counter == r.fifty_move_rule_counter and
            [a.w_ck, a.w_cq, a.b_ck, a.b_cq] == [r.w_ck, r.w_cq, r.b_ck, r.b_cq] and
            [tuple(map(tuple, [m[0]])) + (float(m[1]),) for m in a.hist_white] == [tuple(map(tuple, [m[0]])) + (float(m[1]),) for m in r.hist_white] and
            [tuple(map(tuple, [m[0]])) + (float(m[1]),) for m in a.hist_black] == [tuple(map(tuple, [m[0]])) + (float(m[1]),) for m in r.hist_black])


def test_chess_play_move_on_arbitrary_moves(ref_chess):
    from zeroclone_b200.games.chess import chess_backend as cb
    rng = np.random.default_rng(11)
    boards = odd_boards(200, seed=3)
    n = 0
    for b, turn in boards:
        flags = [bool(x) for x in rng.integers(0, 2, 4)]
        fifty = int(rng.integers(0, 60))
        occupied = [i for i in range(64) if b[i] != 32]
        for _ in range(12):
            fr = int(rng.choice(occupied))
            to = int(rng.integers(0, 64))
            if rng.random() < 0.25:                       # two-file king-style hop along the row (the castling branch, :388-391)
                to = (fr // 8) * 8 + min(7, max(0, fr % 8 + int(rng.choice([-2, 2]))))
            mv = ((fr // 8, fr % 8, to // 8, to % 8), float(rng.integers(0, 10)))
            a = cb.play_move(mine_state(cb, b, turn, fifty, flags), mv)
            r = ref_chess.play_move(ref_chess.State(list(b), turn, fifty, *flags, [], []), mv)
            assert same_state(a, r), (b, turn, mv)
            n += 1
    assert n == 2400


def test_chess_check_draw_with_random_histories(ref_chess):
    from zeroclone_b200.games.chess import chess_backend as cb
    rng = np.random.default_rng(5)
    init = cb.create_init_state()
    hits = 0
    for trial in range(400):
        def hist():
            k = int(rng.integers(0, 14))
            base = [((int(rng.integers(0, 8)), int(rng.integers(0, 8)), int(rng.integers(0, 8)), int(rng.integers(0, 8))), float(rng.integers(0, 3)))
                    for _ in range(int(rng.integers(1, 4)))]
            if rng.random() < 0.6:
                return [base[i % len(base)] for i in range(k)]
            return [((int(rng.integers(0, 8)), 0, 1, 2), 0.0) for _ in range(k)]
        hw, hb = hist(), hist()
        fifty = int(rng.choice([0, 10, 49, 50, 51, 120]))
        a = cb.check_draw(cb.State(init.board, 0, fifty, True, True, True, True, hw, hb))
        r = ref_chess.check_draw(ref_chess.State(list(init.board), 0, fifty, True, True, True, True, hw, hb))
        assert a == r, (trial, fifty, hw, hb)
        hits += a
    assert 20 < hits < 380


def test_chess_tensor_and_fen_on_odd_inputs(ref_chess):
    from zeroclone_b200.games.chess import chess_backend as cb
    for b, turn in odd_boards(120, seed=9):
        for flags in ([True, False, True, False], [False, True, False, True]):
            a = cb.state_to_tensor(mine_state(cb, b, turn, 3, flags))
            r = np.asarray(ref_chess.state_to_tensor(ref_chess.State(list(b), turn, 3, *flags, [], [])))
            assert a.shape == r.shape == (17, 8, 8) and np.array_equal(a, r)
    fens = ["rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1", "8/8/8/8/8/8/8/K6k b - - 12 40", "r3k2r/8/8/8/8/8/8/R3K2R w Kq - 3 9",
            "4k3/PPPPPPPP/8/8/8/8/pppppppp/4K3 b - - 49 1", "QQQQQQQQ/QQQQQQQQ/8/8/8/8/qqqqqqqq/qqqqkqqq w - - 0 1"]
    for fen in fens:
        a, r = cb.state_from_fen(fen), ref_chess.state_from_fen(fen)
        assert same_state(a, r), fen
        assert [(tuple(m[0]), float(m[1])) for m in cb.get_legal_moves(a)] == [(tuple(m[0]), float(m[1])) for m in ref_chess.get_legal_moves(r)], fen


def test_connect4_on_hand_made_boards(ref_c4):
    from zeroclone_b200.games.connect4 import c4_backend as c4
    rng = np.random.default_rng(2)
    for trial in range(1500):
        p_empty = rng.random()
        rows = [[" " if rng.random() < p_empty else ("X" if rng.random() < 0.5 else "O") for _ in range(7)] for _ in range(6)]
        turn = int(rng.integers(0, 2))
        a, r = c4.State([row[:] for row in rows], turn), ref_c4.State([row[:] for row in rows], turn)
        assert list(c4.get_legal_moves(a)) == list(ref_c4.get_legal_moves(r)), rows          # same moves in the same (set) order
        assert c4.check_win(a) == ref_c4.check_win(r) and c4.check_draw(a) == ref_c4.check_draw(r), rows
        assert np.array_equal(c4.state_to_tensor(a), ref_c4.state_to_tensor(r))
        for col in range(7):                                # any column, full ones included (the reference then only flips the turn)
            na, nr = c4.play_move(a, (col, 0)), ref_c4.play_move(r, (col, 0))
            assert [list(x) for x in na.board] == [list(x) for x in nr.board] and na.turn == nr.turn, (rows, col)


def test_chess_random_games_every_function_every_ply(ref_chess):
    """Whole random games played in lockstep by both backends: at every ply the legal move lists (order and capture values),
    check_win, check_draw (incl. the fifty-ply counter and the repetition test on the accumulated histories), the state after
    the move and the network planes must be the same.  Also from odd starting boards and with castling flags cleared."""
    from zeroclone_b200.games.chess import chess_backend as cb
    rng = np.random.default_rng(2024)
    starts = [(list(cb.create_init_state().board), 0)] * 40 + odd_boards(60, seed=9)
    plies = games_over = 0
    for gi, (b, turn) in enumerate(starts):
        flags = [True] * 4 if gi < 40 else [bool(x) for x in rng.integers(0, 2, 4)]
        a = mine_state(cb, b, turn, 0, flags)
        r = ref_chess.State(list(b), turn, 0, *flags, [], [])
        for ply in range(160):
            ma, mr = cb.get_legal_moves(a), ref_chess.get_legal_moves(r)
            assert [(tuple(m[0]), float(m[1])) for m in ma] == [(tuple(m[0]), float(m[1])) for m in mr], (gi, ply)
            wa, wr = cb.check_win(a), ref_chess.check_win(r)
            da, dr = cb.check_draw(a), ref_chess.check_draw(r)
            assert (wa, da) == (wr, dr), (gi, ply)
            assert np.array_equal(cb.state_to_tensor(a), np.asarray(ref_chess.state_to_tensor(r))), (gi, ply)
            if wa or da or not ma:
                games_over += 1
                break
            k = int(rng.integers(len(ma)))
            a, r = cb.play_move(a, ma[k]), ref_chess.play_move(r, mr[k])
            assert same_state(a, r), (gi, ply)
            plies += 1
    assert plies > 3000 and games_over > 20, (plies, games_over)


def test_rest_wire_format_equals_the_stock_serializer():
    """server/main.py:14-36 `serialize_state` is the REST wire format: what the UNMODIFIED function makes of the reference's
    states must be what this repo's server makes of this repo's states (same keys, same values, same JSON)."""
    import json
    import subprocess
    import sys
    repo = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    fens = ["rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1", "r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R b Kq - 7 1",
            "8/8/8/1k6/8/8/4K3/5B2 w - - 49 1"]
    script = r'''
import json, sys, importlib
sys.path.insert(0, REPO)
if STOCK:
    from oracle import ref_harness as rh
    s = rh.stock(need_torch=True)
    chess, c4 = s.chess, s.c4
    import importlib.util, os
    spec = importlib.util.spec_from_file_location("_stock_server_main", os.path.join(rh.pyref_dir(), "server", "main.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)            # imports the stock engine.engine (sys.path: pyref first)
    ser = mod.serialize_state
    assert sys.modules["engine.engine"].__file__.startswith(rh.pyref_dir())
else:
    from zeroclone_b200.games.chess import chess_backend as chess
    from zeroclone_b200.games.connect4 import c4_backend as c4
    from server.main import serialize_state as ser
out = []
for fen in FENS:
    st = chess.state_from_fen(fen)
    out.append(ser(st))
    for k in (0, 3, 5):
        mv = chess.get_legal_moves(st)
        if mv:
            st = chess.play_move(st, mv[k % len(mv)])
            out.append(ser(st))
st = c4.create_init_state()
out.append(ser(st))
for col in (3, 3, 4, 0, 6):
    st = c4.play_move(st, (col, 0))
    out.append(ser(st))
print("WIRE " + json.dumps(out, sort_keys=True))
'''
    got = []
    for stock in (True, False):
        code = f"REPO, STOCK, FENS = {repo!r}, {stock!r}, {fens!r}\n" + script
        run = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=repo if not stock else "/tmp",
                             env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
        assert run.returncode == 0, run.stderr[-3000:]
        got.append(json.loads([l for l in run.stdout.splitlines() if l.startswith("WIRE ")][-1][5:]))
    assert len(got[0]) == len(got[1]) > 12
    for a, b in zip(got[1], got[0]):
        assert a == b, (a, b)
