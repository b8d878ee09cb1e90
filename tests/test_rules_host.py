"""CPU: the single-state rule helpers of the C-ABI (host execution of the same host+device rule
code the kernels run) against the golden vectors of the reference and the oracle."""
import ctypes as C

import numpy as np
import pytest

from conftest import load_golden
from oracle import zc_oracle as zo
from zeroclone_b200 import _ffi
from zeroclone_b200.build import build
from zeroclone_b200.search import c4_pack_rows, c4_unpack_rows


@pytest.fixture(scope="module")
def L():
    build()
    L = _ffi.lib()
    t = zo.python_c4_order()
    _ffi.check(L.zc_c4_set_move_order(t.ctypes.data_as(C.c_void_p)))
    return L


def chess_state(rec):
    s = _ffi.ChessState()
    for i, ch in enumerate(rec["board"]):
        s.board[i] = ord(ch)
    s.turn, s.fifty_move_rule_counter = rec["turn"], rec["fifty"]
    s.w_ck, s.w_cq, s.b_ck, s.b_cq = rec["flags"]
    return s


def legal(L, s):
    mv = (_ffi.ChessMove * 256)()
    n = L.zc_chess_legal_moves(C.byref(s), mv)
    assert n >= 0
    return [[mv[i].fr, mv[i].fc, mv[i].tr, mv[i].tc, float(mv[i].value)] for i in range(n)]


def to_move(m):
    return _ffi.ChessMove(m[0], m[1], m[2], m[3], float(m[4]))


def hist_arr(h):
    return (_ffi.ChessMove * max(1, len(h)))(*[to_move(m) for m in h])


def test_chess_fens_against_reference(L):
    g = load_golden("chess_rules.json.gz")
    for name, rec in g["fens"].items():
        s = _ffi.ChessState()
        _ffi.check(L.zc_chess_from_fen(rec["fen"].encode(), C.byref(s)))
        assert bytes(s.board).decode() == rec["board"] and s.turn == rec["turn"], name
        assert s.fifty_move_rule_counter == rec["fifty"], name
        assert [s.w_ck, s.w_cq, s.b_ck, s.b_cq] == rec["flags"], name
        assert legal(L, s) == rec["legal"], name
        assert bool(L.zc_chess_check_win(C.byref(s))) == rec["win"], name
        assert bool(L.zc_chess_check_draw(C.byref(s), None, 0, None, 0)) == rec["draw"], name
        t = np.zeros((17, 64), dtype=np.float32)
        _ffi.check(L.zc_chess_to_tensor(C.byref(s), t.ctypes.data_as(C.c_void_p)))
        assert [int(sum(1 << i for i in range(64) if t[p, i] == 1)) for p in range(17)] == rec["tensor_planes"], name


def test_chess_playouts_against_reference(L):
    g = load_golden("chess_rules.json.gz")
    draws = 0
    for trace in g["playouts"]:
        s = chess_state(trace[0])
        hist = [[], []]
        for rec in trace:
            assert bytes(s.board).decode() == rec["board"] and s.turn == rec["turn"]
            assert s.fifty_move_rule_counter == rec["fifty"]
            assert [s.w_ck, s.w_cq, s.b_ck, s.b_cq] == rec["flags"]
            assert legal(L, s) == rec["legal"]
            assert bool(L.zc_chess_check_win(C.byref(s))) == rec["win"]
            d = L.zc_chess_check_draw(C.byref(s), hist_arr(hist[0]), len(hist[0]), hist_arr(hist[1]), len(hist[1]))
            assert bool(d) == rec["draw"]
            draws += rec["draw"]
            if "tensor_planes" in rec:
                t = np.zeros((17, 64), dtype=np.float32)
                _ffi.check(L.zc_chess_to_tensor(C.byref(s), t.ctypes.data_as(C.c_void_p)))
                assert [int(sum(1 << i for i in range(64) if t[p, i] == 1)) for p in range(17)] == rec["tensor_planes"]
            if "played" in rec:
                m = rec["played"]
                hist[s.turn].insert(0, m)
                o = _ffi.ChessState()
                mv = to_move(m)
                _ffi.check(L.zc_chess_play_move(C.byref(s), C.byref(mv), C.byref(o)))
                s = o
    assert draws > 10


def test_chess_reference_test_vectors(L):
    # the reference's tests/test_cb.py:39-51,105-116 through the C-ABI
    s = _ffi.ChessState()
    _ffi.check(L.zc_chess_init_state(C.byref(s)))
    assert len(legal(L, s)) == 20 and not L.zc_chess_check_win(C.byref(s))
    for fen, win, draw in [
        ("rnb1kbnr/pppp1ppp/8/4p3/6Pq/5P2/PPPPP2P/RNBQKBNR w KQkq - 0 1", 1, 0),
        ("r1bqkbnr/ppp2Qpp/n2p4/4p3/2B1P3/8/PPPP1PPP/RNB1K1NR b KQkq - 0 1", 1, 0),
        ("7k/5Q2/6K1/8/8/8/8/8 b - - 0 1", 0, 1), ("8/8/8/8/8/8/2n5/2K4k w - - 0 1", 0, 1),
        ("8/8/8/1k6/8/8/4K3/5B2 w - - 0 1", 0, 1)]:
        _ffi.check(L.zc_chess_from_fen(fen.encode(), C.byref(s)))
        assert L.zc_chess_check_win(C.byref(s)) == win
        assert L.zc_chess_check_draw(C.byref(s), None, 0, None, 0) == draw


def test_chess_castling_move_fed_in_hops_the_rook(L):
    # chess_backend.cpp:388-391: never generated, but play_move executes it
    s = _ffi.ChessState()
    _ffi.check(L.zc_chess_from_fen(b"r3k2r/8/8/8/8/8/8/R3K2R w KQkq - 0 1", C.byref(s)))
    o = _ffi.ChessState()
    for mv, oracle_mv in [((7, 4, 7, 6), ((7, 4, 7, 6), 0.0)), ((7, 4, 7, 2), ((7, 4, 7, 2), 0.0))]:
        m = _ffi.ChessMove(*mv, 0.0)
        _ffi.check(L.zc_chess_play_move(C.byref(s), C.byref(m), C.byref(o)))
        ref = zo.ch_play(zo.ch_from_fen("r3k2r/8/8/8/8/8/8/R3K2R w KQkq - 0 1"), oracle_mv)
        assert bytes(o.board).decode() == zo.ch_board_str(ref)
        assert [o.w_ck, o.w_cq, o.b_ck, o.b_cq] == [ref.w_ck, ref.w_cq, ref.b_ck, ref.b_cq]


def test_chess_random_positions_against_oracle(L):
    rng = np.random.default_rng(17)
    for g in range(10):
        s = _ffi.ChessState()
        _ffi.check(L.zc_chess_init_state(C.byref(s)))
        o = zo.ch_init()
        for ply in range(150):
            mine, theirs = legal(L, s), zo.ch_legal(o)
            assert mine == [list(m[0]) + [m[1]] for m in theirs]
            assert bool(L.zc_chess_check_win(C.byref(s))) == zo.ch_check_win(o)
            if not mine:
                break
            k = int(rng.integers(len(mine)))
            n = _ffi.ChessState()
            mv = to_move(mine[k])
            _ffi.check(L.zc_chess_play_move(C.byref(s), C.byref(mv), C.byref(n)))
            s, o = n, zo.ch_play(o, theirs[k])
            assert bytes(s.board).decode() == zo.ch_board_str(o) and s.fifty_move_rule_counter == o.fifty


def test_c4_rules_against_reference_playouts(L):
    for trace in load_golden("c4_rules.json.gz"):
        s = _ffi.C4State()
        _ffi.check(L.zc_c4_init_state(C.byref(s)))
        for rec in trace:
            assert ["".join(r) for r in c4_unpack_rows(s.x, s.o)] == rec["rows"] and s.turn == rec["turn"]
            cols = (C.c_int32 * 8)()
            n = L.zc_c4_legal_moves(C.byref(s), cols)
            assert [cols[i] for i in range(n)] == rec["legal"]
            assert bool(L.zc_c4_check_win(C.byref(s))) == rec["win"]
            assert bool(L.zc_c4_check_draw(C.byref(s))) == rec["draw"]
            t = np.zeros(84, dtype=np.float32)
            _ffi.check(L.zc_c4_to_tensor(C.byref(s), t.ctypes.data_as(C.c_void_p)))
            assert t.astype(int).tolist() == rec["tensor"]
            if "played" in rec:
                o = _ffi.C4State()
                _ffi.check(L.zc_c4_play_move(C.byref(s), rec["played"], C.byref(o)))
                s = o


def test_c4_boards_with_gaps_follow_the_reference(L):
    # hand-made board with a floating disc: the disc falls to the LOWEST empty cell (c4_backend.py:18-21)
    rows = [list("       ") for _ in range(6)]
    rows[2][3] = 'X'
    x, o, t = c4_pack_rows(rows, 1)
    s = _ffi.C4State(x, o, t, 0)
    n = _ffi.C4State()
    _ffi.check(L.zc_c4_play_move(C.byref(s), 3, C.byref(n)))
    ref = zo.c4_play(zo.c4_from_rows(rows, 1), 3)
    assert c4_unpack_rows(n.x, n.o) == zo.c4_rows(ref) and n.turn == ref.turn


def test_treeview_decodes_stub_leaves():
    """host-side decoding of the arena layout (tree.cuh): a chess leaf is a stub (k = 0xFFFF, header + state)"""
    from zeroclone_b200.search import TreeView
    ss = 2
    slots = np.zeros((16, 4), dtype=np.uint32)
    # root at slot 0: N = 1, k = 2 moves, 1 expanded; edges at slots 3, 4; moves at slot 5
    slots[0] = (1, 2 | (1 << 16), 0, 0)
    w = np.array([-3.0], dtype=np.float64).view(np.uint32)
    slots[3] = (w[0], w[1], 1, 6)            # edge 0: Wa = -3, Na = 1, child at slot 6
    # stub child at slot 6: N = 1, k unknown, parent 0, parent edge 0, depth 1
    slots[6] = (1, 0xFFFF, 0, 0 | (1 << 16))
    tv = TreeView(slots, ss, _ffi.GAME_CHESS)
    child = tv.node(6)
    assert child["stub"] and child["k"] == 0 and child["N"] == 1 and child["depth"] == 1
    assert tv.check_invariants(max_abs_value=1000.0) == 2
    slots[6][0] = 2                           # a stub is evaluated exactly once
    with pytest.raises(AssertionError):
        TreeView(slots, ss, _ffi.GAME_CHESS).check_invariants(max_abs_value=1000.0)
