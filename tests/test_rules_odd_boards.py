"""Hand-made boards far outside normal play (up to 33 pieces a side, two kings or none, pawns anywhere): the host rule code
(the same __host__ __device__ functions the kernels run), the oracle and -- where it was built -- the UNMODIFIED reference
module (oracle/_ref/chess_backend*.so) must list the same legal moves in the same order.  In particular a side WITHOUT a king:
the reference's find_king leaves (-1,-1) and its bounds-checked king_attacked tests that phantom square (chess_backend.cpp:68-144)."""
import ctypes as C

import numpy as np
import pytest

from oracle import ref_harness as rh
from oracle import zc_oracle as zo
from zeroclone_b200 import _ffi


def odd_boards(n=600, seed=7):
    rng = np.random.default_rng(seed)
    pieces_w, pieces_b = b"PNBRQK", b"pnbrqk"
    out = []
    for _ in range(n):
        board = bytearray(b" " * 64)
        n_w, n_b = int(rng.integers(1, 34)), int(rng.integers(1, 34))
        squares = rng.permutation(64)
        kings_w, kings_b = int(rng.integers(0, 3)), int(rng.integers(0, 3))
        k = 0
        for n_side, kings, alphabet in ((n_w, kings_w, pieces_w), (n_b, kings_b, pieces_b)):
            for j in range(min(n_side, 64 - k)):
                board[int(squares[k])] = alphabet[5] if j < kings else alphabet[int(rng.integers(0, 5))]
                k += 1
        out.append((bytes(board), int(rng.integers(0, 2))))
    return out


def test_host_rules_oracle_and_reference_agree_on_odd_boards():
    L = _ffi.lib()
    ref = None
    try:
        _, ref = rh.ref_modules()
    except Exception:
        pass
    kingless = 0
    for i, (b, turn) in enumerate(odd_boards()):
        st = zo.ChState()
        st.board[:] = b
        st.turn = turn
        want = [(m[0], float(m[1])) for m in zo.ch_legal(st)]
        s = _ffi.ChessState()
        s.board[:] = b
        s.turn = turn
        mv = (_ffi.ChessMove * 256)()
        k = L.zc_chess_legal_moves(C.byref(s), mv)
        got = [((mv[j].fr, mv[j].fc, mv[j].tr, mv[j].tc), float(mv[j].value)) for j in range(k)]
        assert got == want, i
        assert bool(L.zc_chess_check_win(C.byref(s))) == zo.ch_check_win(st), i
        if ref is not None:
            rs = ref.State(list(b), turn, 0, False, False, False, False, [], [])
            assert [(tuple(m[0]), float(m[1])) for m in ref.get_legal_moves(rs)] == got, i
            assert ref.check_win(rs) == bool(L.zc_chess_check_win(C.byref(s))), i
        kingless += (b"K" if turn == 0 else b"k") not in b
    assert kingless > 100
