"""Synthetic root sets for benchmarks and tests (SURVEY.md §8d "Synthetic inputs").

Set A: every tree at the initial position (what Engine.__init__ does, engine.py:37-39).
Set B: from the initial position play r = tree_id mod 13 uniformly random legal plies with
       numpy.random.Generator(PCG64(1234 + tree_id)) indexing the backend-ordered move list,
       re-drawing the whole line if the result is terminal (check_win or check_draw).
Pure host-side bit twiddling on the packed root format of include/zc_b200.h.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _ffi

_COL0 = 0x3F
_FULL = sum(_COL0 << (7 * c) for c in range(7))


def _c4_order_table() -> np.ndarray:
    t = np.zeros((128, 8), dtype=np.uint8)
    _ffi.check(_ffi.lib().zc_c4_get_move_order(t.ctypes.data_as(C.c_void_p)))
    return t


def _four(b: int) -> bool:
    for s in (7, 1, 6, 8):
        m = b & (b >> s)
        if m & (m >> (2 * s)):
            return True
    return False


def c4_roots_set_a(n: int) -> np.ndarray:
    return np.zeros(n, dtype=_ffi.C4_STATE_DTYPE)


def c4_roots_set_b(n: int, first_tree_id: int = 0) -> np.ndarray:
    order = _c4_order_table()
    out = np.zeros(n, dtype=_ffi.C4_STATE_DTYPE)
    for i in range(n):
        tid = first_tree_id + i
        rng = np.random.Generator(np.random.PCG64(1234 + tid))
        plies = tid % 13
        while True:
            bb = [0, 0]
            turn = 0
            for _ in range(plies):
                occ = bb[0] | bb[1]
                mask = sum(1 << c for c in range(7) if not occ >> (7 * c + 5) & 1)
                moves = [int(x) for x in order[mask] if x != 255]
                col = moves[int(rng.integers(len(moves)))]
                empty = ~occ & (_COL0 << (7 * col))
                bb[turn] |= empty & -empty
                turn ^= 1
            just_moved = bb[turn ^ 1]
            if not _four(just_moved) and (bb[0] | bb[1]) != _FULL:
                break
        out[i] = (bb[0], bb[1], turn, 0)
    return out


def chess_roots_set_a(n: int) -> np.ndarray:
    s = _ffi.ChessState()
    _ffi.check(_ffi.lib().zc_chess_init_state(C.byref(s)))
    one = np.frombuffer(bytes(s), dtype=_ffi.CHESS_STATE_DTYPE)[0]
    out = np.zeros(n, dtype=_ffi.CHESS_STATE_DTYPE)
    out[:] = one
    return out


def chess_roots_set_b(n: int, first_tree_id: int = 0) -> np.ndarray:
    L = _ffi.lib()
    out = np.zeros(n, dtype=_ffi.CHESS_STATE_DTYPE)
    mv = (_ffi.ChessMove * _ffi.MAX_MOVES)()
    for i in range(n):
        tid = first_tree_id + i
        rng = np.random.Generator(np.random.PCG64(1234 + tid))
        plies = tid % 13
        while True:
            s = _ffi.ChessState()
            L.zc_chess_init_state(C.byref(s))
            hist = [[], []]
            for _ in range(plies):
                k = L.zc_chess_legal_moves(C.byref(s), mv)
                if k == 0:
                    break
                m = _ffi.ChessMove.from_buffer_copy(mv[int(rng.integers(k))])
                hist[s.turn].insert(0, m)
                nxt = _ffi.ChessState()
                L.zc_chess_play_move(C.byref(s), C.byref(m), C.byref(nxt))
                s = nxt
            hw = (_ffi.ChessMove * max(1, len(hist[0])))(*hist[0])
            hb = (_ffi.ChessMove * max(1, len(hist[1])))(*hist[1])
            if not L.zc_chess_check_win(C.byref(s)) and not L.zc_chess_check_draw(C.byref(s), hw, len(hist[0]), hb, len(hist[1])):
                break
        out[i] = np.frombuffer(bytes(s), dtype=_ffi.CHESS_STATE_DTYPE)[0]
    return out
