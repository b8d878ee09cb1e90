"""Connect Four backend module: the six-function interface of the reference's
engine/games/connect4/c4_backend.py (engine/README.md:13-28) with identical names, argument
meaning and return types -- `State(board, turn)` namedtuple with a 6x7 list of ' '/'X'/'O'
(row 0 = top), moves `(col, 0)`, `get_legal_moves` returns a set -- implemented over the
bitboard rule code of libzc_b200 (include/zc_b200.h: zc_c4_*).
"""
from __future__ import annotations

import ctypes as C
from collections import namedtuple

import numpy as np

from ... import _ffi

State = namedtuple('State', ['board', 'turn'])

ROWS, COLS = 6, 7
tokens = ['X', 'O']
ZC_GAME = _ffi.GAME_C4
STATE_DTYPE = _ffi.C4_STATE_DTYPE
TENSOR_SHAPE = (2, ROWS, COLS)


def _install_move_order() -> None:
    """get_legal_moves returns a *set* (c4_backend.py:49-50), so move order -- which decides MCTS
    tie-breaks -- is CPython's set-iteration order.  Derive it from the running interpreter and hand
    it to the library, exactly what the reference would produce under this Python."""
    t = np.full((128, 8), 255, dtype=np.uint8)
    for mask in range(128):
        order = [m[0] for m in list({(i, 0) for i in range(COLS) if mask >> i & 1})]
        t[mask, :len(order)] = order
    _ffi.check(_ffi.lib().zc_c4_set_move_order(t.ctypes.data_as(C.c_void_p)))


_install_move_order()


# ------------------------------------------------------------------ packed <-> State
def pack_state(state) -> tuple:
    x = o = 0
    board = state.board
    for r in range(ROWS):
        row = board[r]
        for c in range(COLS):
            cell = row[c]
            if cell == 'X':
                x |= 1 << (c * 7 + 5 - r)
            elif cell == 'O':
                o |= 1 << (c * 7 + 5 - r)
    return (x, o, int(state.turn), 0)


def unpack_state(x: int, o: int, turn: int) -> State:
    board = [[' '] * COLS for _ in range(ROWS)]
    for c in range(COLS):
        for h in range(ROWS):
            b = 1 << (c * 7 + h)
            if x & b:
                board[5 - h][c] = 'X'
            elif o & b:
                board[5 - h][c] = 'O'
    return State(board, turn)


def _c(state) -> _ffi.C4State:
    x, o, t, _ = pack_state(state)
    return _ffi.C4State(x, o, t, 0)


def move_from_result(best_move) -> tuple:
    """zc_root_result.best_move -> the backend's move object"""
    return (int(best_move[0]), 0)


# ------------------------------------------------------------------ the six functions
def create_init_state():
    return State([[' ' for _ in range(COLS)] for _ in range(ROWS)], 0)


def play_move(state, move):
    out = _ffi.C4State()
    _ffi.check(_ffi.lib().zc_c4_play_move(C.byref(_c(state)), int(move[0]), C.byref(out)))
    return unpack_state(out.x, out.o, out.turn)


def check_win(state):
    return bool(_ffi.lib().zc_c4_check_win(C.byref(_c(state))))


def check_draw(state):
    return bool(_ffi.lib().zc_c4_check_draw(C.byref(_c(state))))


def get_legal_moves(state):
    cols = (C.c_int32 * 8)()
    n = _ffi.lib().zc_c4_legal_moves(C.byref(_c(state)), cols)
    open_cols = {cols[i] for i in range(n)}
    # inserted in ascending column order like c4_backend.py:49-50: a set's iteration order depends on its insertion
    # history, and callers that iterate the set (engine.py's random move, the reference's tests) must see the same order
    return {(i, 0) for i in range(COLS) if i in open_cols}


def state_to_tensor(state):
    out = np.empty(TENSOR_SHAPE, dtype=np.float32)
    _ffi.check(_ffi.lib().zc_c4_to_tensor(C.byref(_c(state)), out.ctypes.data_as(C.c_void_p)))
    return out
