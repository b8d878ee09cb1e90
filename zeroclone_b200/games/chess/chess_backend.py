"""Chess backend module: same surface as the reference's pybind11 module `chess_backend`
(engine/games/chess/src/bindings_chess.cpp:13-57): class `State` with nine read/write fields and
the functions create_init_state, get_legal_moves, play_move, check_win, check_draw,
state_to_tensor, state_from_fen.  Moves are `((fr, fc, tr, tc), value)` tuples in a list.
Rules run in libzc_b200 (include/zc_b200.h: zc_chess_*); this file only marshals.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ... import _ffi

ZC_GAME = _ffi.GAME_CHESS
STATE_DTYPE = _ffi.CHESS_STATE_DTYPE
TENSOR_SHAPE = (17, 8, 8)


class State:
    """state.h:9-23 as bound at bindings_chess.cpp:13-41.  board: list of 64 ints (ASCII codes,
    32 = empty), index r*8+c, row 0 = rank 8; hist_*: that side's moves, most recent first."""
    __slots__ = ("board", "turn", "fifty_move_rule_counter", "w_ck", "w_cq", "b_ck", "b_cq", "hist_white", "hist_black")

    def __init__(self, board, turn, fifty_move_rule_counter, w_ck, w_cq, b_ck, b_cq, hist_white, hist_black):
        self.board = list(board)
        if len(self.board) != 64:
            raise TypeError("board must have 64 entries")
        self.turn = int(turn)
        self.fifty_move_rule_counter = int(fifty_move_rule_counter)
        self.w_ck, self.w_cq, self.b_ck, self.b_cq = bool(w_ck), bool(w_cq), bool(b_ck), bool(b_cq)
        self.hist_white = list(hist_white)
        self.hist_black = list(hist_black)

    def __repr__(self):
        return f"<chess State turn={self.turn} fifty={self.fifty_move_rule_counter} board={bytes(self.board).decode()!r}>"


def pack_state(state):
    """State -> one zc_chess_state record (numpy void of STATE_DTYPE)"""
    return np.frombuffer(bytes(_c(state)), dtype=STATE_DTYPE)[0]


def _c(state) -> _ffi.ChessState:
    s = _ffi.ChessState()
    C.memmove(s.board, bytes(state.board), 64)
    s.turn, s.fifty_move_rule_counter = state.turn, state.fifty_move_rule_counter & 0xFF
    s.w_ck, s.w_cq, s.b_ck, s.b_cq = state.w_ck, state.w_cq, state.b_ck, state.b_cq
    return s


def _from_c(s: _ffi.ChessState, hist_white, hist_black) -> State:
    return State(list(s.board), s.turn, s.fifty_move_rule_counter, s.w_ck, s.w_cq, s.b_ck, s.b_cq, hist_white, hist_black)


def _cm(move) -> _ffi.ChessMove:
    (fr, fc, tr, tc), val = move
    return _ffi.ChessMove(int(fr), int(fc), int(tr), int(tc), float(val))


def _hist(moves):
    n = len(moves)
    arr = (_ffi.ChessMove * max(1, n))()
    for i, m in enumerate(moves):
        arr[i] = _cm(m)
    return arr, n


def move_from_result(best_move, value) -> tuple:
    return ((int(best_move[0]), int(best_move[1]), int(best_move[2]), int(best_move[3])), float(value))


def create_init_state():
    s = _ffi.ChessState()
    _ffi.check(_ffi.lib().zc_chess_init_state(C.byref(s)))
    return _from_c(s, [], [])


def state_from_fen(fen):
    s = _ffi.ChessState()
    _ffi.check(_ffi.lib().zc_chess_from_fen(str(fen).encode(), C.byref(s)))
    return _from_c(s, [], [])


def get_legal_moves(state):
    mv = (_ffi.ChessMove * _ffi.MAX_MOVES)()
    n = _ffi.lib().zc_chess_legal_moves(C.byref(_c(state)), mv)
    if n < 0:
        _ffi.check(n)
    return [((mv[i].fr, mv[i].fc, mv[i].tr, mv[i].tc), float(mv[i].value)) for i in range(n)]


def play_move(state, move):
    out = _ffi.ChessState()
    m = _cm(move)
    _ffi.check(_ffi.lib().zc_chess_play_move(C.byref(_c(state)), C.byref(m), C.byref(out)))
    move = ((int(m.fr), int(m.fc), int(m.tr), int(m.tc)), float(m.value))
    hw, hb = list(state.hist_white), list(state.hist_black)
    (hw if state.turn == 0 else hb).insert(0, move)          # push_front, chess_backend.cpp:374
    return _from_c(out, hw, hb)


def check_win(state):
    return bool(_ffi.lib().zc_chess_check_win(C.byref(_c(state))))


def check_draw(state):
    hw, nw = _hist(state.hist_white)
    hb, nb = _hist(state.hist_black)
    rc = _ffi.lib().zc_chess_check_draw(C.byref(_c(state)), hw, nw, hb, nb)
    if rc < 0:
        _ffi.check(rc)
    return bool(rc)


def state_to_tensor(state):
    out = np.empty(TENSOR_SHAPE, dtype=np.float32)
    _ffi.check(_ffi.lib().zc_chess_to_tensor(C.byref(_c(state)), out.ctypes.data_as(C.c_void_p)))
    return out
