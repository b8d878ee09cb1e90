"""ctypes front end of oracle/libzc_oracle.so -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product (zeroclone_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libzc_oracle.so")

GAME_C4, GAME_CHESS = 0, 1
EVAL_C4_TERMINAL, EVAL_C4_POSITIONAL, EVAL_CHESS_CRUDE, EVAL_EXTERNAL = 0, 1, 2, 3
POLICY_FIRST, POLICY_LAST = 0, 1


class C4State(C.Structure):
    _fields_ = [("cell", (C.c_char * 7) * 6), ("turn", C.c_int32)]


class ChState(C.Structure):
    _fields_ = [("board", C.c_uint8 * 64), ("turn", C.c_uint8), ("fifty", C.c_uint8), ("w_ck", C.c_uint8),
                ("w_cq", C.c_uint8), ("b_ck", C.c_uint8), ("b_cq", C.c_uint8), ("pad", C.c_uint8 * 2)]


class ChMove(C.Structure):
    _fields_ = [("fr", C.c_uint8), ("fc", C.c_uint8), ("tr", C.c_uint8), ("tc", C.c_uint8), ("val", C.c_float)]


class SearchResult(C.Structure):
    _fields_ = [("n_moves", C.c_int32), ("best", C.c_int32), ("Na", C.c_int32 * 256), ("Wa", C.c_double * 256),
                ("moves", (C.c_uint8 * 4) * 256), ("move_val", C.c_float * 256), ("nodes_created", C.c_int64),
                ("sum_leaf_depth", C.c_int64), ("max_leaf_depth", C.c_int64), ("reevaluated_leaves", C.c_int64),
                ("tree_hash", C.c_uint64), ("root_N", C.c_int32)]


BATCH_EVAL = C.CFUNCTYPE(None, C.POINTER(C.c_uint8), C.c_int, C.c_int, C.POINTER(C.c_double), C.c_void_p)


def build(force: bool = False) -> str:
    """Compile the port (gcc, < 1 s).  Building the checker is not using it."""
    src = os.path.join(_HERE, "zc_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "port"], check=True, stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.zo_ch_perft.restype = C.c_uint64
        L.zo_eval_state.restype = C.c_double
        L.zo_search.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, BATCH_EVAL,
                                C.c_void_p, C.POINTER(SearchResult)]
        assert L.zo_sizeof_c4_state() == C.sizeof(C4State)
        assert L.zo_sizeof_ch_state() == C.sizeof(ChState)
        assert L.zo_sizeof_search_result() == C.sizeof(SearchResult)
        _lib = L
        install_c4_order(python_c4_order())
    return _lib


# --------------------------------------------------------------------------- C4 move order
def python_c4_order() -> np.ndarray:
    """The 128x8 table of CPython set-iteration orders the reference's
    get_legal_moves (c4_backend.py:49-50) produces under THIS interpreter."""
    t = np.full((128, 8), 255, dtype=np.uint8)
    for mask in range(128):
        order = [m[0] for m in list({(i, 0) for i in range(7) if mask >> i & 1})]
        t[mask, :len(order)] = order
    return t


def install_c4_order(table: np.ndarray) -> None:
    t = np.ascontiguousarray(table, dtype=np.uint8)
    assert t.shape == (128, 8)
    lib().zo_c4_set_order(t.ctypes.data_as(C.c_void_p))


def baked_c4_order() -> np.ndarray:
    """The table compiled into the .so (CPython 3.12), read before any override."""
    L = C.CDLL(_LIB_PATH) if _lib is None else _lib
    t = np.zeros((128, 8), dtype=np.uint8)
    L.zo_c4_get_order(t.ctypes.data_as(C.c_void_p))
    return t


# --------------------------------------------------------------------------- C4 helpers
def c4_init() -> C4State:
    s = C4State()
    lib().zo_c4_init(C.byref(s))
    return s


def c4_from_rows(rows: Sequence[Sequence[str]], turn: int) -> C4State:
    s = C4State()
    for r in range(6):
        for c in range(7):
            s.cell[r][c] = rows[r][c].encode()
    s.turn = turn
    return s


def c4_rows(s: C4State) -> List[List[str]]:
    return [[s.cell[r][c].decode() for c in range(7)] for r in range(6)]


def c4_play(s: C4State, col: int) -> C4State:
    o = C4State()
    lib().zo_c4_play(C.byref(s), int(col), C.byref(o))
    return o


def c4_legal(s: C4State) -> List[int]:
    cols = (C.c_int32 * 8)()
    n = lib().zo_c4_legal(C.byref(s), cols)
    return [cols[i] for i in range(n)]


def c4_check_win(s: C4State) -> bool:
    return bool(lib().zo_c4_check_win(C.byref(s)))


def c4_check_draw(s: C4State) -> bool:
    return bool(lib().zo_c4_check_draw(C.byref(s)))


def c4_to_tensor(s: C4State) -> np.ndarray:
    out = np.zeros((2, 6, 7), dtype=np.float32)
    lib().zo_c4_to_tensor(C.byref(s), out.ctypes.data_as(C.c_void_p))
    return out


def c4_from_moves(cols: Sequence[int]) -> C4State:
    s = c4_init()
    for c in cols:
        s = c4_play(s, c)
    return s


# --------------------------------------------------------------------------- chess helpers
def ch_init() -> ChState:
    s = ChState()
    lib().zo_ch_init(C.byref(s))
    return s


def ch_from_fen(fen: str) -> ChState:
    s = ChState()
    lib().zo_ch_from_fen(fen.encode(), C.byref(s))
    return s


def ch_legal(s: ChState) -> List[tuple]:
    mv = (ChMove * 256)()
    n = lib().zo_ch_legal(C.byref(s), mv)
    return [((mv[i].fr, mv[i].fc, mv[i].tr, mv[i].tc), float(mv[i].val)) for i in range(n)]


def _mv(move) -> ChMove:
    (fr, fc, tr, tc), val = move
    return ChMove(fr, fc, tr, tc, float(val))


def ch_play(s: ChState, move) -> ChState:
    o = ChState()
    m = _mv(move)
    lib().zo_ch_play(C.byref(s), C.byref(m), C.byref(o))
    return o


def ch_check_win(s: ChState) -> bool:
    return bool(lib().zo_ch_check_win(C.byref(s)))


def ch_check_draw(s: ChState, hist_white: Sequence = (), hist_black: Sequence = ()) -> bool:
    hw = (ChMove * max(1, len(hist_white)))(*[_mv(m) for m in hist_white])
    hb = (ChMove * max(1, len(hist_black)))(*[_mv(m) for m in hist_black])
    return bool(lib().zo_ch_check_draw(C.byref(s), hw, len(hist_white), hb, len(hist_black)))


def ch_to_tensor(s: ChState) -> np.ndarray:
    out = np.zeros((17, 8, 8), dtype=np.float32)
    lib().zo_ch_to_tensor(C.byref(s), out.ctypes.data_as(C.c_void_p))
    return out


def ch_perft(s: ChState, depth: int) -> int:
    return int(lib().zo_ch_perft(C.byref(s), depth))


def ch_board_str(s: ChState) -> str:
    return bytes(s.board).decode()


def eval_state(evaluator: int, s) -> float:
    return float(lib().zo_eval_state(evaluator, C.byref(s)))


# --------------------------------------------------------------------------- search
@dataclass
class OracleSearch:
    n_moves: int
    best: int
    Na: List[int]
    Wa: List[float]
    moves: List[tuple]
    nodes_created: int
    sum_leaf_depth: int
    max_leaf_depth: int
    reevaluated_leaves: int
    tree_hash: int
    root_N: int


def search(game: int, state, simulations: int, c: float = 1.4, batch_size: int = 32, evaluator: int = 0,
           policy: int = POLICY_FIRST,
           external: Optional[Callable[[np.ndarray], np.ndarray]] = None) -> OracleSearch:
    """Run the restated reference search (mcts.cpp:102-160) on one root.

    `external(states_u8[n, state_bytes]) -> float64[n]` supplies leaf values when
    evaluator == EVAL_EXTERNAL (used for the neural evaluator)."""
    res = SearchResult()

    def _cb(ptr, n, sbytes, out, _user):
        arr = np.ctypeslib.as_array(ptr, shape=(n, sbytes)).copy()
        vals = np.asarray(external(arr), dtype=np.float64)
        for i in range(n):
            out[i] = vals[i]

    cb = BATCH_EVAL(_cb) if external is not None else BATCH_EVAL()
    rc = lib().zo_search(game, C.byref(state), simulations, c, batch_size, evaluator, policy, cb, None, C.byref(res))
    assert rc == 0
    n = res.n_moves
    if game == GAME_C4:
        moves = [(int(res.moves[i][0]), 0) for i in range(n)]
    else:
        moves = [((int(res.moves[i][0]), int(res.moves[i][1]), int(res.moves[i][2]), int(res.moves[i][3])),
                  float(res.move_val[i])) for i in range(n)]
    return OracleSearch(n, res.best, [res.Na[i] for i in range(n)], [res.Wa[i] for i in range(n)], moves,
                        res.nodes_created, res.sum_leaf_depth, res.max_leaf_depth, res.reevaluated_leaves,
                        res.tree_hash, res.root_N)


def search_puct(game: int, state, simulations: int, c: float = 1.4, batch_size: int = 32, evaluator: int = 0,
                virtual_loss: float = 1.0, prior_weight: int = 0, root_priors: Optional[Sequence[float]] = None,
                external: Optional[Callable[[np.ndarray], np.ndarray]] = None) -> OracleSearch:
    """The checker of the library's opt-in PUCT mode (zo_search_puct; NOT a reference behaviour: mcts.cpp is UCB1)."""
    res = SearchResult()

    def _cb(ptr, n, sbytes, out, _user):
        arr = np.ctypeslib.as_array(ptr, shape=(n, sbytes)).copy()
        vals = np.asarray(external(arr), dtype=np.float64)
        for i in range(n):
            out[i] = vals[i]

    cb = BATCH_EVAL(_cb) if external is not None else BATCH_EVAL()
    pri = None
    if root_priors is not None:
        pri = (C.c_float * len(root_priors))(*[float(x) for x in root_priors])
    L = lib()
    L.zo_search_puct.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_int, C.c_double, C.c_int, C.c_void_p,
                                 BATCH_EVAL, C.c_void_p, C.c_void_p]
    rc = L.zo_search_puct(game, C.byref(state), simulations, c, batch_size, evaluator, virtual_loss, prior_weight, pri, cb, None, C.byref(res))
    assert rc == 0
    n = res.n_moves
    if game == GAME_C4:
        moves = [(int(res.moves[i][0]), 0) for i in range(n)]
    else:
        moves = [((int(res.moves[i][0]), int(res.moves[i][1]), int(res.moves[i][2]), int(res.moves[i][3])),
                  float(res.move_val[i])) for i in range(n)]
    return OracleSearch(n, res.best, [res.Na[i] for i in range(n)], [res.Wa[i] for i in range(n)], moves,
                        res.nodes_created, res.sum_leaf_depth, res.max_leaf_depth, res.reevaluated_leaves,
                        res.tree_hash, res.root_N)
