"""ctypes binding of libzc_b200.so (include/zc_b200.h).  Fails loudly: there is no CPU path."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ZC_B200_LIB") or os.path.join(_HERE, "libzc_b200.so")     # override: A/B timing of a library variant

GAME_C4, GAME_CHESS = 0, 1
EVAL_C4_TERMINAL, EVAL_C4_POSITIONAL, EVAL_CHESS_CRUDE, EVAL_EXTERNAL, EVAL_C4_ROLLOUT = 0, 1, 2, 3, 4
POLICY_FIRST, POLICY_LAST, POLICY_RANDOM, POLICY_IMMEDIATE_VALUE = 0, 1, 2, 3
ZC_OK, ZC_EINVAL, ZC_ENODEVICE, ZC_ECUDA, ZC_ECAPACITY, ZC_ESTATE = 0, -1, -2, -3, -4, -5
MAX_MOVES = 256
RESULT_ONGOING = 2
PLANE_BF16, PLANE_F32, PLANE_F16 = 0, 1, 2
SELECT_UCB1, SELECT_PUCT = 0, 1
ABI_VERSION = 2


class ZcError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libzc_b200 error {code}: {msg}")
        self.code = code


class C4State(C.Structure):
    _fields_ = [("x", C.c_uint64), ("o", C.c_uint64), ("turn", C.c_int32), ("reserved", C.c_int32)]


class ChessState(C.Structure):
    _fields_ = [("board", C.c_uint8 * 64), ("turn", C.c_uint8), ("fifty_move_rule_counter", C.c_uint8),
                ("w_ck", C.c_uint8), ("w_cq", C.c_uint8), ("b_ck", C.c_uint8), ("b_cq", C.c_uint8),
                ("reserved", C.c_uint8 * 2)]


class ChessMove(C.Structure):
    _fields_ = [("fr", C.c_uint8), ("fc", C.c_uint8), ("tr", C.c_uint8), ("tc", C.c_uint8), ("value", C.c_float)]


class RootResult(C.Structure):
    _fields_ = [("n_moves", C.c_int32), ("best", C.c_int32), ("root_visits", C.c_int32), ("status", C.c_int32),
                ("best_move", C.c_uint8 * 4), ("best_move_value", C.c_float), ("nodes", C.c_int32),
                ("sum_leaf_depth", C.c_int64), ("max_leaf_depth", C.c_int32), ("reevaluated_leaves", C.c_int32)]


class Counters(C.Structure):
    _fields_ = [("simulations", C.c_int64), ("nodes", C.c_int64), ("sum_leaf_depth", C.c_int64),
                ("sum_path_children", C.c_int64), ("arena_slots_used", C.c_int64), ("kernel_launches", C.c_int64)]


C4_STATE_DTYPE = np.dtype([("x", "<u8"), ("o", "<u8"), ("turn", "<i4"), ("reserved", "<i4")])
CHESS_STATE_DTYPE = np.dtype([("board", "u1", (64,)), ("turn", "u1"), ("fifty_move_rule_counter", "u1"), ("w_ck", "u1"),
                              ("w_cq", "u1"), ("b_ck", "u1"), ("b_cq", "u1"), ("reserved", "u1", (2,))])
CHESS_MOVE_DTYPE = np.dtype([("fr", "u1"), ("fc", "u1"), ("tr", "u1"), ("tc", "u1"), ("value", "<f4")])
ROOT_RESULT_DTYPE = np.dtype([("n_moves", "<i4"), ("best", "<i4"), ("root_visits", "<i4"), ("status", "<i4"),
                              ("best_move", "u1", (4,)), ("best_move_value", "<f4"), ("nodes", "<i4"), ("_pad", "<i4"),
                              ("sum_leaf_depth", "<i8"), ("max_leaf_depth", "<i4"), ("reevaluated_leaves", "<i4")])

_lib = None


def lib() -> C.CDLL:
    """Load the CUDA library.  Raises ImportError when it has not been built -- mirrors
    engine/mcts/__init__.py:3-9 of the reference ('mcts_cpp extension not built')."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"libzc_b200.so not built (expected {LIB_PATH}); run `python -m zeroclone_b200.build`. "
                          "There is no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, u64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_double
    L.zc_last_error.restype = C.c_char_p
    L.zc_search_device_bytes.restype = i64
    L.zc_search_device_bytes.argtypes = [vp]
    L.zc_c4_set_move_order.argtypes = [vp]
    L.zc_c4_get_move_order.argtypes = [vp]
    L.zc_search_create.argtypes = [i32, i32, i32, i32, i64, C.POINTER(vp)]
    L.zc_search_destroy.argtypes = [vp]
    L.zc_search_set_roots.argtypes = [vp, vp, i32, vp]
    L.zc_search_set_roots_dev.argtypes = [vp, vp, i32, vp]
    L.zc_search_run.argtypes = [vp, i32, dbl, i32, i32, i32, u64, vp]
    L.zc_search_set_policy_freedom.argtypes = [vp, dbl]
    L.zc_search_set_mode.argtypes = [vp, i32, dbl, i32]
    L.zc_search_set_root_priors.argtypes = [vp, vp, i32, vp]
    L.zc_search_begin.argtypes = [vp, i32, dbl, i32, i32, u64]
    L.zc_search_pending.argtypes = [vp]
    L.zc_search_select.argtypes = [vp, vp, i32, vp]
    L.zc_search_backprop.argtypes = [vp, vp, vp]
    L.zc_search_results.argtypes = [vp, vp, vp, vp, vp, i32, vp]
    L.zc_search_results_begin.argtypes = [vp, vp]
    L.zc_search_results_end.argtypes = [vp, vp]
    L.zc_search_tree_hash.argtypes = [vp, vp, vp]
    L.zc_search_get_counters.argtypes = [vp, vp, vp]
    L.zc_search_read_tree.argtypes = [vp, i32, vp, i64, vp, vp, vp]
    for name in ("zc_c4_init_state", "zc_c4_check_win", "zc_c4_check_draw", "zc_chess_init_state", "zc_chess_check_win"):
        getattr(L, name).argtypes = [vp]
    L.zc_c4_legal_moves.argtypes = [vp, vp]
    L.zc_c4_play_move.argtypes = [vp, i32, vp]
    L.zc_c4_to_tensor.argtypes = [vp, vp]
    L.zc_chess_from_fen.argtypes = [C.c_char_p, vp]
    L.zc_chess_legal_moves.argtypes = [vp, vp]
    L.zc_chess_play_move.argtypes = [vp, vp, vp]
    L.zc_chess_check_draw.argtypes = [vp, vp, i32, vp, i32]
    L.zc_chess_to_tensor.argtypes = [vp, vp]
    L.zc_chess_legal_moves_batch.argtypes = [i32, vp, i32, vp, vp, vp]
    L.zc_chess_legal_moves_batch_warp.argtypes = [i32, vp, i32, vp, vp, vp]
    L.zc_chess_perft.argtypes = [i32, vp, i32, vp]
    L.zc_c4_rules_batch.argtypes = [i32, vp, i32, vp, vp]
    L.zc_search_advance.argtypes = [vp, vp, vp, vp, vp, i32, vp, vp, vp]
    L.zc_states_to_tensor.argtypes = [i32, vp, i32, vp]
    L.zc_tower_create.argtypes = [i32, i32, i32, i32, vp, vp, vp, C.c_float, C.POINTER(vp)]
    L.zc_tower_plane_dtype.argtypes = [vp]
    L.zc_tower_destroy.argtypes = [vp]
    L.zc_tower_destroy.restype = None
    L.zc_tower_forward.argtypes = [vp, vp, i32, vp, vp]
    L.zc_tower_update_weights.argtypes = [vp, vp, vp, vp, C.c_float, vp]
    L.zc_tower_fault.argtypes = [vp]
    L.zc_tower_fault.restype = C.c_uint
    L.zc_tower_launches.argtypes = [vp]
    L.zc_tower_launches.restype = i64
    assert C.sizeof(RootResult) == ROOT_RESULT_DTYPE.itemsize == 48, (C.sizeof(RootResult), ROOT_RESULT_DTYPE.itemsize)
    assert C.sizeof(C4State) == C4_STATE_DTYPE.itemsize == 24
    assert C.sizeof(ChessState) == CHESS_STATE_DTYPE.itemsize == 72
    if L.zc_abi_version() != ABI_VERSION:
        raise ImportError("libzc_b200.so ABI version mismatch; rebuild")
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        raise ZcError(rc, lib().zc_last_error().decode())
