"""Thin Python wrapper of the zc_search C-ABI handle: thousands of MCTS trees on one GPU.

This is the batched replacement of `engine.mcts.get_move` (engine/mcts/src/mcts.cpp:102-160,
bindings_mcts.cpp:9-11): one `TreeSearch` holds the node arenas of up to `max_trees` trees and
`run()` advances all of them.  Nothing here computes on the CPU.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _ffi
from ._ffi import (C4_STATE_DTYPE, CHESS_MOVE_DTYPE, CHESS_STATE_DTYPE, EVAL_C4_POSITIONAL, EVAL_C4_TERMINAL,
                   EVAL_CHESS_CRUDE, EVAL_EXTERNAL, GAME_C4, GAME_CHESS, POLICY_FIRST, POLICY_LAST, POLICY_RANDOM,
                   ROOT_RESULT_DTYPE, check, lib)

__all__ = ["TreeSearch", "TreeView", "c4_pack_rows", "c4_pack_cols", "c4_unpack_rows"]


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _stream_ptr(stream) -> Optional[int]:
    if stream is None:
        return None
    if isinstance(stream, int):
        return stream or None
    return int(stream.cuda_stream) or None  # torch.cuda.Stream


class TreeSearch:
    def __init__(self, game: int, max_trees: int, max_sims: int, device: int = 0, arena_slots_per_tree: int = 0):
        self._h = C.c_void_p()
        self.game, self.max_trees, self.max_sims, self.device = game, max_trees, max_sims, device
        check(lib().zc_search_create(game, device, max_trees, max_sims, arena_slots_per_tree, C.byref(self._h)))
        self.n_trees = 0
        self.state_dtype = C4_STATE_DTYPE if game == GAME_C4 else CHESS_STATE_DTYPE
        self.max_moves = 7 if game == GAME_C4 else _ffi.MAX_MOVES

    def close(self) -> None:
        if self._h:
            lib().zc_search_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def device_bytes(self) -> int:
        return int(lib().zc_search_device_bytes(self._h))

    # -- roots ---------------------------------------------------------------------------------
    def set_roots(self, states: np.ndarray, stream=None) -> None:
        """states: structured array (C4_STATE_DTYPE / CHESS_STATE_DTYPE) in host memory."""
        states = np.ascontiguousarray(states, dtype=self.state_dtype)
        check(lib().zc_search_set_roots(self._h, _ptr(states), len(states), _stream_ptr(stream)))
        self.n_trees = len(states)

    def set_roots_dev(self, dev_ptr: int, n: int, stream=None) -> None:
        check(lib().zc_search_set_roots_dev(self._h, C.c_void_p(dev_ptr), n, _stream_ptr(stream)))
        self.n_trees = n

    def set_mode(self, select: int = _ffi.SELECT_UCB1, virtual_loss: float = 1.0, prior_weight: int = 0) -> None:
        """Selection rule of the following searches: the reference's UCB1 (default, bit-exact) or the opt-in PUCT with stored
        priors and virtual loss (include/zc_b200.h, csrc/puct.cuh).  Call before set_roots."""
        check(lib().zc_search_set_mode(self._h, int(select), float(virtual_loss), int(prior_weight)))
        self.n_trees = 0

    def set_root_priors(self, priors: np.ndarray, stream=None) -> None:
        """PUCT mode: float32[n_trees, stride] priors of every root's moves (backend order), e.g. from a policy head"""
        priors = np.ascontiguousarray(priors, dtype=np.float32)
        check(lib().zc_search_set_root_priors(self._h, _ptr(priors), priors.shape[1], _stream_ptr(stream)))

    def set_policy_freedom(self, policy_freedom: float) -> None:
        """`policy_freedom` of Policy.immediate_value (policy_functions.py:16)"""
        check(lib().zc_search_set_policy_freedom(self._h, float(policy_freedom)))

    # -- fused search with a built-in evaluator ------------------------------------------------
    def run(self, simulations: int, c: float = 1.4, batch_size: int = 32, evaluator: int = EVAL_C4_TERMINAL,
            policy: int = POLICY_FIRST, seed: int = 0, stream=None) -> None:
        check(lib().zc_search_run(self._h, simulations, float(c), batch_size, evaluator, policy, seed, _stream_ptr(stream)))

    # -- split phase for an external (neural) evaluator -----------------------------------------
    def begin(self, simulations: int, c: float = 1.4, batch_size: int = 32, policy: int = POLICY_FIRST, seed: int = 0):
        check(lib().zc_search_begin(self._h, simulations, float(c), batch_size, policy, seed))

    def pending(self) -> int:
        return int(lib().zc_search_pending(self._h))

    def select(self, planes_ptr: int, plane_dtype: int = _ffi.PLANE_BF16, stream=None) -> None:
        check(lib().zc_search_select(self._h, C.c_void_p(planes_ptr), plane_dtype, _stream_ptr(stream)))

    def backprop(self, values_ptr: int, stream=None) -> None:
        check(lib().zc_search_backprop(self._h, C.c_void_p(values_ptr), _stream_ptr(stream)))

    def run_network(self, evaluator, simulations: int, c: float = 1.4, batch_size: int = 32,
                    policy: int = POLICY_FIRST, seed: int = 0) -> None:
        """get_move's loop with a neural evaluator (mcts.cpp:112-127 calling value.batch): per batch
        one select kernel, one forward over every tree's pending leaves, one backprop kernel.
        `evaluator(planes[B,C,H,W], out=values[B])` is any callable on device tensors (NetEvaluator)."""
        import torch

        dev = torch.device("cuda", self.device)
        dtype = getattr(evaluator, "dtype", torch.float32)
        code = {torch.bfloat16: _ffi.PLANE_BF16, torch.float32: _ffi.PLANE_F32, torch.float16: _ffi.PLANE_F16}[dtype]
        shape = (self.n_trees * batch_size,) + ((2, 6, 7) if self.game == GAME_C4 else (17, 8, 8))
        key = (shape, dtype)
        if getattr(self, "_planes_key", None) != key:
            self._planes = torch.empty(shape, dtype=dtype, device=dev)
            self._values = torch.empty(shape[0], dtype=torch.float32, device=dev)
            self._planes_key = key
        with torch.cuda.device(dev):
            stream = torch.cuda.current_stream().cuda_stream
            self.begin(simulations, c, batch_size, policy, seed)
            nvtx = torch.cuda.nvtx       # ranges for ncu / timeline filtering; a few ns each
            while self.pending() > 0:
                nvtx.range_push("zc.select")
                self.select(self._planes.data_ptr(), code, stream)
                nvtx.range_pop()
                nvtx.range_push("zc.evaluate")
                evaluator(self._planes, out=self._values)
                nvtx.range_pop()
                nvtx.range_push("zc.backprop")
                self.backprop(self._values.data_ptr(), stream)
                nvtx.range_pop()

    # -- readout ---------------------------------------------------------------------------------
    def results(self, stats: bool = True, stream=None, reuse: bool = False) -> dict:
        """Root readout (mcts.cpp:150-159).  reuse=True returns the per-tree result array owned by this
        object (overwritten by the next call) instead of a fresh one -- fresh pages cost more than the copy."""
        n = self.n_trees
        if reuse:
            if getattr(self, "_res_buf", None) is None or len(self._res_buf) < n:
                self._res_buf = np.zeros(self.max_trees, dtype=ROOT_RESULT_DTYPE)
            res = self._res_buf[:n]
        else:
            res = np.zeros(n, dtype=ROOT_RESULT_DTYPE)
        visits = wsum = moves = None
        stride = self.max_moves
        if stats:
            visits = np.zeros((n, stride), dtype=np.int32)
            wsum = np.zeros((n, stride), dtype=np.float64)
            moves = np.zeros((n, stride), dtype=CHESS_MOVE_DTYPE)
        check(lib().zc_search_results(self._h, _ptr(res), _ptr(visits), _ptr(wsum), _ptr(moves), stride, _stream_ptr(stream)))
        return {"result": res, "visits": visits, "value_sums": wsum, "moves": moves}

    def results_begin(self, stream=None) -> None:
        """Enqueue the root readout and its device->host copy without waiting (zc_search_results_begin): the next
        set_roots_dev / run can be enqueued at once.  Collect with results_end()."""
        check(lib().zc_search_results_begin(self._h, _stream_ptr(stream)))
        self._begun_n = self.n_trees

    def results_end(self) -> dict:
        n = getattr(self, "_begun_n", self.n_trees)
        if getattr(self, "_res_buf", None) is None or len(self._res_buf) < n:
            self._res_buf = np.zeros(self.max_trees, dtype=ROOT_RESULT_DTYPE)
        res = self._res_buf[:n]
        check(lib().zc_search_results_end(self._h, _ptr(res)))
        return {"result": res, "visits": None, "value_sums": None, "moves": None}

    def tree_hash(self, stream=None) -> np.ndarray:
        out = np.zeros(self.n_trees, dtype=np.uint64)
        check(lib().zc_search_tree_hash(self._h, _ptr(out), _stream_ptr(stream)))
        return out

    def read_tree(self, tree: int, stream=None) -> "TreeView":
        """Host copy of one tree's arena as a structured view (inspection / invariant checks)."""
        cap = 1 << 16
        while True:
            buf = np.zeros((cap, 4), dtype=np.uint32)
            used, ss = C.c_int64(), C.c_int32()
            check(lib().zc_search_read_tree(self._h, tree, _ptr(buf), cap, C.byref(used), C.byref(ss), _stream_ptr(stream)))
            if used.value <= cap:
                return TreeView(buf[:used.value], ss.value, self.game)
            cap = int(used.value)

    def counters(self, stream=None) -> dict:
        c = _ffi.Counters()
        check(lib().zc_search_get_counters(self._h, C.byref(c), _stream_ptr(stream)))
        return {k: int(getattr(c, k)) for k, _ in c._fields_}


class TreeView:
    """Decoded arena of one tree (layout: zeroclone_b200/csrc/tree.cuh)."""

    def __init__(self, slots: np.ndarray, state_slots: int, game: int):
        self.slots, self.ss, self.game = slots, state_slots, game

    K_UNKNOWN = 0xFFFF      # a stub: header + state only, moves not generated yet (tree.cuh)

    def node(self, slot: int) -> dict:
        h = self.slots[slot]
        k, nexp = int(h[1] & 0xFFFF), int(h[1] >> 16)
        stub = k == self.K_UNKNOWN
        if stub:
            k = 0
        e = self.slots[slot + 1 + self.ss: slot + 1 + self.ss + k]
        W = (e[:, 0].astype(np.uint64) | (e[:, 1].astype(np.uint64) << np.uint64(32))).view(np.float64) if k else np.zeros(0)
        return {"slot": slot, "N": int(h[0]), "k": k, "nexp": nexp, "stub": stub, "parent": int(h[2]), "parent_edge": int(h[3] & 0xFF),
                "depth": int(h[3] >> 16), "Na": e[:, 2].astype(np.int64), "Wa": W, "child": e[:, 3].astype(np.int64)}

    def walk(self):
        """every node, depth-first from the root (slot 0)"""
        stack = [0]
        while stack:
            n = self.node(stack.pop())
            yield n
            stack.extend(int(c) for c in n["child"] if c)

    def check_invariants(self, max_abs_value: float) -> int:
        """Structural invariants of the reference's statistics (mcts.cpp:80-100): returns the node count.
        - an edge has a child iff it was expanded; #children == n_expanded;
        - child.N == parent.Na[edge]  (every backprop through the child passes the edge);
        - non-root N == (evaluations of the node itself) + sum Na, where a node with moves is evaluated
          exactly once and a move-less node N times;  root N == sum Na;
        - |Wa| <= Na * max|value|; depth and parent links consistent."""
        count = 0
        for n in self.walk():
            count += 1
            has = n["child"] != 0
            assert int(has.sum()) == n["nexp"], (n["slot"], "children vs n_expanded")
            assert ((n["Na"] > 0) == has).all(), (n["slot"], "visited edge without child or child without visit")
            assert (np.abs(n["Wa"]) <= n["Na"] * max_abs_value + 1e-9).all(), (n["slot"], "value sum out of range")
            if n["stub"]:
                # evaluated once when it was created; the first descent that reaches it makes it a complete node
                assert n["N"] == 1 and n["slot"] != 0, (n["slot"], "stub with N != 1")
            elif n["slot"] == 0:
                assert n["N"] == int(n["Na"].sum()), "root N"
            elif n["k"] > 0:
                assert n["N"] == 1 + int(n["Na"].sum()), (n["slot"], "N != 1 + sum Na")
            else:
                assert n["N"] >= 1 and n["nexp"] == 0
            for e in np.nonzero(has)[0]:
                c = self.node(int(n["child"][e]))
                assert c["parent"] == n["slot"] and c["parent_edge"] == int(e) and c["depth"] == n["depth"] + 1
                assert c["N"] == int(n["Na"][e]), (n["slot"], int(e), "child N != edge Na")
        return count


# ------------------------------------------------------------------------------------ C4 packing
def c4_pack_rows(rows: Sequence[Sequence[str]], turn: int) -> tuple:
    """6x7 board of ' '/'X'/'O' (row 0 = top, c4_backend.py:11-12) -> (x_bits, o_bits, turn)."""
    x = o = 0
    for r in range(6):
        for c in range(7):
            cell = rows[r][c]
            if cell == 'X':
                x |= 1 << (c * 7 + (5 - r))
            elif cell == 'O':
                o |= 1 << (c * 7 + (5 - r))
    return x, o, turn


def c4_unpack_rows(x: int, o: int) -> list:
    rows = [[' '] * 7 for _ in range(6)]
    for r in range(6):
        for c in range(7):
            b = 1 << (c * 7 + (5 - r))
            if x & b:
                rows[r][c] = 'X'
            elif o & b:
                rows[r][c] = 'O'
    return rows


def c4_pack_cols(cols: Sequence[int]) -> tuple:
    """Position after playing `cols` from the empty board (X first)."""
    x = o = 0
    turn = 0
    for c in cols:
        occ = x | o
        for h in range(6):
            b = 1 << (c * 7 + h)
            if not occ & b:
                if turn == 0:
                    x |= b
                else:
                    o |= b
                break
        turn = 1 - turn
    return x, o, turn
