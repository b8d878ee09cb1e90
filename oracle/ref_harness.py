"""TEST/BENCH INFRASTRUCTURE: drive the UNMODIFIED reference search (oracle/_ref/mcts*.so,
compiled from the reference's engine/mcts/src/*.cpp by oracle/Makefile) on the box's host cores.

The reference's Connect Four backend and its Value/Policy classes are Python files that cannot
travel to the GPU box, so the callbacks the C++ search needs are restated here, with the same
object shapes the reference uses (namedtuple state with list-of-lists board, set-valued
get_legal_moves, value.batch(states, backend=), policy(list)).  Used only by bench.py
(--impl reference and the cpu_baseline leg) and by tests.
"""
from __future__ import annotations

import os
import sys
from collections import namedtuple

import numpy as np

_REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")


def ref_available() -> bool:
    return os.path.isdir(_REF_DIR) and any(f.startswith("mcts") and f.endswith(".so") for f in os.listdir(_REF_DIR))


def ref_modules():
    if _REF_DIR not in sys.path:
        sys.path.insert(0, _REF_DIR)
    import chess_backend
    import mcts
    return mcts, chess_backend


# ---- Connect Four backend with the reference's object shapes (c4_backend.py:4-61) ------------
C4State = namedtuple("State", ["board", "turn"])
_TOK = "XO"


class C4Backend:
    @staticmethod
    def create_init_state():
        return C4State([[" "] * 7 for _ in range(6)], 0)

    @staticmethod
    def from_bits(x: int, o: int, turn: int):
        b = [[" "] * 7 for _ in range(6)]
        for r in range(6):
            for c in range(7):
                bit = 1 << (c * 7 + (5 - r))
                if x & bit:
                    b[r][c] = "X"
                elif o & bit:
                    b[r][c] = "O"
        return C4State(b, turn)

    @staticmethod
    def play_move(state, move):
        col = move[0]
        board = [list(r) for r in state.board]
        for r in range(5, -1, -1):
            if board[r][col] == " ":
                board[r][col] = _TOK[state.turn]
                break
        return C4State(board, 1 - state.turn)

    @staticmethod
    def check_win(state):
        t = _TOK[1 - state.turn]
        b = state.board
        for r in range(6):
            for c in range(7):
                if b[r][c] != t:
                    continue
                for dr, dc in ((0, 1), (1, 0), (1, 1), (-1, 1)):
                    rr, cc = r + 3 * dr, c + 3 * dc
                    if 0 <= rr < 6 and 0 <= cc < 7 and all(b[r + i * dr][c + i * dc] == t for i in (1, 2, 3)):
                        return True
        return False

    @staticmethod
    def check_draw(state):
        return all(cell != " " for row in state.board for cell in row)

    @staticmethod
    def get_legal_moves(state):
        return {(c, 0) for c in range(7) if state.board[0][c] == " "}

    @staticmethod
    def state_to_tensor(state):
        cur, opp = _TOK[state.turn], _TOK[1 - state.turn]
        a = np.array(state.board)
        return np.stack([(a == cur), (a == opp)]).astype(np.float32)


class TorchValue:
    """value.batch(states, backend=) as engine/value_functions.py:24-32,78-99 compute it (stack the
    state tensors, one fp32 forward, list of floats), minus the worker thread and queues."""

    def __init__(self, model):
        import torch
        self.torch = torch
        self.model = model.eval()

    def batch(self, states, backend=None):
        x = self.torch.from_numpy(np.stack([backend.state_to_tensor(s) for s in states]))
        with self.torch.no_grad():
            return self.model(x).view(-1).tolist()


class FnValue:
    def __init__(self, fn):
        self.fn = fn

    def batch(self, states, backend=None):
        return [self.fn(s, backend) for s in states]


def first_policy(moves):
    return moves[0]
