"""Leaf evaluators: the reference's engine/value_functions.py surface -- `Value(name, **kwargs)`,
`value(state, **kw)`, `value.batch(states, backend=)` -- plus `device_spec()`, which tells the
batched CUDA search how to evaluate leaves for this Value:
   ("builtin", ZC_EVAL_*)   heuristic computed inside the search kernel, or
   ("network", NetEvaluator) one batched bf16 forward per search batch (fused tower kernel).

Names: random_rollout, crude_chess_score, network_latest, network_at_path (the reference's,
value_functions.py:35-129) and c4_terminal / c4_positional (deterministic parity evaluators,
SURVEY.md §8d).  The per-state host methods mirror the reference's semantics for callers that
evaluate single states; the search never goes through them.
"""
from __future__ import annotations

import os
import random
from typing import Optional

import numpy as np

from . import _ffi

_C4_W = (1, 2, 3, 4, 3, 2, 1)
_PIECE = {'P': 1, 'N': 3, 'B': 3, 'R': 5, 'Q': 9, 'p': -1, 'n': -3, 'b': -3, 'r': -5, 'q': -9}


class Value:
    BUILTIN = {"random_rollout": _ffi.EVAL_C4_ROLLOUT, "crude_chess_score": _ffi.EVAL_CHESS_CRUDE,
               "c4_terminal": _ffi.EVAL_C4_TERMINAL, "c4_positional": _ffi.EVAL_C4_POSITIONAL}

    def __init__(self, name, **kwargs):
        self.name = name
        self.init_args = kwargs
        self._net = None          # NetEvaluator, built on first device use
        self.model = None
        if not hasattr(self, str(name)):
            raise AttributeError(f"unknown value function {name!r}")
        init = getattr(self, f"init_{name}", None)
        if init is not None:
            init()

    # ------------------------------------------------------------------ reference surface
    def __call__(self, state, **kwargs):
        return getattr(self, self.name)(state, self.init_args | kwargs)

    def batch(self, states, **kwargs):
        if self.model is None:
            return [self(s, **kwargs) for s in states]
        import torch
        backend = kwargs["backend"]
        planes = torch.from_numpy(np.stack([backend.state_to_tensor(s) for s in states]))
        ev = self.evaluator()
        return ev(planes.to(ev.device, ev.dtype)).cpu().tolist()

    # ------------------------------------------------------------------ device side
    def device_spec(self, game: int):
        if self.model is not None:
            return ("network", self.evaluator())
        ev = self.BUILTIN.get(self.name)
        if ev is None:
            raise NotImplementedError(f"value function {self.name!r} has no device implementation")
        if (ev == _ffi.EVAL_CHESS_CRUDE) != (game == _ffi.GAME_CHESS):
            raise ValueError(f"value function {self.name!r} does not apply to this game")
        return ("builtin", ev)

    def evaluator(self):
        if self._net is None:
            import torch
            from .evaluator import NetEvaluator
            if not torch.cuda.is_available():
                raise RuntimeError("the neural evaluator runs on the GPU only (no CPU fallback)")
            dtypes = {"fp16": torch.float16, "f16": torch.float16, "float16": torch.float16, "bf16": torch.bfloat16, "bfloat16": torch.bfloat16}
            want = self.init_args.get("dtype")       # None: the evaluator's per-game default (bf16 Connect Four, fp16 chess)
            dt = None if want is None else dtypes.get(str(want))
            if want is not None and dt is None:
                raise ValueError("the neural evaluator is the fused sm_100a tower kernel: fp16 or bf16 operands, fp32 accumulation")
            dev = torch.device("cuda", self.init_args.get("device", torch.cuda.current_device()))
            self._net = NetEvaluator(self.model, dev, dt)
        return self._net

    def refresh(self) -> None:
        """self.model was trained further: push its weights into the resident tower (no rebuild, no re-allocation)"""
        if self._net is not None:
            self._net.update_weights(self.model.eval())

    # ------------------------------------------------------------------ heuristics (host, per state)
    def random_rollout(self, state, args):          # value_functions.py:35-45
        backend = args['backend']
        first = state.turn
        while not backend.check_win(state) and not backend.check_draw(state):
            state = backend.play_move(state, random.choice(list(backend.get_legal_moves(state))))
        if backend.check_win(state):
            return -1 if state.turn == first else 1
        return 0

    def crude_chess_score(self, state, args):       # value_functions.py:49-55 (mated side scores +1000, sic)
        if args['backend'].check_win(state):
            return 1000
        sign = 1 - 2 * state.turn
        return sign * sum(_PIECE.get(chr(p), 0) for p in state.board)

    def c4_terminal(self, state, args):
        return -1 if args['backend'].check_win(state) else 0

    def c4_positional(self, state, args):
        if args['backend'].check_win(state):
            return -1
        cur = 'XO'[state.turn]
        return sum((_C4_W[c] if cell == cur else -_C4_W[c]) for row in state.board for c, cell in enumerate(row) if cell != ' ') / 64

    # ------------------------------------------------------------------ networks
    def _load(self, path: Optional[os.PathLike]):
        import torch
        from .models import core
        module, latest = core.get_value_network(self.init_args['model_type'])
        module.add_safe_globals()
        path = latest if path is None else path
        if os.path.exists(path):
            # the safe loader (no arbitrary unpickling): whole-module pickles of this package's classes, of the
            # reference's classes (same sub-module names; registered under both paths), or a plain state_dict
            obj = torch.load(path, map_location="cpu", weights_only=True)
            if isinstance(obj, dict):
                self.model = module.ValueNetwork()
                self.model.load_state_dict(obj)
            else:
                self.model = obj
        else:
            self.model = module.ValueNetwork()       # absent checkpoint => random init (value_functions.py:110,125)
        self.model.eval()
        self.batch_size = self.init_args.get('batch_size', 1)

    def init_network_latest(self):
        self._load(None)

    def init_network_at_path(self):
        self._load(self.init_args['path'])

    def network_latest(self, state, args):
        return self.batch([state], **args)[0]

    def network_at_path(self, state, args):
        return self.batch([state], **args)[0]
