"""Expansion-order policies: the reference's engine/policy_functions.py:1-17 surface
(`Policy(name=None, **kwargs)`, callable on a move list) plus the id of the device policy the
batched search uses for it.  `first` / `last` are the deterministic policies used for parity."""
from __future__ import annotations

import random

from . import _ffi


class Policy:
    DEVICE = {"random": _ffi.POLICY_RANDOM, "first": _ffi.POLICY_FIRST, "last": _ffi.POLICY_LAST,
              "immediate_value": _ffi.POLICY_IMMEDIATE_VALUE}

    def __init__(self, name=None, **kwargs):
        self.name = name if name is not None else "random"
        self.args = kwargs
        if not hasattr(self, self.name):
            raise AttributeError(f"unknown policy {self.name!r}")

    def __call__(self, moves, **kwargs):
        return getattr(self, self.name)(moves, self.args | kwargs)

    @property
    def device_policy(self) -> int:
        """ZC_POLICY_* id for the CUDA search; raises for policies that only exist on the host."""
        try:
            return self.DEVICE[self.name]
        except KeyError:
            raise NotImplementedError(f"policy {self.name!r} has no device implementation (libzc_b200 has no CPU "
                                      "search to fall back to)") from None

    @property
    def device_freedom(self) -> float:
        """policy_freedom handed to the device search (immediate_value only)"""
        return float(self.args.get("policy_freedom", 0))

    # host semantics, one move list at a time (policy_functions.py:10-17)
    def random(self, moves, args):
        return random.choice(moves)

    def immediate_value(self, moves, args):
        best = max(m[1] for m in moves)
        slack = args.get('policy_freedom', 0)
        return random.choice([m for m in moves if m[1] >= best - slack])

    def first(self, moves, args):
        return moves[0]

    def last(self, moves, args):
        return moves[-1]
