"""`Engine`: the reference's orchestration facade (engine/engine.py:15-157) with the same
constructor, attributes, methods, return values and exceptions, re-implemented over the batched
CUDA search.  `play_mcts_parallel(idxs, ...)` is ONE device search over all listed games instead of
a GIL-bound thread pool of per-game searches (engine.py:131-138).

Result convention (engine.py:148-153): +1 first player / white won, -1 second / black won,
0 draw, None game still running.
"""
from __future__ import annotations

import importlib
import threading
from dataclasses import dataclass, field
from typing import Any, Callable, Optional, Sequence

import numpy as np
import yaml

from . import mcts
from .policy_functions import Policy
from .value_functions import Value


@dataclass
class History:
    states: list = field(default_factory=list)
    result: Optional[int] = None


class Engine:
    def __init__(self, config: str | dict, *, value_functions: Sequence[Callable] | None = None):
        if isinstance(config, str):
            with open(config, "r") as fh:
                self.config = yaml.safe_load(fh)
        else:
            self.config = config
        self.backend = importlib.import_module(f"{__package__}.games.{self.config['game']}.{self.config['backend']}")
        # The reference reads the key `policy_functions` (engine.py:27) while its YAMLs write
        # `policy_function`, so every shipped config resolves to the default (random) policy.
        # Kept, so that the same YAML gives the same behaviour.
        self.policy = Policy(name=self.config.get("policy_functions"), **self.config.get("policy", {}))
        if value_functions is None:
            shared = Value(self.config.get("value_function"), **self.config.get("value", {}))
            self.values = [shared, shared]
        else:
            if len(value_functions) != 2:
                raise ValueError("value_functions must have length 2")
            self.values = list(value_functions)
        self.threads = self.config.get("threads", 1)
        self.batch_size = int(self.config.get("mcts", {}).get("batch_size", 32))   # extra key; reference default 32
        self.select_mode = mcts.select_mode(self.config.get("mcts"))               # extra keys; default = the reference's UCB1
        self.last_search: dict = {}
        # the REST server calls these methods from a thread pool: game bookkeeping and the shared device search
        # handle are used by one request at a time (the reference searched a private tree per request)
        self._lock = threading.RLock()
        self.reset_all_games()

    # ------------------------------------------------------------------ bookkeeping
    def add_game(self, init_state=None):
        state = init_state or self.backend.create_init_state()
        with self._lock:
            self.states.append(state)
            self.history.append(History(states=[state], result=None))
            return len(self.states) - 1

    def get_state(self, idx=0):
        return self.states[idx]

    def get_hist(self, idx=0):
        # the reference does list(History) here, which raises TypeError (engine.py:54-55);
        # returning the recorded states is what the name promises
        return list(self.history[idx].states)

    def reset_all_games(self):
        first = self.backend.create_init_state()
        self.states = [first for _ in range(self.threads)]
        self.history = [History(states=[first], result=None) for _ in range(self.threads)]

    def get_dataset(self):
        """(float32[N,C,H,W], float32[N]): every position of every finished game; the last position is
        labelled -1, alternating backwards, draws 0 (engine.py:60-89)."""
        planes, labels = [], []
        for h in self.history:
            if h.result is None:
                continue
            n = len(h.states)
            sign = 0 if h.result == 0 else -1
            planes.extend(self.backend.state_to_tensor(s).astype(np.float32) for s in h.states)
            labels.extend(sign * (-1) ** (n - 1 - i) for i in range(n))
        if not planes:
            shape = self.backend.state_to_tensor(self.backend.create_init_state()).shape
            return np.empty((0,) + shape, dtype=np.float32), np.empty((0,), dtype=np.float32)
        return np.stack(planes, axis=0), np.asarray(labels, dtype=np.float32)

    # ------------------------------------------------------------------ play
    def legal_moves(self, idx=0):
        return self.backend.get_legal_moves(self.states[idx])

    def play_move(self, move, idx=0):
        with self._lock:
            if not self._is_legal(move, idx):
                raise ValueError("Illegal move")
            nxt = self.backend.play_move(self.states[idx], move)
            self.states[idx] = nxt
            h = self.history[idx]
            h.states.append(nxt)
            h.result = self._evaluate(nxt)
            return h.result

    def play_moves_parallel(self, moves, max_workers=None):
        return {idx: self.play_move(mv, idx) for idx, mv in moves.items()}

    def play_mcts(self, idx=0, simulations=1000, c=1.4):
        return self.play_mcts_parallel([idx], simulations, c)[idx]

    def play_mcts_parallel(self, idxs, simulations=1000, c=1.4, max_workers=None):
        """For every listed game: root terminal guard, search, apply the chosen move
        (engine.py:119-129), with all searches in one batch on the GPU."""
        with self._lock:
            return self._play_mcts_parallel(idxs, simulations, c)

    def _play_mcts_parallel(self, idxs, simulations, c):
        results: dict = {}
        todo = {0: [], 1: []}
        for idx in idxs:
            state = self.states[idx]
            done = self._evaluate(state)
            if done is not None:
                self.history[idx].result = done
                results[idx] = done
            else:
                todo[state.turn].append(idx)
        groups = [todo[0] + todo[1]] if self.values[0] is self.values[1] else [todo[0], todo[1]]
        for group in groups:
            if not group:
                continue
            value_fn = self.values[self.states[group[0]].turn]
            out = mcts.search_batch([self.states[i] for i in group], value_fn, self.policy, self.backend,
                                    int(simulations), float(c), self.batch_size, stats=True, mode=self.select_mode)
            for j, idx in enumerate(group):
                self.last_search[idx] = {"visits": out["visits"][j][:out["result"][j]["n_moves"]].copy(),
                                         "value_sums": out["value_sums"][j][:out["result"][j]["n_moves"]].copy(),
                                         "best": int(out["result"][j]["best"])}
                results[idx] = self.play_move(out["moves_out"][j], idx)
        return results

    def last_search_stats(self, idx=0) -> dict:
        """Per-child visit counts / value sums of the last search of game idx (additive API: the
        reference's get_move returns only the move, mcts.cpp:157-159)."""
        return self.last_search[idx]

    # ------------------------------------------------------------------ helpers
    def _evaluate(self, state):
        if self.backend.check_win(state):
            return state.turn * 2 - 1
        if self.backend.check_draw(state):
            return 0
        return None

    def _is_legal(self, mv, idx=0) -> bool:
        # the reference unpacks four coordinates here (engine.py:156), which raises TypeError for
        # Connect Four's (col, 0) moves; comparing the first element works for both games
        return any(legal[0] == mv[0] for legal in self.legal_moves(idx))
