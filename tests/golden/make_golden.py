#!/usr/bin/env python
"""Generate the golden parity fixtures by running the UNMODIFIED reference.

Runs only in the dev container (needs /root/reference and `make -C oracle ref`):
  * engine/mcts/src/mcts.cpp and engine/games/chess/src/chess_backend.cpp are the reference's
    own sources compiled by oracle/Makefile into oracle/_ref/ (nothing is copied),
  * engine/games/connect4/c4_backend.py, engine/value_functions.py and
    models/chess_value/network.py are imported straight from /root/reference.

`mcts.get_move` returns only the chosen move, so per-child visit counts and value sums are
recovered with the shadow tracer of SURVEY.md App. E: one proxy object is passed as backend,
value and policy; it forwards to the real backend, remembers child->(parent, move) by object
identity in play_move, and replays mcts.cpp:80-100 on the leaf lists handed to value.batch.

Usage:  python tests/golden/make_golden.py            (rewrites tests/golden/*.json[.gz])
The GPU box never runs this; it only reads the committed JSON.
"""
from __future__ import annotations

import gzip
import importlib.util
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.abspath(os.path.join(HERE, "..", ".."))
REF = os.environ.get("ZC_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(REPO, "oracle", "_ref"))

import chess_backend as ref_chess  # noqa: E402  (reference C++ built by oracle/Makefile)
import mcts as ref_mcts  # noqa: E402


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


ref_c4 = _load(os.path.join(REF, "engine/games/connect4/c4_backend.py"), "ref_c4_backend")
ref_values = _load(os.path.join(REF, "engine/value_functions.py"), "ref_value_functions")

C4_W = [1, 2, 3, 4, 3, 2, 1]


# ------------------------------------------------------------------ evaluators (reference side)
def c4_terminal(state, backend):
    return -1 if backend.check_win(state) else 0  # terminal branch of random_rollout, value_functions.py:41-43


def c4_positional(state, backend):
    if backend.check_win(state):
        return -1
    cur = ref_c4.tokens[state.turn]
    acc = 0
    for row in state.board:
        for c, cell in enumerate(row):
            if cell != ' ':
                acc += C4_W[c] if cell == cur else -C4_W[c]
    return acc / 64


_crude = ref_values.Value("crude_chess_score")


def chess_crude(state, backend):
    return _crude(state, backend=backend)  # value_functions.py:49-55, unmodified


EVALS = {"c4_terminal": c4_terminal, "c4_positional": c4_positional, "chess_crude": chess_crude}
POLICIES = {"first": lambda moves: moves[0], "last": lambda moves: moves[-1]}


# ------------------------------------------------------------------ shadow tracer (App. E)
class Shadow:
    def __init__(self, backend, evaluator, policy):
        self.b, self.ev, self.pol = backend, evaluator, policy
        self.parent = {}   # id(child) -> (id(parent), move index)
        self.keep = []     # keep states alive so ids stay unique
        self.N, self.Na, self.Wa = {}, {}, {}
        self.moves_of = {}
        self.leaf_depths = []
        self.depth = {}

    # backend interface
    def get_legal_moves(self, s):
        mv = list(self.b.get_legal_moves(s))
        self.moves_of[id(s)] = mv
        self.keep.append(s)
        return mv

    def play_move(self, s, m):
        c = self.b.play_move(s, m)
        self.keep.append(c)
        self.parent[id(c)] = (id(s), self.moves_of[id(s)].index(m))
        self.depth[id(c)] = self.depth.get(id(s), 0) + 1
        return c

    def __getattr__(self, k):
        return getattr(self.b, k)

    # policy interface
    def __call__(self, moves):
        return self.pol(moves)

    # value interface
    def batch(self, states, backend=None):
        vals = [self.ev(s, self.b) for s in states]
        for s, v in zip(states, vals):
            self.leaf_depths.append(self.depth.get(id(s), 0))
            node, r = id(s), float(v)
            while True:
                self.N[node] = self.N.get(node, 0) + 1
                if node not in self.parent:
                    break
                p, a = self.parent[node]
                self.Na[(p, a)] = self.Na.get((p, a), 0) + 1
                self.Wa[(p, a)] = self.Wa.get((p, a), 0.0) - r
                node, r = p, -r
        return vals


def run_reference(backend, root, evaluator, policy, sims, c, batch):
    sh = Shadow(backend, EVALS[evaluator], POLICIES[policy])
    move = ref_mcts.get_move(root, sh, sh, sh, sims, c, batch)
    mv = sh.moves_of[id(root)]
    rid = id(root)
    return {
        "moves": [list(m[0]) + [m[1]] if backend is ref_chess else [m[0]] for m in mv],
        "Na": [sh.Na.get((rid, i), 0) for i in range(len(mv))],
        "Wa": [sh.Wa.get((rid, i), 0.0) for i in range(len(mv))],
        "best": mv.index(move),
        "nodes_created": len(sh.parent) + 1,
        "sum_leaf_depth": sum(sh.leaf_depths),
        "max_leaf_depth": max(sh.leaf_depths) if sh.leaf_depths else 0,
    }


# ------------------------------------------------------------------ C4 fixtures
def c4_state_from_cols(cols):
    s = ref_c4.create_init_state()
    for c in cols:
        s = ref_c4.play_move(s, (c, 0))
    return s


def c4_rows(s):
    return ["".join(r) for r in s.board]


def gen_c4_search():
    rng = random.Random(20261018)
    roots = [[], [3, 3, 3, 3, 3, 3, 0, 1, 0, 1, 0, 1], [3], [0, 6, 0, 6, 0, 6], [3, 2, 3, 2, 3, 2]]
    # nearly full boards: no-move (full board) nodes get re-selected inside a batch
    filler = []
    for c in (0, 1, 2, 4, 5, 6, 3):
        filler += [c] * 6
    roots += [filler[:36], filler[:39], filler[:41]]
    for _ in range(8):
        n = rng.randrange(2, 30)
        s, cols = ref_c4.create_init_state(), []
        while len(cols) < n:
            legal = sorted(m[0] for m in ref_c4.get_legal_moves(s))
            if not legal:
                break
            col = rng.choice(legal)
            cols.append(col)
            s = ref_c4.play_move(s, (col, 0))
        roots.append(cols)
    cases = []
    for ri, cols in enumerate(roots):
        for ev in ("c4_terminal", "c4_positional"):
            variants = [(800, 1.4, 32, "first")]
            if ri < 3:
                variants += [(1, 1.4, 32, "first"), (31, 1.4, 32, "first"), (33, 1.4, 32, "first"), (100, 2.5, 32, "first"),
                             (200, 0.5, 7, "first"), (64, 1.4, 1, "first"), (800, 1.4, 32, "last"), (300, 1.25, 16, "last")]
            for sims, c, batch, pol in variants:
                root = c4_state_from_cols(cols)
                if not ref_c4.get_legal_moves(root):
                    continue
                out = run_reference(ref_c4, root, ev, pol, sims, c, batch)
                cases.append({"cols": cols, "rows": c4_rows(root), "turn": root.turn, "evaluator": ev, "policy": pol,
                              "sims": sims, "c": c, "batch": batch, **out})
    return cases


def gen_c4_rules():
    rng = random.Random(7)
    out = []
    for g in range(40):
        s = ref_c4.create_init_state()
        trace = []
        while True:
            legal = list(ref_c4.get_legal_moves(s))
            rec = {"rows": c4_rows(s), "turn": s.turn, "legal": [m[0] for m in legal], "win": ref_c4.check_win(s),
                   "draw": ref_c4.check_draw(s), "tensor": ref_c4.state_to_tensor(s).astype(int).reshape(-1).tolist()}
            trace.append(rec)
            # like random_rollout (value_functions.py:39) stop at win/draw -- but keep some games going past a win
            if not legal or ((rec["win"] or rec["draw"]) and g % 4 != 0):
                break
            m = rng.choice(sorted(legal))
            rec["played"] = m[0]
            s = ref_c4.play_move(s, m)
        out.append(trace)
    return out


# ------------------------------------------------------------------ chess fixtures
# Every FEN carries explicit clock fields: with them missing the reference reads an
# uninitialised int (chess_backend.cpp:529-531, `iss >> hm` after EOF leaves hm untouched).
FENS = {
    "start": "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1",
    "kiwipete": "r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1",
    "pos3": "8/2p5/3p4/KP5r/1R3p1k/8/4P1P1/8 w - - 0 1",
    "pos4": "r3k2r/Pppp1ppp/1b3nbN/nP6/BBP1P3/q4N2/Pp1P2PP/R2Q1RK1 w kq - 0 1",
    "pos5": "rnbq1k1r/pp1Pbppp/2p5/8/2B5/8/PPP1NnPP/RNBQK2R w KQ - 1 8",
    "mate_in_1": "6k1/5ppp/8/8/8/8/8/R3K3 w - - 0 1",
    "fools_mate": "rnb1kbnr/pppp1ppp/8/4p3/6Pq/5P2/PPPPP2P/RNBQKBNR w KQkq - 0 1",       # tests/test_cb.py:106
    "scholars_mate": "r1bqkbnr/ppp2Qpp/n2p4/4p3/2B1P3/8/PPPP1PPP/RNB1K1NR b KQkq - 0 1",  # :107
    "stalemate": "7k/5Q2/6K1/8/8/8/8/8 b - - 0 1",                                        # :108
    "k_n_vs_k": "8/8/8/8/8/8/2n5/2K4k w - - 0 1",                                         # :109
    "k_b_vs_k": "8/8/8/1k6/8/8/4K3/5B2 w - - 0 1",                                        # :110
    "kq_vs_k": "8/8/8/3k4/8/8/4K3/3Q4 w - - 0 1",
    "two_minors": "8/8/8/3k4/8/2N5/4K3/5B2 b - - 12 40",
    "promo_race": "8/P6k/8/8/8/8/p6K/8 w - - 0 1",
    "black_promo": "8/8/8/8/8/k7/p6K/1R6 b - - 3 1",
    "pins": "4k3/4r3/8/8/4B3/8/4K3/8 w - - 0 1",
    "in_check": "rnbqkbnr/ppp2ppp/8/1B1pp3/4P3/8/PPPP1PPP/RNBQK1NR b KQkq - 1 3",
}


def ch_rec(s):
    return {"board": bytes(s.board).decode(), "turn": s.turn, "fifty": s.fifty_move_rule_counter,
            "flags": [int(s.w_ck), int(s.w_cq), int(s.b_ck), int(s.b_cq)]}


def mv_list(moves):
    return [list(m[0]) + [m[1]] for m in moves]


def perft(s, d):
    mv = ref_chess.get_legal_moves(s)
    if d <= 1:
        return len(mv)
    return sum(perft(ref_chess.play_move(s, m), d - 1) for m in mv)


def gen_chess_rules():
    out = {"fens": {}, "perft": {}, "playouts": []}
    depth = {"start": 4, "kiwipete": 3, "pos3": 4, "pos4": 3, "pos5": 3}
    for name, fen in FENS.items():
        s = ref_chess.state_from_fen(fen)
        out["fens"][name] = {"fen": fen, **ch_rec(s), "legal": mv_list(ref_chess.get_legal_moves(s)),
                             "win": ref_chess.check_win(s), "draw": ref_chess.check_draw(s),
                             "tensor_planes": tensor_bits(ref_chess.state_to_tensor(s))}
        if name in depth:
            out["perft"][name] = [perft(s, d) for d in range(1, depth[name] + 1)]
    rng = random.Random(99)
    for g in range(12):
        s = ref_chess.create_init_state() if g % 3 else ref_chess.state_from_fen(rng.choice(list(FENS.values())))
        trace = []
        for ply in range(160):
            mv = ref_chess.get_legal_moves(s)
            # hist_white / hist_black are not stored: they are the moves "played" so far, per side, newest first
            rec = {**ch_rec(s), "legal": mv_list(mv), "win": ref_chess.check_win(s), "draw": ref_chess.check_draw(s)}
            if ply % 8 == 0:
                rec["tensor_planes"] = tensor_bits(ref_chess.state_to_tensor(s))
            trace.append(rec)
            if not mv:
                break
            # bias toward shuffling so that repetition / 50-ply draws show up
            m = mv[0] if (g % 4 == 3 and ply % 2 == 0) else rng.choice(mv)
            rec["played"] = list(m[0]) + [m[1]]
            s = ref_chess.play_move(s, m)
        out["playouts"].append(trace)
    # the knight shuffle of tests/test_cb.py:53-79 (repetition draw appears before the 50-ply one)
    s = ref_chess.create_init_state()
    trace = []
    wb, bb = {(7, 6, 5, 5), (5, 5, 7, 6)}, {(0, 6, 2, 5), (2, 5, 0, 6)}
    for ply in range(52):
        mv = ref_chess.get_legal_moves(s)
        m = next(x for x in mv if x[0] in (wb if s.turn == 0 else bb))
        trace.append({**ch_rec(s), "legal": mv_list(mv), "win": ref_chess.check_win(s), "draw": ref_chess.check_draw(s),
                      "played": list(m[0]) + [m[1]]})
        s = ref_chess.play_move(s, m)
    out["playouts"].append(trace)
    return out


def tensor_bits(t):
    """17 planes of 64 cells -> 17 integers (bit i = cell i), exact for a 0/1 tensor."""
    flat = t.reshape(17, 64)
    assert ((flat == 0) | (flat == 1)).all()
    return [int(sum(1 << i for i in range(64) if flat[p, i] == 1)) for p in range(17)]


def gen_chess_search():
    cases = []
    plan = [("start", 1000, 1.4, 32, "first"), ("start", 1600, 1.4, 32, "first"), ("kiwipete", 800, 1.4, 32, "first"),
            ("mate_in_1", 400, 1.4, 32, "first"), ("start", 200, 1.4, 32, "last"), ("kiwipete", 300, 2.0, 16, "last"),
            ("pos3", 800, 1.4, 32, "first"), ("pos4", 500, 1.4, 32, "first"), ("pos5", 500, 1.4, 32, "first"),
            ("kq_vs_k", 600, 1.4, 32, "first"), ("two_minors", 400, 1.4, 32, "first"), ("promo_race", 600, 1.4, 32, "first"),
            ("black_promo", 300, 1.4, 8, "first"), ("pins", 300, 1.4, 32, "first"), ("in_check", 300, 1.4, 32, "first"),
            ("start", 1, 1.4, 32, "first"), ("start", 33, 1.4, 32, "first"), ("kiwipete", 100, 1.4, 1, "first")]
    for name, sims, c, batch, pol in plan:
        root = ref_chess.state_from_fen(FENS[name])
        out = run_reference(ref_chess, root, "chess_crude", pol, sims, c, batch)
        cases.append({"name": name, "fen": FENS[name], **ch_rec(root), "evaluator": "chess_crude", "policy": pol,
                      "sims": sims, "c": c, "batch": batch, **out})
    return cases


def main():
    py_order = {str(mask): [m[0] for m in list({(i, 0) for i in range(7) if mask >> i & 1})] for mask in range(128)}
    outputs = {
        "c4_order_py312.json": {"python": sys.version.split()[0], "order": py_order},
        "c4_search.json": gen_c4_search(),
        "c4_rules.json.gz": gen_c4_rules(),
        "chess_rules.json.gz": gen_chess_rules(),
        "chess_search.json": gen_chess_search(),
    }
    for name, obj in outputs.items():
        opener = gzip.open if name.endswith(".gz") else open
        with opener(os.path.join(HERE, name), "wt") as f:
            json.dump(obj, f, separators=(",", ":"))
        print(name, os.path.getsize(os.path.join(HERE, name)), "bytes")


if __name__ == "__main__":
    main()
