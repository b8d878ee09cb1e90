"""Drop-in check at the Engine API: the UNMODIFIED reference `Engine` (engine/engine.py:14-157 over its own mcts module and
backends, from oracle/_ref/pyref.zip, in a subprocess with CUDA hidden) and this repo's `Engine` on the GPU play the same games
through the same calls -- add_game, play_move, play_mcts_parallel, get_state, get_dataset -- and must produce the same states
ply by ply, the same results and the same dataset.  Deterministic ingredients only: crude_chess_score and the first-untried
expansion order (the stock engine gets `engine.policy = first`, an attribute the reference exposes, engine.py:27).
"""
import hashlib
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import ref_harness as rh

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not rh.ref_available(), reason="oracle/_ref/pyref.zip not built")]

FENS = ["6k1/5ppp/8/8/8/8/5PPP/3R2K1 w - - 0 1",              # back-rank mate in one
        "7k/5Q2/5K2/8/8/8/8/8 w - - 0 1",                      # queen + king: mate in one
        "r1bqkb1r/pppp1ppp/2n2n2/4p2Q/2B1P3/8/PPPP1PPP/RNB1K1NR w KQkq - 0 1",   # scholar's mate in one
        "8/8/8/8/8/5k2/6q1/7K w - - 0 1"]                      # white is already mated
SIMS, PLIES = 300, 6

SCRIPT = r"""
import hashlib, json
def run(Engine, backend, first_policy):
    eng = Engine({'game': 'chess', 'backend': 'chess_backend', 'value_function': 'crude_chess_score',
                  'policy_functions': FIRST_NAME, 'threads': 2, 'mcts': {'simulations': SIMS, 'c_puct': 1.4}})
    if first_policy is not None:
        eng.policy = first_policy
    for fen in FENS:
        eng.add_game(backend.state_from_fen(fen))
    # the two default games leave the opening by hand, through play_move's legality check
    mv = eng.legal_moves(0)
    eng.play_move(mv[7], 0)
    mv = eng.legal_moves(1)
    eng.play_move(mv[12], 1)
    def plain(x):
        return [plain(e) for e in x] if isinstance(x, (tuple, list)) else float(x)
    def snap(s):
        return [bytes(s.board).hex(), int(s.turn), int(s.fifty_move_rule_counter), bool(s.w_ck), bool(s.w_cq), bool(s.b_ck), bool(s.b_cq),
                plain(list(s.hist_white)), plain(list(s.hist_black))]
    log = []
    idxs = list(range(2 + len(FENS)))
    for ply in range(PLIES):
        res = eng.play_mcts_parallel(idxs, simulations=SIMS, c=1.4)
        log.append({'results': [res[i] for i in idxs], 'states': [snap(eng.get_state(i)) for i in idxs]})
    x, y = eng.get_dataset()
    return {'log': log, 'dataset_shape': list(x.shape), 'labels': [float(v) for v in y],
            'dataset_sha': hashlib.sha256(x.tobytes()).hexdigest(), 'hist_results': [h.result for h in eng.history],
            'hist_len': [len(h.states) for h in eng.history]}
"""

STOCK_TAIL = r"""
import sys
sys.path.insert(0, {repo!r})
from oracle import ref_harness as rh
s = rh.stock(need_torch=True)
import importlib
eng_mod = importlib.import_module('engine.engine')
assert eng_mod.__file__.startswith(rh.pyref_dir()) and 'zeroclone_b200' not in sys.modules
FENS, SIMS, PLIES, FIRST_NAME = {fens!r}, {sims}, {plies}, 'random'
print('OUT ' + json.dumps(run(eng_mod.Engine, s.chess, rh.first_policy)))
"""


def test_same_games_through_the_engine_api():
    code = SCRIPT + STOCK_TAIL.format(repo=REPO, fens=FENS, sims=SIMS, plies=PLIES)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900,
                         env=dict(os.environ, CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="1"))
    assert out.returncode == 0, out.stderr[-3000:]
    want = json.loads([l for l in out.stdout.splitlines() if l.startswith("OUT ")][-1][4:])

    from zeroclone_b200.engine import Engine
    from zeroclone_b200.games.chess import chess_backend
    env = {"FENS": FENS, "SIMS": SIMS, "PLIES": PLIES, "FIRST_NAME": "first", "hashlib": hashlib, "json": json}
    exec(SCRIPT, env)
    got = env["run"](Engine, chess_backend, None)

    assert got["hist_results"] == want["hist_results"] and got["hist_len"] == want["hist_len"]
    for ply, (g, w) in enumerate(zip(got["log"], want["log"])):
        assert g["results"] == w["results"], ply
        for i, (gs, ws) in enumerate(zip(g["states"], w["states"])):
            assert gs == ws, (ply, i)
    assert got["dataset_shape"] == want["dataset_shape"] and got["labels"] == want["labels"]
    assert got["dataset_sha"] == want["dataset_sha"]
    assert any(r is not None for r in want["hist_results"])          # the mates were found and labelled
    assert np.isfinite(np.array(want["labels"])).all()
