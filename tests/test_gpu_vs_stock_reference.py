"""The CUDA path against the UNMODIFIED reference, live: the stock `engine.mcts.get_move` (mcts.cpp:102-160 behind
bindings_mcts.cpp:9-11) and the stock backends, as packed by `make -C oracle ref` into oracle/_ref/pyref.zip (byte-identical
files, tests/test_reference_arm.py), search roots of the benchmark's own root sets in a subprocess with CUDA hidden; the
product searches the same roots on the GPU through the C-ABI.  `get_move` returns the chosen move only (mcts.cpp:150-159),
so that is what is compared here -- per-child statistics and whole-tree hashes are compared against the oracle
(test_gpu_parity_bench_sets.py), which is itself pinned to this same stock module (tests/test_oracle_vs_reference*.py).

Deterministic evaluators and the first-untried policy only: `Policy.random` draws from Python's global `random`.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import ref_harness as rh
from zeroclone_b200 import _ffi
from zeroclone_b200.search import TreeSearch
from zeroclone_b200.workloads import c4_roots_set_b, chess_roots_set_b

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not rh.ref_available(), reason="oracle/_ref/pyref.zip not built")]

STOCK = r"""
import json, sys
sys.path.insert(0, {repo!r})
from oracle import ref_harness as rh
s = rh.stock(need_torch=True)
assert 'zeroclone_b200' not in sys.modules
game, evaluator, sims, starts, per = {game!r}, {evaluator!r}, {sims}, {starts!r}, {per}
value = rh.make_value(game, evaluator)
out = []
for st in starts:
    if game == 'chess':
        backend, states = s.chess, [rh.chess_state_from_bytes(r) for r in rh.chess_roots_set_b(per, first_tree_id=st)]
    else:
        backend, states = s.c4, [rh.c4_state_from_bits(*r) for r in rh.c4_roots_set_b(per, first_tree_id=st)]
    for state in states:
        mv = s.mcts.get_move(state, value, rh.first_policy, backend, sims, 1.4, 32)
        out.append([int(x) for x in mv[0]] + [float(mv[1])] if game == 'chess' else [int(mv[0]), int(mv[1])])
print('MOVES ' + json.dumps(out))
"""


def stock_moves(game, evaluator, sims, starts, per):
    code = STOCK.format(repo=REPO, game=game, evaluator=evaluator, sims=sims, starts=starts, per=per)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900,
                         env=dict(os.environ, CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="1"))
    assert out.returncode == 0, out.stderr[-3000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("MOVES ")][-1]
    return json.loads(line[6:])


def test_c4_chosen_moves_equal_the_stock_reference():
    starts, per, sims = [0, 10900, 21800, 32704], 64, 800          # c4_heuristic: 32768 trees x 800 sims, c4_positional
    want = stock_moves("connect4", "c4_positional", sims, starts, per)
    roots = np.concatenate([c4_roots_set_b(per, first_tree_id=s) for s in starts])
    ts = TreeSearch(_ffi.GAME_C4, len(roots), sims)
    ts.set_roots(roots)
    ts.run(sims, 1.4, 32, _ffi.EVAL_C4_POSITIONAL, _ffi.POLICY_FIRST)
    got = [[int(r["best_move"][0]), 0] for r in ts.results()["result"]]
    assert got == want


def test_chess_chosen_moves_equal_the_stock_reference():
    starts, per, sims = [0, 8000, 16352], 32, 1600                 # chess_crude: 16384 trees x 1600 sims
    want = stock_moves("chess", "chess_crude", sims, starts, per)
    roots = np.concatenate([chess_roots_set_b(per, first_tree_id=s) for s in starts])
    ts = TreeSearch(_ffi.GAME_CHESS, len(roots), sims)
    ts.set_roots(roots)
    ts.run(sims, 1.4, 32, _ffi.EVAL_CHESS_CRUDE, _ffi.POLICY_FIRST)
    res = ts.results()["result"]
    got = [[int(x) for x in r["best_move"]] + [float(r["best_move_value"])] for r in res]
    assert got == want
