"""`Value.random_rollout` (engine/value_functions.py:35-45) on the device against the UNMODIFIED reference function, as a
distribution: the in-kernel rollout of a leaf (uniform random legal moves to the end, -1 / 0 / +1 from the leaf's side to
move) must produce the same win / draw / loss frequencies as the stock Python function sampled on the same leaf.  The
device side evaluates ONE leaf per tree (1 simulation, first-untried order) on 20 000 trees with different random streams;
the stock side (oracle/_ref/pyref.zip, CUDA hidden, `random.seed` fixed) samples its function 10 000 times.  Chi-square
homogeneity test, 2 degrees of freedom, p = 1e-4 per position (critical value 18.42); both sides are seeded, so the test
is deterministic.  (A first version with 4 000 stock samples and p = 0.001 tripped on the initial position at chi2 = 13.87; 40 000 stock
samples then gave 57.40 / 0.24 / 42.36 % against the device's 57.24 / 0.28 / 42.49 %: the small sample was the outlier.)
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import ref_harness as rh
from zeroclone_b200 import _ffi
from zeroclone_b200.games.connect4 import c4_backend as c4
from zeroclone_b200.search import TreeSearch
from zeroclone_b200.workloads import c4_roots_set_b

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not rh.ref_available(), reason="oracle/_ref/pyref.zip not built")]
N_DEVICE, N_STOCK, CRIT = 20000, 10000, 18.42

STOCK = r"""
import json, random, sys
sys.path.insert(0, {repo!r})
from oracle import ref_harness as rh
s = rh.stock(need_torch=True)
random.seed(999)
v = s.Value("random_rollout")
out = []
for x, o, turn, col in {cases!r}:
    leaf = s.c4.play_move(rh.c4_state_from_bits(x, o, turn), (col, 0))
    c = {{-1: 0, 0: 0, 1: 0}}
    for _ in range({n}):
        c[v(leaf, backend=s.c4)] += 1
    out.append([c[-1], c[0], c[1]])
print("COUNTS " + json.dumps(out))
"""


def test_rollout_outcome_frequencies_match_the_stock_function():
    roots = [c4_roots_set_b(1, first_tree_id=t)[0] for t in (0, 5, 9, 12)]      # 0, 5, 9 and 12 random plies from the start
    cases, device = [], []
    for rec in roots:
        batch = np.repeat(np.array([rec]), N_DEVICE)
        ts = TreeSearch(_ffi.GAME_C4, N_DEVICE, 32)
        ts.set_roots(batch)
        ts.run(1, 1.4, 1, _ffi.EVAL_C4_ROLLOUT, _ffi.POLICY_FIRST, seed=7)
        out = ts.results()
        assert (out["visits"].sum(axis=1) == 1).all()
        first = out["visits"].argmax(axis=1)
        assert (first == first[0]).all()                                        # every tree expanded the same (first) move
        state = c4.unpack_state(int(rec["x"]), int(rec["o"]), int(rec["turn"]))
        col = list(c4.get_legal_moves(state))[int(first[0])][0]
        v = -out["value_sums"][np.arange(N_DEVICE), first]                      # backprop stores -value at the parent's edge (mcts.cpp:91)
        assert set(np.unique(v)).issubset({-1.0, 0.0, 1.0})
        device.append([int((v == -1).sum()), int((v == 0).sum()), int((v == 1).sum())])
        cases.append((int(rec["x"]), int(rec["o"]), int(rec["turn"]), int(col)))
    code = STOCK.format(repo=REPO, cases=cases, n=N_STOCK)
    run = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900,
                         env=dict(os.environ, CUDA_VISIBLE_DEVICES="", OMP_NUM_THREADS="1"))
    assert run.returncode == 0, run.stderr[-3000:]
    stock = json.loads([l for l in run.stdout.splitlines() if l.startswith("COUNTS ")][-1][7:])
    for case, d, s in zip(cases, device, stock):
        d, s = np.array(d, float), np.array(s, float)
        pooled = (d + s) / (d.sum() + s.sum())
        keep = pooled > 0
        chi2 = float((((d - d.sum() * pooled) ** 2 / (d.sum() * pooled))[keep]).sum() + (((s - s.sum() * pooled) ** 2 / (s.sum() * pooled))[keep]).sum())
        print(f"root {case}: device loss/draw/win {d / d.sum()}, stock {s / s.sum()}, chi2 {chi2:.2f}")
        assert chi2 < CRIT, (case, d.tolist(), s.tolist(), chi2)
