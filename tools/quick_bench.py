"""Ad-hoc timing of the fused C4 search (development aid; bench.py is the contract)."""
import sys
import time

import numpy as np
import torch

from zeroclone_b200 import _ffi
from zeroclone_b200.search import TreeSearch, c4_pack_cols

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
sims = int(sys.argv[2]) if len(sys.argv) > 2 else 800
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
rng = np.random.default_rng(0)
roots = np.zeros(n, dtype=_ffi.C4_STATE_DTYPE)
for i in range(n):
    cols = []
    x = o = 0
    x, o, t = c4_pack_cols([int(c) for c in rng.integers(0, 7, size=i % 13)])
    roots[i] = (x, o, t, 0)
ts = TreeSearch(_ffi.GAME_C4, n, sims)
print("device bytes", ts.device_bytes / 1e6, "MB")
for ev in (_ffi.EVAL_C4_TERMINAL, _ffi.EVAL_C4_POSITIONAL):
    for r in range(reps):
        ts.set_roots(roots)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        ts.run(sims, 1.4, 32, ev, _ffi.POLICY_FIRST)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b)
        print(f"eval={ev} n={n} sims={sims} {ms:.3f} ms  {n * sims / ms * 1e3:.3e} sims/s")
    c = ts.counters()
    print(c, "mean depth", c["sum_leaf_depth"] / c["simulations"])
