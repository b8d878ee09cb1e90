import sys, numpy as np, torch
from zeroclone_b200 import _ffi
from zeroclone_b200.search import TreeSearch
from zeroclone_b200.workloads import chess_roots_set_b
n, sims = 16384, 1600
roots = chess_roots_set_b(n)
for per_node in (72, 24, 12):
    ts = TreeSearch(_ffi.GAME_CHESS, n, sims, arena_slots_per_tree=(sims + 1) * per_node)
    for r in range(3):
        ts.set_roots(roots)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        ts.run(sims, 1.4, 32, _ffi.EVAL_CHESS_CRUDE, _ffi.POLICY_FIRST)
        b.record(); torch.cuda.synchronize()
    c = ts.counters()
    print(per_node, f"{a.elapsed_time(b):.2f} ms", "slots/node", c["arena_slots_used"] / c["nodes"], "MB", ts.device_bytes / 1e6, flush=True)
    res = ts.results(stats=False)
    ts.close()
