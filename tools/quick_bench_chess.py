"""Ad-hoc timing of the fused chess search (development aid)."""
import sys
import numpy as np
import torch
from zeroclone_b200 import _ffi
from zeroclone_b200.search import TreeSearch
import ctypes as C

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
sims = int(sys.argv[2]) if len(sys.argv) > 2 else 1600
policy = {"first": _ffi.POLICY_FIRST, "random": _ffi.POLICY_RANDOM, "immediate_value": _ffi.POLICY_IMMEDIATE_VALUE}[sys.argv[3] if len(sys.argv) > 3 else "first"]
L = _ffi.lib()
s = _ffi.ChessState()
L.zc_chess_init_state(C.byref(s))
roots = np.zeros(n, dtype=_ffi.CHESS_STATE_DTYPE)
rng = np.random.default_rng(0)
for i in range(n):
    st = _ffi.ChessState.from_buffer_copy(s)
    for ply in range(i % 13):
        mv = (_ffi.ChessMove * 256)()
        k = L.zc_chess_legal_moves(C.byref(st), mv)
        if k == 0:
            break
        o = _ffi.ChessState()
        L.zc_chess_play_move(C.byref(st), C.byref(mv[int(rng.integers(k))]), C.byref(o))
        st = o
    roots[i] = np.frombuffer(bytes(st), dtype=_ffi.CHESS_STATE_DTYPE)[0]
ts = TreeSearch(_ffi.GAME_CHESS, n, sims)
print("device MB", ts.device_bytes / 1e6)
for r in range(3):
    ts.set_roots(roots)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    ts.run(sims, 1.4, 32, _ffi.EVAL_CHESS_CRUDE, policy)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    print(f"chess n={n} sims={sims} {ms:.2f} ms {n * sims / ms * 1e3:.3e} sims/s")
res = ts.results(stats=False)
c = ts.counters()
print(c, "mean depth", c["sum_leaf_depth"] / c["simulations"], "slots/node", c["arena_slots_used"] / c["nodes"])
