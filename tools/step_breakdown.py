"""Development aid: wall-clock breakdown of one heuristic search step (set_roots_dev / run / results)."""
import sys, time
import numpy as np
import torch
from zeroclone_b200 import _ffi
from zeroclone_b200.search import TreeSearch
from zeroclone_b200.workloads import c4_roots_set_b, chess_roots_set_b

game = sys.argv[1] if len(sys.argv) > 1 else "c4"
trees = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
sims = int(sys.argv[3]) if len(sys.argv) > 3 else 800
chess = game == "chess"
roots = (chess_roots_set_b if chess else c4_roots_set_b)(trees)
rd = torch.from_numpy(roots.view(np.uint8).reshape(trees, -1).copy()).cuda()
ts = TreeSearch(_ffi.GAME_CHESS if chess else _ffi.GAME_C4, trees, sims)
ev = _ffi.EVAL_CHESS_CRUDE if chess else _ffi.EVAL_C4_POSITIONAL
st = torch.cuda.current_stream().cuda_stream
for it in range(6):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ts.set_roots_dev(rd.data_ptr(), trees, st)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    ts.run(sims, 1.4, 32, ev, _ffi.POLICY_FIRST, 0, st)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    r = ts.results(stats=False, stream=st)
    t3 = time.perf_counter()
    print(f"set_roots {1e3*(t1-t0):.3f} ms  run {1e3*(t2-t1):.3f} ms  results {1e3*(t3-t2):.3f} ms", flush=True)
