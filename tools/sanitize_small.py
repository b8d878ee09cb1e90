"""Small run of every kernel family for compute-sanitizer (memcheck): tower forward, fused and split search."""
import numpy as np
import torch
from zeroclone_b200 import _ffi
from zeroclone_b200.evaluator import NetEvaluator
from zeroclone_b200.models.connect4_value.network import ValueNetwork
from zeroclone_b200.search import TreeSearch
from zeroclone_b200.workloads import c4_roots_set_b, chess_roots_set_b

torch.manual_seed(0)
ev = NetEvaluator(ValueNetwork().eval(), "cuda")
x = (torch.rand(11, 2, 6, 7) < 0.3).to("cuda", ev.dtype)
print("tower", ev(x).cpu().tolist()[:3])
ts = TreeSearch(_ffi.GAME_C4, 32, 64)
ts.set_roots(c4_roots_set_b(32))
ts.run(64, 1.4, 32, _ffi.EVAL_C4_POSITIONAL, _ffi.POLICY_FIRST)
print("c4 fused", ts.results()["result"]["best"][:4])
ts.set_roots(c4_roots_set_b(32))
ts.run_network(ev, 64, 1.4, 32, _ffi.POLICY_FIRST)
print("c4 net", ts.results()["result"]["best"][:4])
tc = TreeSearch(_ffi.GAME_CHESS, 8, 64)
tc.set_roots(chess_roots_set_b(8))
tc.run(64, 1.4, 32, _ffi.EVAL_CHESS_CRUDE, _ffi.POLICY_RANDOM, seed=3)
print("chess fused", tc.results()["result"]["best"][:4])
torch.cuda.synchronize()
print("done")
