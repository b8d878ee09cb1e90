"""Development aid: the full value-net search (stub leaves, fused tower, backprop) is deterministic run to run:
identical whole-tree hashes for BASELINE configs[1] and configs[3]."""
import numpy as np, torch
from zeroclone_b200 import _ffi
from zeroclone_b200.evaluator import NetEvaluator
from zeroclone_b200.search import TreeSearch
from zeroclone_b200.workloads import c4_roots_set_b, chess_roots_set_b
for game, n, sims in (("c4", 4096, 800), ("chess", 2048, 800)):
    if game == "c4":
        from zeroclone_b200.models.connect4_value.network import ValueNetwork
        roots, G = c4_roots_set_b(n), _ffi.GAME_C4
    else:
        from zeroclone_b200.models.chess_value.network import ValueNetwork
        roots, G = chess_roots_set_b(n), _ffi.GAME_CHESS
    torch.manual_seed(0)
    ev = NetEvaluator(ValueNetwork().eval(), "cuda")
    ts = TreeSearch(G, n, sims)
    hs = []
    for r in range(3):
        ts.set_roots(roots)
        ts.run_network(ev, sims, 1.4, 32, _ffi.POLICY_FIRST)
        hs.append(ts.tree_hash().copy())
    print(game, "identical whole-tree hashes over 3 runs:", bool((hs[0] == hs[1]).all() and (hs[1] == hs[2]).all()), flush=True)
