"""Freeze the library's own PUCT definition (zeroclone_b200/csrc/puct.cuh, oracle/zc_oracle.c:zo_search_puct) as golden vectors.
The reference has no PUCT, so these are NOT reference outputs: they pin the definition against accidental change.
Also records SHA-256 digests of the benchmark's root sets (workloads.py, set B).   python tests/golden/make_golden_puct.py"""
import hashlib
import json
import os
import sys

REPO = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, REPO)
from oracle import zc_oracle as zo  # noqa: E402
from zeroclone_b200.workloads import c4_roots_set_b, chess_roots_set_b  # noqa: E402

cases = []
for cols, sims, c, batch, ev, vl in (([], 200, 1.4, 32, "c4_positional", 1.0), ([3, 3, 2, 4], 333, 2.5, 8, "c4_terminal", 0.5),
                                     ([0, 1, 0, 1, 0, 1], 150, 0.7, 1, "c4_positional", 1.0)):
    o = zo.search_puct(zo.GAME_C4, zo.c4_from_moves(cols), sims, c, batch, getattr(zo, "EVAL_" + ev.upper()), vl, 0)
    cases.append({"game": "c4", "cols": cols, "sims": sims, "c": c, "batch": batch, "evaluator": ev, "virtual_loss": vl, "prior_weight": 0,
                  "Na": o.Na, "Wa": o.Wa, "best": o.best, "nodes_created": o.nodes_created, "max_leaf_depth": o.max_leaf_depth, "tree_hash": str(o.tree_hash)})
for fen, sims, c, batch, pw in (("rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1", 300, 1.4, 32, 0),
                                ("r1bqkbnr/pppp1ppp/2n5/4p3/2B1P3/5Q2/PPPP1PPP/RNB1K1NR w KQkq - 0 1", 400, 30.0, 32, 2)):
    o = zo.search_puct(zo.GAME_CHESS, zo.ch_from_fen(fen), sims, c, batch, zo.EVAL_CHESS_CRUDE, 1.0, pw)
    cases.append({"game": "chess", "fen": fen, "sims": sims, "c": c, "batch": batch, "evaluator": "chess_crude", "virtual_loss": 1.0, "prior_weight": pw,
                  "Na": o.Na, "Wa": o.Wa, "best": o.best, "nodes_created": o.nodes_created, "max_leaf_depth": o.max_leaf_depth, "tree_hash": str(o.tree_hash)})
roots = {"c4_set_b_4096": hashlib.sha256(c4_roots_set_b(4096).tobytes()).hexdigest(),
         "c4_set_b_first_id_30000_x64": hashlib.sha256(c4_roots_set_b(64, first_tree_id=30000).tobytes()).hexdigest(),
         "chess_set_b_2048": hashlib.sha256(chess_roots_set_b(2048).tobytes()).hexdigest()}
json.dump({"puct": cases, "root_sets_sha256": roots}, open(os.path.join(os.path.dirname(__file__), "puct_and_roots.json"), "w"), indent=1)
print("wrote", len(cases), "PUCT cases and", len(roots), "root-set digests")
