"""Development aid: where the device-resident chess self-play loop spends its time (search vs per-ply overhead)."""
import sys, time
import torch
from zeroclone_b200.games.chess import chess_backend as backend
from zeroclone_b200.policy_functions import Policy
from zeroclone_b200.selfplay import DeviceSelfPlay
from zeroclone_b200.value_functions import Value

games = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
for sims in (32, 1600):
    sp = DeviceSelfPlay(backend, Value("crude_chess_score"), Policy("random"), n_slots=games)
    sp.play(min(games, 1024), sims, 1.4, seed=1, record=False)     # warm-up
    t0 = time.time()
    out = sp.play(games, sims, 1.4, seed=0, record=False)
    dt = time.time() - t0
    print(f"sims={sims}: {out['moves']} plies in {dt:.2f} s ({out['moves'] * sims / dt:.3e} sims/s, {games / dt * 3600:.3e} games/h)", flush=True)
