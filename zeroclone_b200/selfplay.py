"""Self-play driver: the reference's `simulate_games` loop (scripts/train.py:151-170) over the
batched engine -- every iteration is ONE device search over all unfinished games, one move per
game; finished games are replaced by fresh ones until `total_games` have been started.
Reports games/hour, the second half of BASELINE.json's metric."""
from __future__ import annotations

import time
from typing import Callable, List, Optional


def simulate_games(engine, total_games: int, simulations: Optional[int] = None, c: Optional[float] = None,
                   on_game_done: Optional[Callable[[int, int], None]] = None) -> dict:
    """Play `total_games` games with at most `engine.threads` in flight.
    Returns {"results": [...], "moves": plies played, "seconds": wall, "games_per_hour", "sims_per_sec"}."""
    sims = simulations if simulations is not None else engine.config["mcts"]["simulations"]
    c_val = c if c is not None else engine.config["mcts"]["c_puct"]
    first = min(total_games, engine.threads)
    unfinished = set(range(first))
    final: List[Optional[int]] = [None] * first
    plies = searched = 0
    t0 = time.time()
    while unfinished:
        batch = sorted(unfinished)
        results = engine.play_mcts_parallel(batch, simulations=sims, c=c_val)
        plies += len(batch)
        searched += sum(1 for i in batch if i in engine.last_search)
        for idx in batch:
            if results[idx] is None:
                continue
            unfinished.discard(idx)
            final[idx] = results[idx]
            if on_game_done is not None:
                on_game_done(sum(r is not None for r in final), total_games)
            if len(final) < total_games:            # refill: a new game takes a new index (train.py:166-168)
                final.append(None)
                unfinished.add(engine.add_game())
    dt = time.time() - t0
    return {"results": final, "moves": plies, "seconds": dt, "games_per_hour": len(final) / dt * 3600.0,
            "sims_per_sec": plies * sims / dt}
