"""Arena: the latest network against the first and the previous checkpoint, each side evaluated by its own
network (`Engine(cfg, value_functions=[white, black])`, `values[state.turn]`: engine.py:29-35,127 of the
reference; command line and scoring of its scripts/evaluate.py).  Every ply is one batched device search per
side over all unfinished games, so the two networks are two resident tower handles.

    python scripts/evaluate.py -c configs/chess_value.yaml -n 20
"""
from __future__ import annotations

import argparse
import os
import sys
from typing import Callable, Dict, List, Optional, Sequence

import yaml

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from zeroclone_b200.engine import Engine          # noqa: E402
from zeroclone_b200.models import core as mcore   # noqa: E402
from zeroclone_b200.value_functions import Value  # noqa: E402

WHITE_WON, DRAW, BLACK_WON = 1, 0, -1             # Engine's result convention (engine.py:148-153)


def simulate(engine, total: int, log: Callable[[str], None] = lambda msg: None) -> List[Optional[int]]:
    """`total` games, at most engine.threads in flight; when one ends its result is kept and, while fewer than
    `total` games exist, a fresh game joins the pool -- the refill rule of the reference's simulate()."""
    in_flight = min(total, engine.threads)
    outcome: List[Optional[int]] = [None] * in_flight
    running = list(range(in_flight))
    search = dict(simulations=engine.config["mcts"]["simulations"], c=engine.config["mcts"]["c_puct"])
    while running:
        step: Dict[int, Optional[int]] = engine.play_mcts_parallel(running, search["simulations"], search["c"])
        survivors = []
        for game in running:
            if step[game] is None:
                survivors.append(game)
                continue
            outcome[game] = step[game]
            log(f"game {game} finished: {step[game]:+d}")
            if len(outcome) < total:
                outcome.append(None)
                survivors.append(engine.add_game())
        running = sorted(survivors)
    return outcome


def win_rate(results: Sequence[int], latest_is_white: bool) -> float:
    """score of the latest network over these games: win 1, draw 1/2, loss 0"""
    ours = WHITE_WON if latest_is_white else BLACK_WON
    points = {ours: 1.0, DRAW: 0.5}
    return sum(points.get(r, 0.0) for r in results) / len(results)


def evaluate_pair(cfg: dict, latest_v, other_v, games: int) -> float:
    """half the games with the latest network as white, the rest as black; game-weighted mean score"""
    plan = [(games // 2, True), (games - games // 2, False)]
    total = 0.0
    for count, latest_white in plan:
        if count == 0:
            continue
        sides = [latest_v, other_v] if latest_white else [other_v, latest_v]
        total += count * win_rate(simulate(Engine(cfg, value_functions=sides), count), latest_white)
    return total / games


def main() -> None:
    cli = argparse.ArgumentParser()
    cli.add_argument("-c", "--config", required=True, help="YAML config")
    cli.add_argument("-n", "--games", type=int, default=10, help="games per match-up")
    opts = cli.parse_args()
    with open(opts.config, "r", encoding="utf-8") as fh:
        cfg = yaml.safe_load(fh)
    kind = cfg["value"]["model_type"]
    shared = dict(model_type=kind, batch_size=cfg["value"].get("batch_size", 1))
    history = mcore.list_checkpoints(kind)
    if not history:
        raise RuntimeError(f"No checkpoints found for {kind}")
    latest = Value("network_latest", **shared)
    opponents = {"first": Value("network_at_path", path=str(history[0]), **shared),
                 "prev ": Value("network_at_path", path=str(history[-1]), **shared)}
    print(f"Evaluating {kind}  -  {opts.games} games each match-up\n")
    scores = {name: evaluate_pair(cfg, latest, other, opts.games) for name, other in opponents.items()}
    print("Win-rates for *latest* network")
    print("--------------------------------")
    for name, score in scores.items():
        print(f"vs {name} checkpoint : {score:.2%}")


if __name__ == "__main__":
    main()
