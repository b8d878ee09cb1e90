"""Arena: the latest network against the first and the previous checkpoint, each side evaluated by its
own network (`Engine(cfg, value_functions=[white, black])`, `values[state.turn]`, engine.py:29-35,127 of the
reference; CLI and win-rate convention of scripts/evaluate.py:14-91).  Every ply is one batched device
search per side over all unfinished games."""
from __future__ import annotations

import argparse
import os
import sys
from typing import Callable, List, Optional, Sequence

import yaml

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from zeroclone_b200.engine import Engine          # noqa: E402
from zeroclone_b200.models import core as mcore   # noqa: E402
from zeroclone_b200.value_functions import Value  # noqa: E402


def simulate(engine: Engine, total: int, log=lambda msg: None) -> List[Optional[int]]:
    """Play `total` games, at most engine.threads at a time; a finished game's slot is refilled with
    add_game() until `total` games exist (evaluate.py:14-31)."""
    first = min(total, engine.threads)
    results: List[Optional[int]] = [None] * first
    unfinished = set(range(first))
    sims, c = engine.config["mcts"]["simulations"], engine.config["mcts"]["c_puct"]
    while unfinished:
        partial = engine.play_mcts_parallel(sorted(unfinished), sims, c)
        for idx in sorted(unfinished):
            if partial[idx] is None:
                continue
            log(f"game {idx} finished: {partial[idx]:+d}")
            results[idx] = partial[idx]
            unfinished.discard(idx)
            if len(results) < total:
                results.append(None)
                unfinished.add(engine.add_game())
    return results


def win_rate(results: Sequence[int], latest_is_white: bool) -> float:
    """(wins + draws/2) / games from the latest network's side (+1 = white / first player won)."""
    mine = 1 if latest_is_white else -1
    return sum(1.0 if r == mine else 0.5 if r == 0 else 0.0 for r in results) / len(results)


def evaluate_pair(cfg: dict, latest_v: Callable, other_v: Callable, games: int) -> float:
    as_white = games // 2
    as_black = games - as_white
    score = 0.0
    if as_white:
        score += as_white * win_rate(simulate(Engine(cfg, value_functions=[latest_v, other_v]), as_white), True)
    if as_black:
        score += as_black * win_rate(simulate(Engine(cfg, value_functions=[other_v, latest_v]), as_black), False)
    return score / games


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("-c", "--config", required=True, help="YAML config")
    ap.add_argument("-n", "--games", type=int, default=10, help="games per match-up")
    args = ap.parse_args()
    with open(args.config, "r", encoding="utf-8") as fh:
        cfg = yaml.safe_load(fh)
    model_type = cfg["value"]["model_type"]
    batch = cfg["value"].get("batch_size", 1)
    ckpts = mcore.list_checkpoints(model_type)
    if not ckpts:
        raise RuntimeError(f"No checkpoints found for {model_type}")
    latest = Value("network_latest", model_type=model_type, batch_size=batch)
    first = Value("network_at_path", model_type=model_type, path=str(ckpts[0]), batch_size=batch)
    prev = Value("network_at_path", model_type=model_type, path=str(ckpts[-1]), batch_size=batch)
    print(f"Evaluating {model_type}  -  {args.games} games each match-up\n")
    wr_first = evaluate_pair(cfg, latest, first, args.games)
    wr_prev = evaluate_pair(cfg, latest, prev, args.games)
    print("Win-rates for *latest* network")
    print("--------------------------------")
    print(f"vs first checkpoint : {wr_first:.2%}")
    print(f"vs prev  checkpoint : {wr_prev:.2%}")


if __name__ == "__main__":
    main()
