"""Time `engine.mcts.get_move` end to end -- the reference's scripts/timing_check.py:15-58 for this engine: same flags
(-c/--config, --sims, --batch, --loops), same five printed lines (one warm-up call, then `loops` timed calls of get_move on the
initial position with the configuration's network evaluator and policy), so whatever parsed the reference's output keeps working.

One tree per call is the reference's shape, not this engine's: a single search occupies one warp of one SM and pays the launch
latency of every batch.  `--trees N` (additive) also times N roots in one call of the batched search the Engine uses
(`play_mcts_parallel`), which is where the GPU is.
"""
from __future__ import annotations

import argparse
import importlib
import time
from pathlib import Path

import yaml

from engine.mcts import get_move
from engine.policy_functions import Policy
from engine.value_functions import Value


def timed(fn) -> float:
    t0 = time.perf_counter()
    fn()
    return time.perf_counter() - t0


def main() -> None:
    ap = argparse.ArgumentParser(description="Time get_move() end-to-end")
    ap.add_argument("-c", "--config", required=True, help="YAML config file")
    ap.add_argument("--sims", type=int, default=2048, help="MCTS playouts")
    ap.add_argument("--batch", type=int, default=32, help="Leaf batch size")
    ap.add_argument("--loops", type=int, default=10, help="Number of moves to time")
    ap.add_argument("--trees", type=int, default=0, help="also time this many roots in ONE batched search (not in the reference)")
    args = ap.parse_args()

    with open(Path(args.config).expanduser(), "r", encoding="utf-8") as fh:
        cfg = yaml.safe_load(fh)
    backend = importlib.import_module(f"engine.games.{cfg['game']}.{cfg['backend']}")
    v_cfg = cfg.get("value")
    if v_cfg and "model_type" in v_cfg:          # the reference requires a network configuration (timing_check.py:37-41)
        value_fn = Value("network_latest", model_type=v_cfg["model_type"], batch_size=v_cfg.get("batch_size", 1))
    else:                                        # additive: heuristic configurations can be timed too
        value_fn = Value(cfg.get("value_function"), **(v_cfg or {}))
    policy = Policy(name=cfg.get("policy_function", "random"), **cfg.get("policy", {}))     # the reference reads this key here
    state = backend.create_init_state()

    def one():
        get_move(state, value_fn, policy, backend, simulations=args.sims, c=1.4, batch_size=args.batch)

    one()                                        # warm-up (timing_check.py:49)
    times = sorted(timed(one) for _ in range(args.loops))
    print(f"--- get_move timing ({args.loops} runs) ---")
    print(f"simulations : {args.sims}")
    print(f"batch size  : {args.batch}")
    print(f"mean  time  : {sum(times) / len(times):.3f} s")
    print(f"median time : {times[len(times) // 2]:.3f} s")

    if args.trees > 0:
        from zeroclone_b200.mcts import search_batch
        states = [state] * args.trees

        def many():
            search_batch(states, value_fn, policy, backend, args.sims, 1.4, args.batch)

        many()
        t = sorted(timed(many) for _ in range(max(1, args.loops // 2)))[0]
        print(f"batched     : {args.trees} trees in {t:.3f} s = {args.trees * args.sims / t:.3e} simulations/s")


if __name__ == "__main__":
    main()
