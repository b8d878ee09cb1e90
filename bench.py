#!/usr/bin/env python
"""bench.py -- MCTS simulations/sec on B200 (BASELINE.json metric), one JSON line on rank 0.

  python bench.py --gpus N --steps K --warmup W [--workload NAME] [--impl reference]

A step = one pass of the hot path over one batch of synthetic roots: `trees` independent searches
of `sims` simulations each (Engine.play_mcts_parallel's search phase: select, expand, evaluate,
backprop, root readout).  Default workload = BASELINE.json configs[1]: Connect Four, 4096
concurrent trees x 800 sims with the value network (random-init weights, bf16), roots = set B.
Per-GPU work is fixed as N grows (weak scaling): every rank owns its own 4096 trees; the search
needs no collective.

Keys beyond the base contract: roofline (dominant kernel), roofline_tree (our CUDA kernels),
cpu_baseline (the CPU port timed on this box's host cores, N=1 only), e2e (host buffers in,
host results out), gpu_launches, clocks.
`--impl reference` times the reference's own CPU search (oracle/_ref/mcts*.so, compiled from
the reference sources) on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

# stdout carries exactly one JSON line: NCCL's version banner / debug output goes to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

WORKLOADS = {
    # name: (game, evaluator, trees per GPU, sims)
    "c4_value_net": dict(game="connect4", evaluator="value_net", trees=4096, sims=800,
                         desc="BASELINE configs[1]: Connect Four batched self-play search, 4096 concurrent trees x 800 sims, "
                              "value tower 128ch x 8 blocks on 2x6x7 planes"),
    "c4_heuristic": dict(game="connect4", evaluator="c4_positional", trees=32768, sims=800,
                         desc="Connect Four, deterministic evaluator (parity configuration), 32768 trees x 800 sims"),
    "chess_crude": dict(game="chess", evaluator="chess_crude", trees=16384, sims=1600,
                        desc="BASELINE configs[4]: chess configs/crude_chess.yaml heuristic evaluator, 16384 trees x 1600 sims per GPU"),
    "chess_value_net": dict(game="chess", evaluator="value_net", trees=2048, sims=800,
                            desc="BASELINE configs[3]: chess configs/chess_value.yaml, movegen kernel + value-net leaf batching, "
                                 "2048 trees x 800 sims"),
}
C_UCT, BATCH = 1.4, 32


def read_peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops_sustained"], "measured"
    return 6650.0, 1400.0, "fallback"   # B200_PROFILING.md fallback (1.59 PF burst, ~1.4 sustained)


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except (ValueError, IndexError):
                continue
            for nme, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        if not sm:   # timed region shorter than the sampling period: take one reading now
            try:
                r = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                   capture_output=True, text=True, timeout=10).stdout.strip().split(",")
                sm, mx = [float(r[0])], float(r[1])
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------
# dram__bytes_read.sum + dram__bytes_write.sum per evaluated leaf, from the ncu captures under profiles/
TOWER_DRAM_BYTES_PER_LEAF = {"connect4": 204.9, "chess": 2325.1}


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from zeroclone_b200 import _ffi
    from zeroclone_b200.evaluator import NetEvaluator, tower_flops_per_leaf
    from zeroclone_b200.search import TreeSearch
    from zeroclone_b200.workloads import c4_roots_set_b, chess_roots_set_b

    wl = WORKLOADS[args.workload]
    chess = wl["game"] == "chess"
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    trees = args.trees or wl["trees"]
    sims = args.sims or wl["sims"]
    hbm_peak, tensor_peak, peak_src = read_peaks()

    roots = (chess_roots_set_b if chess else c4_roots_set_b)(trees, first_tree_id=rank * trees)
    roots_pinned = torch.from_numpy(roots.view(np.uint8).reshape(trees, -1).copy()).pin_memory()
    roots_dev = roots_pinned.to(dev)
    ts = TreeSearch(_ffi.GAME_CHESS if chess else _ffi.GAME_C4, trees, sims, device=local)
    use_net = wl["evaluator"] == "value_net"
    ev = None
    if use_net:
        if chess:
            from zeroclone_b200.models.chess_value.network import ValueNetwork
        else:
            from zeroclone_b200.models.connect4_value.network import ValueNetwork
        torch.manual_seed(0)
        ev = NetEvaluator(ValueNetwork().eval(), dev)      # fused sm_100a tower kernel (csrc/tower.cuh)
        flops_leaf = tower_flops_per_leaf(17, 8, 8) if chess else tower_flops_per_leaf(2, 6, 7)
    heur = {"c4_positional": _ffi.EVAL_C4_POSITIONAL, "c4_terminal": _ffi.EVAL_C4_TERMINAL,
            "chess_crude": _ffi.EVAL_CHESS_CRUDE}.get(wl["evaluator"])
    stream = torch.cuda.current_stream().cuda_stream
    phase_events = []   # (kind, start, end) CUDA events on the launching stream

    def timed(kind, fn):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        phase_events.append((kind, a, b))

    def search_resident(record):
        """inputs already in HBM; ends with the root readout (chosen move per tree) on the host"""
        ts.set_roots_dev(roots_dev.data_ptr(), trees, stream)
        if use_net:
            if record:
                run_network_timed()
            else:
                ts.run_network(ev, sims, C_UCT, BATCH, _ffi.POLICY_FIRST)
        else:
            if record:
                timed("tree", lambda: ts.run(sims, C_UCT, BATCH, heur, _ffi.POLICY_FIRST, 0, stream))
            else:
                ts.run(sims, C_UCT, BATCH, heur, _ffi.POLICY_FIRST, 0, stream)
        return ts.results(stats=False, stream=stream, reuse=True)

    def run_network_timed():
        planes, values = ts._planes, ts._values
        ts.begin(sims, C_UCT, BATCH, _ffi.POLICY_FIRST, 0)
        while ts.pending() > 0:
            timed("tree", lambda: ts.select(planes.data_ptr(), _ffi.PLANE_BF16, stream))
            timed("net", lambda: ev(planes, out=values))
            timed("tree", lambda: ts.backprop(values.data_ptr(), stream))

    def search_e2e():
        """the public call with HOST buffers: roots H2D, search, per-tree results D2H"""
        ts.set_roots(roots, stream)
        if use_net:
            ts.run_network(ev, sims, C_UCT, BATCH, _ffi.POLICY_FIRST)
        else:
            ts.run(sims, C_UCT, BATCH, heur, _ffi.POLICY_FIRST, 0, stream)
        return ts.results(stats=False, stream=stream, reuse=True)    # what get_move returns: the chosen move per tree (mcts.cpp:157-159)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        search_resident(False)
    launches0 = ts.counters()["kernel_launches"] + (ev.launches if ev is not None else 0)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        res = search_resident(True)
    t1.record()
    barrier()
    clocks = sampler.stop()
    ms = t0.elapsed_time(t1)
    launches = ts.counters()["kernel_launches"] + (ev.launches if ev is not None else 0) - launches0
    cnt = ts.counters()
    assert int(res["result"]["root_visits"].min()) == sims, "a tree did not finish its simulations"

    # e2e: same metric through host buffers
    for _ in range(2):
        search_e2e()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        out = search_e2e()
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)

    tmax = torch.tensor([ms, e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms, e2e_ms = float(tmax[0]), float(tmax[1])
    total_sims = world * trees * sims * args.steps
    value = total_sims / (ms * 1e-3)
    e2e_value = total_sims / (e2e_ms * 1e-3)

    tree_ms = sum(a.elapsed_time(b) for k, a, b in phase_events if k == "tree")
    net_ms = sum(a.elapsed_time(b) for k, a, b in phase_events if k == "net")
    n_tree = sum(1 for k, _, _ in phase_events if k == "tree")
    n_net = sum(1 for k, _, _ in phase_events if k == "net")
    # algorithmic bytes of the tree kernels: per simulation one new node is written
    # (header + state + k' zeroed edges = 16*(2+k') B) plus its link (4 B) and first statistics (16 B edge + 4 B N);
    # per batch one descent reads every node on the path (16*(2+k) B per level) and the backprop rewrites one
    # edge (16 B) and one N (4 B) per level.
    sims_done = trees * sims
    depth = cnt["sum_leaf_depth"] / max(1, cnt["simulations"])
    node_bytes = 16.0 * cnt["arena_slots_used"] / max(1, cnt["nodes"])       # measured mean node size (header+state+edges+moves)
    batches = sims_done / BATCH
    tree_bytes_step = sims_done * (node_bytes + 4 + 20) + batches * (depth + 1) * (node_bytes + 2 * 20)
    # measured DRAM traffic per simulation of the tree kernels: dram__bytes_read+write of one `ncu --set full`
    # capture of k_search_fused divided by the simulations of that launch (profiles/r1d_k_search_fused_chess_ncu.txt, r1b_k_search_fused_c4_ncu.txt)
    ncu_bytes_per_sim = 80.0 if chess else 148.0
    roofline_tree = {"bound": "hbm", "achieved": tree_bytes_step * args.steps / (tree_ms * 1e-3) / 1e9 if tree_ms else None,
                     "peak": hbm_peak, "unit": "GB/s",
                     "traffic": None if use_net else ncu_bytes_per_sim * sims_done,
                     "traffic_source": "ncu capture of the same kernel (c4 8192 trees, chess 16384 trees x 800 sims), scaled per simulation", "launches": n_tree,
                     "avg_launch_ms": tree_ms / max(1, n_tree), "share_of_step": tree_ms / ms}
    if roofline_tree["achieved"]:
        roofline_tree["frac"] = roofline_tree["achieved"] / hbm_peak
    if use_net:
        fl = sims_done * flops_leaf * args.steps
        ach = fl / (net_ms * 1e-3) / 1e12
        roofline = {"bound": "tensor", "achieved": ach, "peak": tensor_peak, "unit": "TFLOP/s", "frac": ach / tensor_peak,
                    # dram__bytes_read+write of one `ncu --set full` capture of k_value_tower per leaf
                    # (profiles/r1d_k_value_tower_{c4,chess}_ncu.txt): planes in, values out, weights from L2
                    "traffic": TOWER_DRAM_BYTES_PER_LEAF[wl["game"]] * sims_done / (sims // BATCH + (1 if sims % BATCH else 0)),
                    "traffic_source": "ncu capture of k_value_tower at 131072 leaves, scaled per leaf",
                    "kernel": "k_value_tower (fused tcgen05 residual tower, bf16 x bf16 -> fp32), %d launches" % n_net,
                    "avg_launch_ms": net_ms / max(1, n_net), "share_of_step": net_ms / ms, "peak_source": peak_src}
    else:
        roofline = dict(roofline_tree, kernel="k_search_fused", peak_source=peak_src)

    line = {
        "metric": "mcts_simulations_per_sec", "value": value, "unit": "sims/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16" if use_net else "f64", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {wl['desc']}", "trees_per_gpu": trees, "sims": sims, "batch_size": BATCH,
                   "c": C_UCT, "policy": "first untried", "roots": "set B (tree_id mod 13 random plies, PCG64(1234+id))",
                   "weights": "random init, torch.manual_seed(0)" if use_net else None,
                   "l2": "working set (node arenas %.0f MB + activations) larger than L2, no flush" % (ts.device_bytes / 1e6),
                   "parallelism": f"{world} independent shards, no collective"},
        "roofline": roofline, "roofline_tree": roofline_tree,
        "e2e": {"value": e2e_value, "unit": "sims/s", "ms_per_step": e2e_ms / args.steps,
                "h2d_bytes_per_step": int(roots.nbytes), "d2h_bytes_per_step": int(sum(v.nbytes for v in out.values() if v is not None))},
        "gpu_launches": int(launches), "clocks": clocks,
        "tree_stats": {"mean_leaf_depth": depth, "nodes_per_tree": cnt["nodes"] / trees, "mean_node_bytes": node_bytes,
                       "algorithmic_bytes_per_sim": tree_bytes_step / sims_done},
    }
    if args.selfplay_games < 0:     # enough games in flight to occupy the GPU, few enough to finish in seconds
        args.selfplay_games = {"c4_value_net": 2048, "c4_heuristic": 16384, "chess_crude": 16384, "chess_value_net": 512}[args.workload]
    if args.selfplay_games > 0:
        # secondary half of the metric: self-play games/hour with the full move loop on the device
        # (search, apply move, win/draw detection, refill); not part of the timed steps above
        from zeroclone_b200.policy_functions import Policy
        from zeroclone_b200.selfplay import DeviceSelfPlay
        from zeroclone_b200.value_functions import Value
        if chess:
            from zeroclone_b200.games.chess import chess_backend as backend
        else:
            from zeroclone_b200.games.connect4 import c4_backend as backend
        vname = "network_latest" if use_net else wl["evaluator"].replace("chess_crude", "crude_chess_score")
        vkw = {"model_type": "chess_value" if chess else "connect4_value"} if use_net else {}
        del ts
        sp = DeviceSelfPlay(backend, Value(vname, **vkw), Policy("random"), n_slots=args.selfplay_games, device=local)
        spo = sp.play(args.selfplay_games, sims, C_UCT, seed=rank, record=False)
        gph = torch.tensor([spo["games_per_hour"]], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(gph, op=dist.ReduceOp.SUM)
        line["selfplay"] = {"games_per_hour": float(gph[0]), "games": args.selfplay_games * world, "sims": sims, "policy": "random",
                            "plies": spo["moves"], "seconds": spo["seconds"], "note": "device-resident loop (zc_search_advance), games in flight = games"}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(args.workload, roots, sims, budget_s=args.cpu_budget)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------
# CPU legs (the only places that may execute oracle/)
# ----------------------------------------------------------------------------------------------
def _port_worker(job):
    """oracle port (C restatement) on one core: searches its share of roots, returns sims done."""
    workload, rows, sims, deadline = job
    import numpy as np
    from oracle import zc_oracle as zo
    wl = WORKLOADS[workload]
    ext = None
    if wl["game"] == "chess":
        done = 0
        for raw in rows:
            st = zo.ChState.from_buffer_copy(raw)
            zo.search(zo.GAME_CHESS, st, sims, C_UCT, BATCH, zo.EVAL_CHESS_CRUDE, zo.POLICY_FIRST)
            done += sims
            if time.time() > deadline:
                break
        return done
    if wl["evaluator"] == "value_net":
        import torch
        torch.set_num_threads(1)
        from zeroclone_b200.models.connect4_value.network import ValueNetwork
        torch.manual_seed(0)
        model = ValueNetwork().eval()

        def ext(states_u8):
            k = states_u8.shape[0]
            cells = states_u8[:, :42].reshape(k, 6, 7)
            turn = states_u8[:, 44]
            cur = np.where(turn == 0, ord('X'), ord('O'))[:, None, None]
            opp = np.where(turn == 0, ord('O'), ord('X'))[:, None, None]
            with torch.no_grad():
                return model(torch.from_numpy(np.stack([cells == cur, cells == opp], axis=1).astype(np.float32))).view(-1).double().numpy()
    done = 0
    for x, o, t in rows:
        from zeroclone_b200.search import c4_unpack_rows
        st = zo.c4_from_rows(c4_unpack_rows(x, o), t)
        zo.search(zo.GAME_C4, st, sims, C_UCT, BATCH, zo.EVAL_EXTERNAL if ext else zo.EVAL_C4_POSITIONAL, zo.POLICY_FIRST, external=ext)
        done += sims
        if time.time() > deadline:
            break
    return done


def _ref_worker(job):
    """the UNMODIFIED reference search (oracle/_ref/mcts*.so) on one core"""
    workload, rows, sims, deadline = job
    from oracle import ref_harness as rh
    mcts, ref_chess = rh.ref_modules()
    wl = WORKLOADS[workload]
    if wl["game"] == "chess":
        import torch
        torch.set_num_threads(1)
        if wl["evaluator"] == "value_net":
            from zeroclone_b200.models.chess_value.network import ValueNetwork
            torch.manual_seed(0)
            value = rh.TorchValue(ValueNetwork())
        else:
            pv = {'P': 1, 'N': 3, 'B': 3, 'R': 5, 'Q': 9, 'p': -1, 'n': -3, 'b': -3, 'r': -5, 'q': -9}

            def crude(s, b):   # engine/value_functions.py:49-55
                if b.check_win(s):
                    return 1000
                return (s.turn * -2 + 1) * sum(pv.get(chr(p), 0) for p in s.board)
            value = rh.FnValue(crude)
        done = 0
        for raw in rows:
            board = list(raw[:64])
            st = ref_chess.State(board, raw[64], raw[65], bool(raw[66]), bool(raw[67]), bool(raw[68]), bool(raw[69]), [], [])
            mcts.get_move(st, value, rh.first_policy, ref_chess, sims, C_UCT, BATCH)
            done += sims
            if time.time() > deadline:
                break
        return done
    if wl["evaluator"] == "value_net":
        import torch
        torch.set_num_threads(1)
        from zeroclone_b200.models.connect4_value.network import ValueNetwork
        torch.manual_seed(0)
        value = rh.TorchValue(ValueNetwork())
    else:
        w = [1, 2, 3, 4, 3, 2, 1]

        def positional(s, b):
            if b.check_win(s):
                return -1
            cur = "XO"[s.turn]
            return sum((w[c] if cell == cur else -w[c]) for row in s.board for c, cell in enumerate(row) if cell != " ") / 64
        value = rh.FnValue(positional)
    done = 0
    for x, o, t in rows:
        mcts.get_move(rh.C4Backend.from_bits(x, o, t), value, rh.first_policy, rh.C4Backend, sims, C_UCT, BATCH)
        done += sims
        if time.time() > deadline:
            break
    return done


def _cpu_pool_run(worker, workload, roots, sims, budget_s, trees_per_core=None):
    """Each of P worker processes loops over its own roots for ~budget_s seconds of wall clock;
    returns (sims/s aggregate, cores, sample description)."""
    import multiprocessing as mp
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    if WORKLOADS[workload]["game"] == "chess":
        rows = [r.tobytes() for r in roots]
    else:
        rows = [(int(r["x"]), int(r["o"]), int(r["turn"])) for r in roots]
    per = max(1, len(rows) // cores)
    ctx = mp.get_context("fork")
    t0 = time.time()
    deadline = t0 + budget_s
    jobs = [(workload, rows[i * per:(i + 1) * per] or rows[:1], sims, deadline) for i in range(cores)]
    with ctx.Pool(cores) as pool:
        done = pool.map(worker, jobs)
    dt = time.time() - t0
    return sum(done) / dt, cores, f"{sum(done) // sims} trees x {sims} sims of the same root set in {dt:.1f} s wall on {cores} processes"


def cpu_baseline(workload, roots, sims, budget_s=15.0):
    from oracle.ref_harness import ref_available
    try:
        if ref_available():
            v, cores, sample = _cpu_pool_run(_ref_worker, workload, roots, sims, budget_s)
            kind = "reference"
        else:
            v, cores, sample = _cpu_pool_run(_port_worker, workload, roots, sims, budget_s)
            kind = "port"
        return {"value": v, "unit": "sims/s", "cores": cores, "kind": kind, "sample": sample}
    except Exception as e:   # the baseline must never take the bench line down
        return {"value": None, "unit": "sims/s", "cores": 0, "kind": "port", "sample": f"failed: {e!r}"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np  # noqa: F401
    from oracle.ref_harness import ref_available
    from zeroclone_b200.workloads import c4_roots_set_b, chess_roots_set_b
    wl = WORKLOADS[args.workload]
    trees = args.trees or wl["trees"]
    sims = args.sims or wl["sims"]
    roots = (chess_roots_set_b if wl["game"] == "chess" else c4_roots_set_b)(min(trees, 4096))
    worker, kind = (_ref_worker, "reference") if ref_available() else (_port_worker, "port")
    per_step = max(2.0, min(20.0, 120.0 / (args.steps + args.warmup)))
    vals, sample, cores = [], "", 0
    for step in range(args.warmup + args.steps):
        v, cores, sample = _cpu_pool_run(worker, args.workload, roots, sims, per_step)
        if step >= args.warmup:
            vals.append(v)
    value = sum(vals) / len(vals)
    line = {"impl": "reference", "metric": "mcts_simulations_per_sec", "value": value, "unit": "sims/s",
            "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if wl["evaluator"] == "value_net" else "f64", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {wl['desc']}", "sims": sims, "batch_size": BATCH, "c": C_UCT,
                       "note": "reference mcts.get_move (engine/mcts/src/mcts.cpp, unmodified, compiled into oracle/_ref) "
                               "driven per tree on all host cores; each step is a bounded sample"},
            "cpu_baseline": {"value": value, "unit": "sims/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "sims/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4_value_net", choices=sorted(WORKLOADS))
    ap.add_argument("--trees", type=int, default=0, help="trees per GPU (default: the workload's)")
    ap.add_argument("--sims", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--selfplay-games", type=int, default=-1,
                    help="games (all in flight at once) of the games/hour side measurement; 0 = skip, -1 = per-workload default")
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
