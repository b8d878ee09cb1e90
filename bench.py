#!/usr/bin/env python
"""bench.py -- MCTS simulations/sec on B200 (BASELINE.json metric), one JSON line on rank 0.

  python bench.py --gpus N --steps K --warmup W [--workload NAME] [--impl reference] [--no-extra]

A step = one pass of the hot path over one batch of synthetic roots: `trees` independent searches of `sims`
simulations each (Engine.play_mcts_parallel's search phase: select, expand, evaluate, backprop, root readout).
The headline line is BASELINE.json configs[1]: Connect Four, 4096 concurrent trees x 800 sims with the value network
(random-init weights, bf16 operands / fp32 accumulation), roots = set B.  Per-GPU work is fixed as N grows (weak scaling): every rank owns its own
trees; the search needs no collective.

The same invocation also measures the other configurations of BASELINE.json and reports them under "workloads":
  c4_heuristic      the deterministic (parity) evaluator, 32768 trees x 800 sims per GPU -- the north-star target
  chess_crude       configs[4], 16384 trees x 1600 sims per GPU (weak) and, at N > 1, "chess_crude_strong":
                    16384 trees in total split over the N GPUs
  chess_value_net   configs[3], 2048 trees x 800 sims per GPU
  c4_selfplay_train (N > 1) configs[2]: 4096 self-play games per GPU x 800 sims with the value net, then the
                    data-parallel training step with its NCCL gradient all-reduce, timed
each with value, e2e, roofline, clocks (and cpu_baseline at N = 1).

Keys beyond the base contract: roofline (dominant kernel), roofline_tree (our tree kernels), cpu_baseline (the
reference's CPU path timed on this box's host cores, N=1 only), e2e (host buffers in, host results out), gpu_launches,
clocks, selfplay (games/hour, device-resident move loop).
`--impl reference` times the reference's own CPU implementation (oracle/_ref/pyref.zip: its compiled mcts / chess modules
and its stock value_functions.py / c4_backend.py) on every host core over a bounded sample of the same workload; the
clock runs inside persistent workers after their imports and model construction (oracle/ref_harness.py).
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
    # name: game, evaluator, trees per GPU, sims; sub_steps = timed steps when run as a side record (the region
    # should last a few hundred ms: ms-scale steps need many)
    "c4_value_net": dict(game="connect4", evaluator="value_net", trees=4096, sims=800, sub_steps=5,
                         desc="BASELINE configs[1]: Connect Four batched self-play search, 4096 concurrent trees x 800 sims, "
                              "value tower 128ch x 8 blocks on 2x6x7 planes"),
    "c4_heuristic": dict(game="connect4", evaluator="c4_positional", trees=32768, sims=800, sub_steps=200,
                         desc="Connect Four, deterministic evaluator (parity configuration), 32768 trees x 800 sims"),
    "chess_crude": dict(game="chess", evaluator="chess_crude", trees=16384, sims=1600, sub_steps=40,
                        desc="BASELINE configs[4]: chess configs/crude_chess.yaml heuristic evaluator, 16384 trees x 1600 sims per GPU"),
    "chess_value_net": dict(game="chess", evaluator="value_net", trees=2048, sims=800, sub_steps=5,
                            desc="BASELINE configs[3]: chess configs/chess_value.yaml, movegen kernel + value-net leaf batching, "
                                 "2048 trees x 800 sims"),
}
# the opt-in PUCT selection mode (stored priors, virtual loss; csrc/puct.cuh) on the c4_heuristic workload: a side record only,
# the reference has no such mode and therefore no baseline for it
PUCT_SIDE = dict(base="c4_heuristic", sub_steps=20, virtual_loss=1.0, prior_weight=0)
C_UCT, BATCH = 1.4, 32
NVLINK_GBS_PER_DIR = 900.0      # NVLink 5 per GPU and direction (nominal), the denominator of the all-reduce figure

# dram__bytes_read.sum + dram__bytes_write.sum from `ncu --set full` captures kept under profiles/ (NOT re-measured in this
# run: a profiler cannot run inside the timed region).  Tower: per evaluated leaf; tree kernels: per simulation.
NCU_TRAFFIC = {
    "tower": {"connect4": (205.4, "profiles/r2_k_value_tower_c4_ncu.txt"), "chess": (2337.7, "profiles/r2_k_value_tower_chess_ncu.txt")},
    "tree": {"connect4": (152.4, "profiles/r2_k_search_fused_c4_ncu.txt"), "chess": (91.2, "profiles/r2h_k_search_fused_chess_ncu.txt")},
}


def workload_config(name: str, trees: int, sims: int, world: int, scaling: str = "weak") -> dict:
    """The `config` object of a bench line; both arms print exactly this for the same arguments."""
    wl = WORKLOADS[name]
    return {"workload": f"{name}: {wl['desc']}", "trees_per_gpu": trees, "sims": sims, "batch_size": BATCH, "c": C_UCT,
            "policy": "first untried", "roots": "set B (tree_id mod 13 random plies, PCG64(1234+id))",
            "weights": "random init, torch.manual_seed(0)" if wl["evaluator"] == "value_net" else None,
            "l2": "inputs larger than L2 (node arenas of 0.5-10 GB per GPU are rebuilt every step), no flush",
            "parallelism": f"{world} independent shards, no collective ({scaling} scaling)"}


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
class Ctx:
    """per-process state shared by every measured workload"""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the engine has no CPU path")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.hbm_peak, self.tensor_peak, self.peak_src = read_peaks()
        self.net_dtype = "auto"     # tensor-core operand format of the value tower: the engine's per-game default, "f16" or "bf16"

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, vals):
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t]

    def gather_floats(self, v: float):
        """v of every rank, in rank order"""
        if self.world == 1:
            return [float(v)]
        out = [self.torch.zeros(1, dtype=self.torch.float64, device=self.dev) for _ in range(self.world)]
        self.dist.all_gather(out, self.torch.tensor([v], dtype=self.torch.float64, device=self.dev))
        return [float(t[0]) for t in out]

    def sum_over_ranks(self, vals):
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [float(v) for v in t]


def measure_search(cx: Ctx, name: str, trees: int, sims: int, steps: int, warmup: int, scaling: str = "weak",
                   first_tree_id: int | None = None, puct: dict | None = None):
    """Time `steps` searches of `trees` trees x `sims` sims on this rank (all ranks together: world x trees).
    Returns (record, roots) -- the record carries value, e2e, roofline, roofline_tree, clocks, gpu_launches."""
    import numpy as np

    from zeroclone_b200 import _ffi
    from zeroclone_b200.evaluator import NetEvaluator, tower_flops_per_leaf
    from zeroclone_b200.search import TreeSearch
    from zeroclone_b200.workloads import c4_roots_set_b, chess_roots_set_b

    torch = cx.torch
    wl = WORKLOADS[name]
    chess = wl["game"] == "chess"
    use_net = wl["evaluator"] == "value_net"
    first = cx.rank * trees if first_tree_id is None else first_tree_id
    roots = (chess_roots_set_b if chess else c4_roots_set_b)(trees, first_tree_id=first)
    roots_pinned = torch.from_numpy(roots.view(np.uint8).reshape(trees, -1).copy()).pin_memory()
    roots_host = roots_pinned.numpy().view(roots.dtype).reshape(trees)       # the e2e input: pinned host memory
    roots_dev = roots_pinned.to(cx.dev)
    ts = TreeSearch(_ffi.GAME_CHESS if chess else _ffi.GAME_C4, trees, sims, device=cx.local)
    if puct is not None:
        ts.set_mode(_ffi.SELECT_PUCT, puct["virtual_loss"], puct["prior_weight"])
    ev, flops_leaf = None, 0.0
    if use_net:
        if chess:
            from zeroclone_b200.models.chess_value.network import ValueNetwork
        else:
            from zeroclone_b200.models.connect4_value.network import ValueNetwork
        torch.manual_seed(0)
        want = {"bf16": torch.bfloat16, "f16": torch.float16, "auto": None}[cx.net_dtype]
        ev = NetEvaluator(ValueNetwork().eval(), cx.dev, want)      # fused sm_100a tower kernel (csrc/tower.cuh)
        net_dtype = "bf16" if ev.dtype == torch.bfloat16 else "f16"
        plane_code = _ffi.PLANE_BF16 if ev.dtype == torch.bfloat16 else _ffi.PLANE_F16
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

    def run_network_timed():
        planes, values = ts._planes, ts._values
        ts.begin(sims, C_UCT, BATCH, _ffi.POLICY_FIRST, 0)
        while ts.pending() > 0:
            timed("tree", lambda: ts.select(planes.data_ptr(), plane_code, stream))
            timed("net", lambda: ev(planes, out=values))
            timed("tree", lambda: ts.backprop(values.data_ptr(), stream))

    def search_resident(record):
        """inputs already in HBM; the root readout (chosen move per tree) is copied to pinned host memory beside the
        next step's kernels (zc_search_results_begin) and collected after the last step"""
        ts.set_roots_dev(roots_dev.data_ptr(), trees, stream)
        if use_net:
            if record:
                run_network_timed()
            else:
                ts.run_network(ev, sims, C_UCT, BATCH, _ffi.POLICY_FIRST)
        elif record:
            timed("tree", lambda: ts.run(sims, C_UCT, BATCH, heur, _ffi.POLICY_FIRST, 0, stream))
        else:
            ts.run(sims, C_UCT, BATCH, heur, _ffi.POLICY_FIRST, 0, stream)
        ts.results_begin(stream)

    def search_e2e():
        """the public call with HOST buffers: roots H2D from pinned memory, search, per-tree results D2H, host waits"""
        ts.set_roots(roots_host, stream)
        if use_net:
            ts.run_network(ev, sims, C_UCT, BATCH, _ffi.POLICY_FIRST)
        else:
            ts.run(sims, C_UCT, BATCH, heur, _ffi.POLICY_FIRST, 0, stream)
        return ts.results(stats=False, stream=stream, reuse=True)    # what get_move returns: the chosen move per tree (mcts.cpp:157-159)

    for _ in range(max(warmup, 3)):
        search_resident(False)
    ts.results_end()
    launches0 = ts.counters()["kernel_launches"] + (ev.launches if ev is not None else 0)
    tower0 = ev.launches if ev is not None else 0
    sampler = ClockSampler(cx.local)
    cx.barrier()
    sampler.start()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(steps):
        search_resident(True)
    res = ts.results_end()
    t1.record()
    cx.barrier()
    clocks = sampler.stop()
    ms = t0.elapsed_time(t1)
    cnt = ts.counters()
    launches = cnt["kernel_launches"] + (ev.launches if ev is not None else 0) - launches0
    assert int(res["result"]["root_visits"].min()) == sims, "a tree did not finish its simulations"
    if ev is not None:
        per_step = -(-sims // BATCH)
        assert ev.launches - tower0 == steps * per_step, f"tower launches {ev.launches - tower0} != steps x ceil(sims/batch) = {steps * per_step}"

    # e2e: same metric through host buffers
    for _ in range(2):
        search_e2e()
    cx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = search_e2e()
    e1.record()
    cx.barrier()
    e2e_ms = e0.elapsed_time(e1)
    ms_by_rank = cx.gather_floats(ms / steps)
    ms, e2e_ms = cx.max_over_ranks([ms, e2e_ms])
    total_sims = cx.world * trees * sims * steps
    value = total_sims / (ms * 1e-3)
    e2e_value = total_sims / (e2e_ms * 1e-3)

    tree_ms = sum(a.elapsed_time(b) for k, a, b in phase_events if k == "tree")
    net_ms = sum(a.elapsed_time(b) for k, a, b in phase_events if k == "net")
    n_tree = sum(1 for k, _, _ in phase_events if k == "tree")
    n_net = sum(1 for k, _, _ in phase_events if k == "net")
    local_ms = t0.elapsed_time(t1)
    # algorithmic bytes of the tree kernels (Roofline accounting): per simulation one new node is written
    # (header + state + edges) plus its link (4 B) and first statistics (16 B edge + 4 B N); per batch one descent
    # reads every node on the path and the backprop rewrites one edge (16 B) and one N (4 B) per level.
    sims_done = trees * sims
    depth = cnt["sum_leaf_depth"] / max(1, cnt["simulations"])
    node_bytes = 16.0 * cnt["arena_slots_used"] / max(1, cnt["nodes"])       # measured mean node size
    batches = sims_done / BATCH
    tree_bytes_step = sims_done * (node_bytes + 4 + 20) + batches * (depth + 1) * (node_bytes + 2 * 20)
    tr_b, tr_src = NCU_TRAFFIC["tree"][wl["game"]]
    roofline_tree = {"bound": "hbm", "achieved": tree_bytes_step * steps / (tree_ms * 1e-3) / 1e9 if tree_ms else None,
                     "peak": cx.hbm_peak, "unit": "GB/s",
                     "traffic": None if use_net else tr_b * sims_done,
                     "traffic_source": f"ncu --set full capture {tr_src} (dram bytes per simulation x simulations per launch); not re-measured in this run",
                     "launches": n_tree, "avg_launch_ms": tree_ms / max(1, n_tree), "share_of_step": tree_ms / local_ms}
    if roofline_tree["achieved"]:
        roofline_tree["frac"] = roofline_tree["achieved"] / cx.hbm_peak
    if use_net:
        fl = sims_done * flops_leaf * steps
        ach = fl / (net_ms * 1e-3) / 1e12
        tw_b, tw_src = NCU_TRAFFIC["tower"][wl["game"]]
        roofline = {"bound": "tensor", "achieved": ach, "peak": cx.tensor_peak, "unit": "TFLOP/s", "frac": ach / cx.tensor_peak,
                    "traffic": tw_b * trees * BATCH,
                    "traffic_source": f"ncu --set full capture {tw_src} (dram bytes per leaf x leaves per launch); not re-measured in this run",
                    "kernel": "k_value_tower (fused tcgen05 residual tower, %s x %s -> fp32), %d launches of %d leaves" % (net_dtype, net_dtype, n_net, trees * BATCH),
                    "avg_launch_ms": net_ms / max(1, n_net), "share_of_step": net_ms / local_ms, "peak_source": cx.peak_src}
    else:
        roofline = dict(roofline_tree, kernel="%s (%d simulations per launch)" % ("k_search_fused_puct" if puct else "k_search_fused", sims_done),
                        peak_source=cx.peak_src)
        if puct:            # no ncu capture of this kernel; its algorithmic bytes are counted like the UCB1 kernel's (one descent per simulation more)
            roofline["traffic"], roofline["traffic_source"] = None, "not captured"
    rec = {"value": value, "unit": "sims/s", "ms_per_step": ms / steps, "ms_per_step_by_rank": ms_by_rank, "steps": steps, "scaling": scaling,
           "dtype": net_dtype if use_net else "f64",
           "config": workload_config(name, trees, sims, cx.world, scaling),
           "roofline": roofline, "roofline_tree": roofline_tree,
           "e2e": {"value": e2e_value, "unit": "sims/s", "ms_per_step": e2e_ms / steps, "h2d_bytes_per_step": int(roots.nbytes),
                   "d2h_bytes_per_step": int(sum(v.nbytes for v in out.values() if v is not None))},
           "gpu_launches": int(launches), "clocks": clocks, "device_bytes": int(ts.device_bytes),
           "tree_stats": {"mean_leaf_depth": depth, "nodes_per_tree": cnt["nodes"] / trees, "mean_node_bytes": node_bytes,
                          "algorithmic_bytes_per_sim": tree_bytes_step / sims_done}}
    ts.close()
    if ev is not None:
        ev.close()
    return rec, roots


def measure_selfplay(cx: Ctx, name: str, games: int, sims: int) -> dict:
    """secondary half of the metric: self-play games/hour with the full move loop on the device (search, apply move,
    win/draw detection, refill); not part of the timed steps"""
    from zeroclone_b200.policy_functions import Policy
    from zeroclone_b200.selfplay import DeviceSelfPlay
    from zeroclone_b200.value_functions import Value
    wl = WORKLOADS[name]
    chess, use_net = wl["game"] == "chess", wl["evaluator"] == "value_net"
    if chess:
        from zeroclone_b200.games.chess import chess_backend as backend
    else:
        from zeroclone_b200.games.connect4 import c4_backend as backend
    vname = "network_latest" if use_net else wl["evaluator"].replace("chess_crude", "crude_chess_score")
    vkw = {"model_type": "chess_value" if chess else "connect4_value"} if use_net else {}
    cx.torch.manual_seed(0)          # the same random-init network as the timed search ("weights" in config)
    sp = DeviceSelfPlay(backend, Value(vname, **vkw), Policy("random"), n_slots=games, device=cx.local)
    spo = sp.play(games, sims, C_UCT, seed=cx.rank, record=False)
    gph = cx.sum_over_ranks([spo["games_per_hour"]])[0]
    return {"games_per_hour": gph, "games": games * cx.world, "sims": sims, "policy": "random", "plies": spo["moves"],
            "seconds": spo["seconds"], "note": "device-resident loop (zc_search_advance), games in flight = games"}


def measure_selfplay_train(cx: Ctx, games: int, sims: int, max_steps: int) -> dict:
    """BASELINE configs[2]: Connect Four self-play sharded over the GPUs (`games` per GPU, value net, `sims` sims per
    move) + the training step of scripts/train.py with its NCCL gradient all-reduce (one 9.53 MB bucket per step)."""
    from zeroclone_b200 import mcts
    from zeroclone_b200.games.connect4 import c4_backend as backend
    from zeroclone_b200.policy_functions import Policy
    from zeroclone_b200.selfplay import DeviceSelfPlay
    from zeroclone_b200.training import train_epochs
    from zeroclone_b200.value_functions import Value
    torch = cx.torch
    torch.manual_seed(0)
    value = Value("network_latest", model_type="connect4_value")
    cx.barrier()
    t0 = time.perf_counter()
    sp = DeviceSelfPlay(backend, value, Policy("random"), n_slots=games, device=cx.local).play(games, sims, C_UCT, seed=1000 + cx.rank)
    cx.barrier()
    sp_s = cx.max_over_ranks([time.perf_counter() - t0])[0]
    states, labels = sp["dataset"]
    mcts.release_all()
    cx.barrier()
    t1 = time.perf_counter()
    st = train_epochs(value.model, states, labels, epochs=1, lr=3e-4, batch_size=256, device=cx.dev, rank=cx.rank,
                      timed=cx.world > 1, verbose=False, max_steps=max_steps)
    value.model.eval()
    value.refresh()
    cx.barrier()
    tr_s = cx.max_over_ranks([time.perf_counter() - t1])[0]
    positions = cx.sum_over_ranks([len(labels)])[0]
    rec = {"config": f"BASELINE configs[2]: Connect Four self-play sharded over {cx.world} GPUs ({games * cx.world} games, {sims} sims per move, "
                     "value net, random expansion policy) + data-parallel training step (Adam, MSE, batch 256 per GPU)",
           "games": games * cx.world, "sims": sims, "selfplay_seconds": sp_s, "games_per_hour": games * cx.world / sp_s * 3600.0,
           "plies": int(cx.sum_over_ranks([sp["moves"]])[0]), "sims_per_sec": cx.sum_over_ranks([sp["moves"]])[0] * sims / sp_s,
           "positions": int(positions), "train_steps": st["steps"], "train_step_ms": cx.max_over_ranks([st["step_ms"]])[0],
           "train_seconds": tr_s, "train_loss": st["loss"], "allreduce_bytes": st["allreduce_bytes"], "collectives_per_step": 1}
    if cx.world > 1:
        n = cx.world
        # (a) inside the training step: CUDA events around every all-reduce on every rank.  A rank that reaches the
        # collective early also waits for the others there, so the collective itself is the per-call MINIMUM over ranks.
        per_call = cx.torch.tensor(st.get("allreduce_ms_list", [0.0]), dtype=cx.torch.float64, device=cx.dev)
        cx.dist.all_reduce(per_call, op=cx.dist.ReduceOp.MIN)
        in_step = float(per_call.median())
        waited = cx.max_over_ranks([st.get("allreduce_ms_median", 0.0)])[0]
        # (b) the same 9.46 MB bucket all-reduced back to back with nothing else on the GPU: the wire time
        bucket = cx.torch.zeros(st["allreduce_bytes"] // 4, dtype=cx.torch.float32, device=cx.dev)
        for _ in range(5):
            cx.dist.all_reduce(bucket)
        cx.barrier()
        a, b = cx.torch.cuda.Event(enable_timing=True), cx.torch.cuda.Event(enable_timing=True)
        reps = 50
        a.record()
        for _ in range(reps):
            cx.dist.all_reduce(bucket)
        b.record()
        cx.torch.cuda.synchronize()
        alone = cx.max_over_ranks([a.elapsed_time(b) / reps])[0]
        gbs = lambda ms: st["allreduce_bytes"] / (ms * 1e-3) / 1e9
        rec.update({"allreduce_ms_in_step": in_step, "allreduce_ms_in_step_incl_rank_skew": waited, "allreduce_ms_back_to_back": alone,
                    "allreduce_algbw_GBs": gbs(alone), "allreduce_busbw_GBs": gbs(alone) * 2 * (n - 1) / n,
                    "allreduce_in_step_algbw_GBs": gbs(in_step) if in_step > 0 else None,
                    "nvlink_GBs_per_direction_nominal": NVLINK_GBS_PER_DIR,
                    "allreduce_share_of_train_step": in_step / cx.max_over_ranks([st["step_ms"]])[0]})
    return rec


def run_ours(args):
    cx = Ctx()
    cx.net_dtype = args.net_dtype
    wl = WORKLOADS[args.workload]
    trees = args.trees or wl["trees"]
    sims = args.sims or wl["sims"]
    rec, roots = measure_search(cx, args.workload, trees, sims, args.steps, args.warmup)
    line = {"metric": "mcts_simulations_per_sec", "value": rec["value"], "unit": "sims/s", "n_gpus": cx.world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": rec["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": rec["dtype"], "data": "synthetic", "config": rec["config"],
            "roofline": rec["roofline"], "roofline_tree": rec["roofline_tree"], "e2e": rec["e2e"],
            "gpu_launches": rec["gpu_launches"], "clocks": rec["clocks"], "tree_stats": rec["tree_stats"]}
    if args.selfplay_games < 0:     # enough games in flight to occupy the GPU, few enough to finish in seconds
        args.selfplay_games = {"c4_value_net": 2048, "c4_heuristic": 16384, "chess_crude": 16384, "chess_value_net": 512}[args.workload]
    if args.selfplay_games > 0:
        line["selfplay"] = measure_selfplay(cx, args.workload, args.selfplay_games, sims)
    solo = cx.rank == 0 and cx.world == 1 and not args.no_cpu_baseline
    if solo:
        line["cpu_baseline"] = cpu_baseline(args.workload, sims, budget_s=args.cpu_budget)
    if not args.no_extra:
        subs = {}
        for name in WORKLOADS:
            if name == args.workload:
                continue
            w = WORKLOADS[name]
            try:
                r, _ = measure_search(cx, name, w["trees"], w["sims"], w["sub_steps"], 3)
                if solo:
                    r["cpu_baseline"] = cpu_baseline(name, w["sims"], budget_s=max(4.0, args.cpu_budget / 2))
                subs[name] = r
            except Exception as e:      # noqa: BLE001 -- a side record must never take the headline down
                subs[name] = {"failed": repr(e)}
                if cx.world > 1:
                    raise               # ranks would desynchronise: fail loudly instead
        try:
            w = WORKLOADS[PUCT_SIDE["base"]]
            r, _ = measure_search(cx, PUCT_SIDE["base"], w["trees"], w["sims"], PUCT_SIDE["sub_steps"], 3, puct=PUCT_SIDE)
            r["config"]["select"] = "PUCT, stored priors (uniform for Connect Four), virtual loss %.1f; opt-in mode, not in the reference" % PUCT_SIDE["virtual_loss"]
            subs["c4_heuristic_puct"] = r
        except Exception as e:      # noqa: BLE001
            subs["c4_heuristic_puct"] = {"failed": repr(e)}
            if cx.world > 1:
                raise
        if cx.world > 1:
            w = WORKLOADS["chess_crude"]
            per = w["trees"] // cx.world
            r, _ = measure_search(cx, "chess_crude", per, w["sims"], w["sub_steps"], 3, scaling="strong", first_tree_id=cx.rank * per)
            r["total_trees"] = per * cx.world
            subs["chess_crude_strong"] = r
            subs["c4_selfplay_train"] = measure_selfplay_train(cx, args.train_games, WORKLOADS["c4_value_net"]["sims"], args.train_steps)
        elif args.train_record:
            subs["c4_selfplay_train"] = measure_selfplay_train(cx, args.train_games, WORKLOADS["c4_value_net"]["sims"], args.train_steps)
        line["workloads"] = subs
    if cx.rank == 0:
        print(json.dumps(line))
    if cx.world > 1:
        cx.dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------
# CPU legs (the only places that may execute oracle/)
# ----------------------------------------------------------------------------------------------
def _ref_rows(name: str, n: int):
    """the first n roots of the workload's root set, generated by the reference's own backends (no libzc_b200)"""
    from oracle import ref_harness as rh
    return (rh.chess_roots_set_b if WORKLOADS[name]["game"] == "chess" else rh.c4_roots_set_b)(n)


def _port_rate(name: str, sims: int, budget_s: float):
    """fallback when oracle/_ref is absent: the C restatement (oracle/libzc_oracle.so) on one core per process"""
    import multiprocessing as mp
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    with mp.get_context("spawn").Pool(cores) as pool:
        res = pool.map(_port_worker, [(name, i, sims, budget_s) for i in range(cores)])
    return sum(d / dt for d, dt in res), cores, sum(d for d, _ in res) // sims


def _port_worker(job):
    name, seed, sims, budget = job
    import numpy as np
    from oracle import zc_oracle as zo
    wl = WORKLOADS[name]
    chess = wl["game"] == "chess"
    ext = None
    if wl["evaluator"] == "value_net":      # the port's search around this repo's PyTorch module in fp32 on one core
        import torch
        torch.set_num_threads(1)
        if chess:
            from zeroclone_b200.models.chess_value.network import ValueNetwork
        else:
            from zeroclone_b200.models.connect4_value.network import ValueNetwork
        torch.manual_seed(0)
        model = ValueNetwork().eval()

        def ext(states_u8):
            k = states_u8.shape[0]
            if chess:
                planes = np.zeros((k, 17, 8, 8), dtype=np.float32)
                board = states_u8[:, :64].reshape(k, 8, 8)
                for p_, ch in enumerate(b"PNBRQKpnbrqk"):
                    planes[:, p_] = board == ch
                planes[:, 12] = (states_u8[:, 64] == 0)[:, None, None]
                for f in range(4):
                    planes[:, 13 + f] = (states_u8[:, 66 + f] != 0)[:, None, None]
            else:
                cells = states_u8[:, :42].reshape(k, 6, 7)
                turn = states_u8[:, 44]
                cur = np.where(turn == 0, ord('X'), ord('O'))[:, None, None]
                opp = np.where(turn == 0, ord('O'), ord('X'))[:, None, None]
                planes = np.stack([cells == cur, cells == opp], axis=1).astype(np.float32)
            with torch.no_grad():
                return model(torch.from_numpy(planes)).view(-1).double().numpy()
    ev = zo.EVAL_EXTERNAL if ext else (zo.EVAL_CHESS_CRUDE if chess else zo.EVAL_C4_POSITIONAL)
    done, t0 = 0, time.perf_counter()
    i = seed * 131
    while True:
        st = zo.ch_init() if chess else zo.c4_from_moves([(i + j) % 7 for j in range(i % 9)])
        zo.search(zo.GAME_CHESS if chess else zo.GAME_C4, st, sims, C_UCT, BATCH, ev, zo.POLICY_FIRST, external=ext)
        i += 1
        done += sims
        dt = time.perf_counter() - t0
        if dt >= budget:
            return done, dt


def cpu_baseline(name: str, sims: int, budget_s: float = 15.0) -> dict:
    """The reference's CPU path on this box's host cores over a bounded sample of the workload's root set."""
    from oracle import ref_harness as rh
    wl = WORKLOADS[name]
    try:
        if rh.ref_available():
            pool = rh.RefPool(wl["game"], wl["evaluator"], _ref_rows(name, 1024), sims)
            try:
                pool.step(min(2.0, budget_s / 4))                  # page in, fill allocator pools
                v, det = pool.step(budget_s)
            finally:
                pool.close()
            return {"value": v, "unit": "sims/s", "cores": pool.cores, "kind": "reference",
                    "sample": f"{det['trees']} trees x {sims} sims of the same root set, {budget_s:.0f} s window per worker "
                              f"(longest {det['longest_worker_s']:.1f} s), {pool.cores} processes, clock inside the workers after setup"}
        v, cores, trees = _port_rate(name, sims, budget_s)
        return {"value": v, "unit": "sims/s", "cores": cores, "kind": "port",
                "sample": f"{trees} trees x {sims} sims (oracle C port, built-in evaluator), {budget_s:.0f} s per worker on {cores} processes"}
    except Exception as e:   # noqa: BLE001 -- the baseline must never take the bench line down
        return {"value": None, "unit": "sims/s", "cores": 0, "kind": "reference", "sample": f"failed: {e!r}"}


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation of the path.  No torch, no CUDA, no libzc_b200 in
    this process; the workers import the reference's files from oracle/_ref/pyref.zip (unpacked to a temporary directory)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ref_harness as rh
    wl = WORKLOADS[args.workload]
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    trees = args.trees or wl["trees"]
    sims = args.sims or wl["sims"]
    n_steps = args.warmup + args.steps
    per_step = max(3.0, min(20.0, 150.0 / n_steps))
    base = {"impl": "reference", "metric": "mcts_simulations_per_sec", "unit": "sims/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if wl["evaluator"] == "value_net" else "f64", "data": "synthetic",
            "config": workload_config(args.workload, trees, sims, world)}
    if not rh.ref_available():
        v, cores, n = _port_rate(args.workload, sims, per_step * args.steps)
        kind, sample, ms = "port", f"{n} trees x {sims} sims, oracle C port", per_step * 1e3
    else:
        pool = rh.RefPool(wl["game"], wl["evaluator"], _ref_rows(args.workload, min(trees, 4096)), sims)
        try:
            vals, dets = [], []
            for step in range(n_steps):
                v, det = pool.step(per_step)
                if step >= args.warmup:
                    vals.append(v)
                    dets.append(det)
        finally:
            pool.close()
        v, cores, kind = sum(vals) / len(vals), pool.cores, "reference"
        ms = sum(d["longest_worker_s"] for d in dets) / len(dets) * 1e3
        sample = (f"per step each of {cores} worker processes runs reference get_move on its share of the first {min(trees, 4096)} roots of "
                  f"the set for {per_step:.1f} s (whole trees; {sum(d['trees'] for d in dets)} trees over the timed steps); "
                  "value = sum of the workers' own sims/s, clock inside the workers after imports and model construction")
    line = dict(base, value=v, ms_per_step=ms,
                note="reference mcts.get_move (engine/mcts/src/mcts.cpp compiled unmodified) with the reference's stock Value / backend "
                     "Python files (oracle/_ref/pyref.zip) on every host core, CUDA hidden (DEVICE=cpu, fp32); bounded sample of the workload",
                cpu_baseline={"value": v, "unit": "sims/s", "cores": cores, "kind": kind, "sample": sample},
                e2e={"value": v, "unit": "sims/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
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
    ap.add_argument("--no-extra", action="store_true", help="only the headline workload (no `workloads` side records)")
    ap.add_argument("--selfplay-games", type=int, default=-1,
                    help="games (all in flight at once) of the games/hour side measurement; 0 = skip, -1 = per-workload default")
    ap.add_argument("--train-record", action="store_true", help="also measure configs[2] (self-play + training step) at N = 1")
    ap.add_argument("--train-games", type=int, default=4096, help="self-play games per GPU of the configs[2] record")
    ap.add_argument("--train-steps", type=int, default=200, help="training steps timed in the configs[2] record")
    ap.add_argument("--cpu-budget", type=float, default=15.0)
    ap.add_argument("--net-dtype", default="auto", choices=["auto", "f16", "bf16"],
                    help="tensor-core operand format of the value tower; auto = the engine's default (bf16 Connect Four, fp16 chess: the "
                         "cheapest format that holds the 1e-3 root-value bar for the game)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
