"""TEST/BENCH INFRASTRUCTURE: drive the UNMODIFIED reference on the box's host cores.

`make -C oracle ref` packs the reference's own files of the hot path into oracle/_ref/pyref.zip as its
`engine` and `models` packages: the two pybind11 modules compiled from its C++ sources
(engine/mcts/src/*.cpp, engine/games/chess/src/*.cpp) and, verbatim, engine/value_functions.py,
engine/policy_functions.py, engine/games/connect4/c4_backend.py, engine/mcts/__init__.py and
models/ (build output: git-ignored, unpacked into a temporary directory at run time).  This file
imports THOSE (never this repo's packages, never libzc_b200.so) and runs
`engine.mcts.get_move(state, Value, policy, backend, sims, c, batch)` -- the call
engine/engine.py:119-129 makes -- one tree at a time on every host core.

Nothing is restated here except three things the reference does not ship:
  * a `first`-untried policy callback (the parity configuration's deterministic policy; the stock
    Policy class only has random / immediate_value, policy_functions.py:10-17),
  * `c4_positional`, the deterministic Connect Four parity evaluator, added as a
    method of a subclass of the stock Value so it runs through the stock Value.batch/__call__,
  * a Connect Four value network: the reference has models/chess_value only, so configs[1] uses the
    stock chess_value.ValueNetwork class with its stem convolution re-made for 2 input planes.
The network evaluator is the stock Value._nn_setup / _batch_worker / batch path
(value_functions.py:61-99) on the CPU (CUDA is hidden from the workers: DEVICE="cpu", fp32).

Timing: worker processes are persistent; imports, model construction and root decoding happen
before the clock; each step every worker times its own get_move loop with perf_counter and the
step's value is the sum of the workers' own sims/second (they run concurrently, one per core).

Used only by bench.py (`--impl reference`, the cpu_baseline leg) and tests.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np

_REF_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
PYREF_ZIP = os.path.join(_REF_DIR, "pyref.zip")
PYREF = None            # where the archive is unpacked for this process tree (pyref_dir())
C_UCT, BATCH = 1.4, 32


def ref_available() -> bool:
    return os.path.exists(PYREF_ZIP)


def pyref_dir() -> str:
    """Unpack oracle/_ref/pyref.zip (the reference's own files, packed by `make -C oracle ref`) into a temporary directory
    named after the archive's digest; compiled modules cannot be imported from inside a zip.  Idempotent and race-free."""
    global PYREF
    if PYREF is None:
        import hashlib
        import tempfile
        import zipfile
        digest = hashlib.sha256(open(PYREF_ZIP, "rb").read()).hexdigest()[:16]
        target = os.path.join(tempfile.gettempdir(), f"zc_pyref_{digest}")
        if not os.path.exists(os.path.join(target, "SHA256SUMS")):
            tmp = tempfile.mkdtemp(prefix="zc_pyref_unpack_")
            with zipfile.ZipFile(PYREF_ZIP) as z:
                z.extractall(tmp)
            try:
                os.rename(tmp, target)              # atomic: a concurrent worker either wins or finds it there
            except OSError:
                import shutil
                shutil.rmtree(tmp, ignore_errors=True)
        PYREF = target
    return PYREF


def ref_modules():
    """the two compiled reference modules on their own (tests)"""
    if _REF_DIR not in sys.path:
        sys.path.insert(0, _REF_DIR)
    import chess_backend
    import mcts
    return mcts, chess_backend


class _Stock:
    pass


_stock = None


def stock(need_torch: bool = False) -> _Stock:
    """Import the reference's own `engine` / `models` packages from oracle/_ref/pyref.  This repo ships
    import aliases of the same names at its root (drop-in for users); they must not win here."""
    global _stock
    PYREF = pyref_dir() if ref_available() else ""
    if _stock is None:
        if not ref_available():
            raise ImportError("oracle/_ref/pyref.zip is missing: run `make -C oracle ref` where /root/reference exists")
        for name in list(sys.modules):
            if name in ("engine", "models") or name.startswith(("engine.", "models.")):
                f = getattr(sys.modules[name], "__file__", None) or ""
                if not f.startswith(PYREF):
                    del sys.modules[name]
        sys.path.insert(0, PYREF)
        import importlib
        s = _Stock()
        s.mcts = importlib.import_module("engine.mcts")
        s.c4 = importlib.import_module("engine.games.connect4.c4_backend")
        s.chess = importlib.import_module("engine.games.chess.chess_backend")
        s.Policy = importlib.import_module("engine.policy_functions").Policy
        for m in (s.mcts, s.c4, s.chess):
            assert m.__file__.startswith(PYREF), (m.__name__, m.__file__)
        assert s.mcts.mcts is not None, "reference engine.mcts could not load its compiled module"
        _stock = s
    if need_torch and not hasattr(_stock, "Value"):
        import importlib
        vf = importlib.import_module("engine.value_functions")      # imports torch; DEVICE is decided here
        assert vf.__file__.startswith(PYREF)
        _stock.vf = vf
        _stock.Value = vf.Value
        _stock.network = importlib.import_module("models.chess_value.network")
    return _stock


def first_policy(moves):
    return moves[0]


# ---- root sets (SURVEY.md §8d set B) from the reference's own backends ---------------------------
def c4_roots_set_b(n: int, first_tree_id: int = 0) -> list:
    """[(x_bits, o_bits, turn)]: the same positions as zeroclone_b200.workloads.c4_roots_set_b, produced by
    playing the reference's c4_backend (legal moves in list(set) order, c4_backend.py:49-50)."""
    b = stock().c4
    out = []
    for i in range(n):
        tid = first_tree_id + i
        rng = np.random.Generator(np.random.PCG64(1234 + tid))
        while True:
            s = b.create_init_state()
            for _ in range(tid % 13):
                moves = list(b.get_legal_moves(s))
                s = b.play_move(s, moves[int(rng.integers(len(moves)))])
            if not b.check_win(s) and not b.check_draw(s):
                break
        x = o = 0
        for r in range(6):
            for c in range(7):
                bit = 1 << (c * 7 + (5 - r))
                if s.board[r][c] == "X":
                    x |= bit
                elif s.board[r][c] == "O":
                    o |= bit
        out.append((x, o, int(s.turn)))
    return out


def c4_state_from_bits(x: int, o: int, turn: int):
    b = stock().c4
    board = [[" "] * 7 for _ in range(6)]
    for r in range(6):
        for c in range(7):
            bit = 1 << (c * 7 + (5 - r))
            if x & bit:
                board[r][c] = "X"
            elif o & bit:
                board[r][c] = "O"
    return b.State(board, turn)


def chess_roots_set_b(n: int, first_tree_id: int = 0) -> list:
    """[72-byte packed zc_chess_state]: same positions as zeroclone_b200.workloads.chess_roots_set_b, produced by the
    reference's chess_backend (get_legal_moves / play_move / check_win / check_draw incl. its own histories)."""
    b = stock().chess
    out = []
    for i in range(n):
        tid = first_tree_id + i
        rng = np.random.Generator(np.random.PCG64(1234 + tid))
        while True:
            s = b.create_init_state()
            for _ in range(tid % 13):
                mv = b.get_legal_moves(s)
                if not mv:
                    break
                s = b.play_move(s, mv[int(rng.integers(len(mv)))])
            if not b.check_win(s) and not b.check_draw(s):
                break
        out.append(bytes(s.board) + bytes([s.turn, s.fifty_move_rule_counter, s.w_ck, s.w_cq, s.b_ck, s.b_cq, 0, 0]))
    return out


def chess_state_from_bytes(raw: bytes):
    b = stock().chess
    return b.State(list(raw[:64]), raw[64], raw[65], bool(raw[66]), bool(raw[67]), bool(raw[68]), bool(raw[69]), [], [])


# ---- evaluators --------------------------------------------------------------------------------
def make_value(game: str, evaluator: str, allow_cuda: bool = False):
    """A stock `Value` for the workload's evaluator."""
    if evaluator == "value_net":
        s = stock(need_torch=True)
        import torch
        torch.set_num_threads(1)
        torch.manual_seed(0)
        model = s.network.ValueNetwork()
        if game == "connect4":       # the reference ships no Connect Four network: same tower, 2 input planes
            model.stem[0] = torch.nn.Conv2d(2, 128, 3, padding=1, bias=False)
        v = s.Value.__new__(s.Value)         # init_network_latest (value_functions.py:104-112) minus the checkpoint lookup
        v.name, v.init_args = "network_latest", {"batch_size": BATCH}
        v._nn_setup(model, BATCH)
        assert allow_cuda or v.device == "cpu", "the reference arm is the reference's CPU path"
        return v
    s = stock(need_torch=True)             # value_functions.py imports torch at module level
    if evaluator == "chess_crude":
        return s.Value("crude_chess_score")
    if evaluator == "c4_positional":
        w = (1, 2, 3, 4, 3, 2, 1)

        class ParityValue(s.Value):
            def c4_positional(self, state, args):
                if args["backend"].check_win(state):
                    return -1
                cur = "XO"[state.turn]
                return sum((w[c] if cell == cur else -w[c]) for row in state.board for c, cell in enumerate(row) if cell != " ") / 64
        return ParityValue("c4_positional")
    if evaluator == "random_rollout":
        return s.Value("random_rollout")
    raise ValueError(evaluator)


# ---- persistent worker pool --------------------------------------------------------------------
def _worker_main(conn, game, evaluator, rows, sims, hide_cuda=True):
    try:
        if hide_cuda:
            os.environ["CUDA_VISIBLE_DEVICES"] = ""    # the reference's Value then picks DEVICE="cpu", fp32
        os.environ.setdefault("OMP_NUM_THREADS", "1")
        s = stock(need_torch=True)
        value = make_value(game, evaluator, allow_cuda=not hide_cuda)
        if game == "chess":
            backend = s.chess
            states = [chess_state_from_bytes(r) for r in rows]
        else:
            backend = s.c4
            states = [c4_state_from_bits(*r) for r in rows]
        get_move = s.mcts.get_move
        conn.send(("ready", os.getpid()))
        nxt = 0
        while True:
            msg = conn.recv()
            if msg[0] == "stop":
                break
            budget = msg[1]
            done = 0
            t0 = time.perf_counter()
            while True:
                get_move(states[nxt % len(states)], value, first_policy, backend, sims, C_UCT, BATCH)
                nxt += 1
                done += sims
                dt = time.perf_counter() - t0
                if dt >= budget:
                    break
            conn.send(("done", done, dt))
    except Exception as e:      # noqa: BLE001 -- reported to the parent, which raises
        import traceback
        conn.send(("error", f"{e!r}\n{traceback.format_exc()}"))


class RefPool:
    """`cores` persistent processes, each owning a slice of the root set.  step(seconds) -> (sims/s, detail)."""

    def __init__(self, game: str, evaluator: str, rows: list, sims: int, cores: int | None = None, hide_cuda: bool = True):
        """hide_cuda=False (tools/ref_gpu_net.py only, never the bench arms): the stock Value puts the network on the GPU in
        fp16 (value_functions.py:5-6) -- how a user of the reference would run it on a GPU box."""
        import multiprocessing as mp
        if cores is None:
            cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        self.cores, self.sims = cores, sims
        # spawn, not fork: the parent may hold a CUDA context (cpu_baseline leg of the GPU arm) and this repo's
        # `engine` aliases; the workers start from a clean interpreter with CUDA hidden
        ctx = mp.get_context("spawn")
        per = max(1, len(rows) // cores)
        self.procs, self.conns = [], []
        saved = {k: os.environ.get(k) for k in ("CUDA_VISIBLE_DEVICES", "OMP_NUM_THREADS", "MKL_NUM_THREADS")}
        os.environ.update({"OMP_NUM_THREADS": "1", "MKL_NUM_THREADS": "1"})
        if hide_cuda:
            os.environ["CUDA_VISIBLE_DEVICES"] = ""
        try:
            for i in range(cores):
                a, b = ctx.Pipe()
                mine = rows[i * per:(i + 1) * per] or rows[:1]
                p = ctx.Process(target=_worker_main, args=(b, game, evaluator, mine, sims, hide_cuda), daemon=True)
                p.start()
                self.procs.append(p)
                self.conns.append(a)
        finally:
            for k, v in saved.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v
        for c in self.conns:
            self._expect(c, "ready")

    @staticmethod
    def _expect(conn, what):
        msg = conn.recv()
        if msg[0] == "error":
            raise RuntimeError("reference worker failed: " + msg[1])
        assert msg[0] == what, msg
        return msg

    def step(self, seconds: float):
        for c in self.conns:
            c.send(("go", seconds))
        res = [self._expect(c, "done") for c in self.conns]
        rate = sum(d / dt for _, d, dt in res)
        total = sum(d for _, d, _ in res)
        longest = max(dt for _, _, dt in res)
        return rate, {"sims": total, "trees": total // self.sims, "longest_worker_s": longest}

    def close(self):
        for c in self.conns:
            try:
                c.send(("stop",))
            except Exception:
                pass
        for p in self.procs:
            p.join(timeout=10)
            if p.is_alive():
                p.kill()
