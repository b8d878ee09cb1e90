"""GPU parity for chess: device move generation (ordered lists, perft) and the search, bit-exact
against the golden vectors of the reference and the oracle."""
import ctypes as C

import numpy as np
import pytest

from conftest import load_golden
from oracle import zc_oracle as zo
from zeroclone_b200 import _ffi
from zeroclone_b200.search import TreeSearch

pytestmark = pytest.mark.gpu
POL = {"first": _ffi.POLICY_FIRST, "last": _ffi.POLICY_LAST}


def pack(rec_or_oracle_state):
    a = np.zeros(1, dtype=_ffi.CHESS_STATE_DTYPE)[0]
    s = rec_or_oracle_state
    if isinstance(s, dict):
        a["board"] = np.frombuffer(s["board"].encode(), dtype=np.uint8)
        a["turn"], a["fifty_move_rule_counter"] = s["turn"], s["fifty"]
        a["w_ck"], a["w_cq"], a["b_ck"], a["b_cq"] = s["flags"]
    else:
        a["board"] = np.frombuffer(bytes(s.board), dtype=np.uint8)
        a["turn"], a["fifty_move_rule_counter"] = s.turn, s.fifty
        a["w_ck"], a["w_cq"], a["b_ck"], a["b_cq"] = s.w_ck, s.w_cq, s.b_ck, s.b_cq
    return a


def states_array(items):
    arr = np.zeros(len(items), dtype=_ffi.CHESS_STATE_DTYPE)
    for i, it in enumerate(items):
        arr[i] = pack(it)
    return arr


def device_legal(arr, warp=False):
    """warp=False: one thread per position (generate); True: one warp per position (generate_warp)"""
    n = len(arr)
    moves = np.zeros((n, _ffi.MAX_MOVES), dtype=_ffi.CHESS_MOVE_DTYPE)
    counts = np.zeros(n, dtype=np.int32)
    flags = np.zeros(n, dtype=np.int32)
    fn = _ffi.lib().zc_chess_legal_moves_batch_warp if warp else _ffi.lib().zc_chess_legal_moves_batch
    _ffi.check(fn(0, arr.ctypes.data_as(C.c_void_p), n, moves.ctypes.data_as(C.c_void_p),
                  counts.ctypes.data_as(C.c_void_p), flags.ctypes.data_as(C.c_void_p)))
    return moves, counts, flags


def as_lists(moves, k):
    return [[int(m["fr"]), int(m["fc"]), int(m["tr"]), int(m["tc"]), float(m["value"])] for m in moves[:k]]


@pytest.mark.parametrize("warp", [False, True])
def test_device_movegen_matches_reference_playouts(warp):
    g = load_golden("chess_rules.json.gz")
    recs = [r for trace in g["playouts"] for r in trace] + list(g["fens"].values())
    moves, counts, flags = device_legal(states_array(recs), warp)
    for i, rec in enumerate(recs):
        assert as_lists(moves[i], counts[i]) == rec["legal"], i
        assert bool(flags[i] & 1) == rec["win"], i
        if not rec["legal"] and not rec["win"]:
            assert flags[i] & 2


@pytest.mark.parametrize("warp", [False, True])
def test_device_movegen_matches_oracle_on_random_positions(warp):
    rng = np.random.default_rng(23)
    states = []
    for g in range(40 if not warp else 120):
        s = zo.ch_init()
        for ply in range(int(rng.integers(0, 140))):
            mv = zo.ch_legal(s)
            if not mv:
                break
            s = zo.ch_play(s, mv[int(rng.integers(len(mv)))])
            states.append(s)
    moves, counts, flags = device_legal(states_array(states), warp)
    for i, s in enumerate(states):
        assert as_lists(moves[i], counts[i]) == [list(m[0]) + [m[1]] for m in zo.ch_legal(s)], i
        assert bool(flags[i] & 1) == zo.ch_check_win(s)


def test_device_perft_matches_reference_tables():
    g = load_golden("chess_rules.json.gz")
    for name, counts in g["perft"].items():
        root = states_array([g["fens"][name]])
        for d, want in enumerate(counts, start=1):
            out = C.c_uint64()
            _ffi.check(_ffi.lib().zc_chess_perft(0, root.ctypes.data_as(C.c_void_p), d, C.byref(out)))
            assert out.value == want, (name, d)
    # startpos depth 5 under the reference's rules
    out = C.c_uint64()
    root = states_array([zo.ch_init()])
    _ffi.check(_ffi.lib().zc_chess_perft(0, root.ctypes.data_as(C.c_void_p), 5, C.byref(out)))
    assert out.value == 4865351


def test_golden_chess_search_cases_bit_exact():
    ts = TreeSearch(_ffi.GAME_CHESS, max_trees=2, max_sims=1600)
    for cs in load_golden("chess_search.json"):
        ts.set_roots(states_array([cs]))
        ts.run(cs["sims"], cs["c"], cs["batch"], _ffi.EVAL_CHESS_CRUDE, POL[cs["policy"]])
        out = ts.results()
        r = out["result"][0]
        k = int(r["n_moves"])
        tag = (cs["name"], cs["policy"], cs["sims"], cs["batch"])
        assert as_lists(out["moves"][0], k) == cs["moves"], tag
        assert out["visits"][0][:k].tolist() == cs["Na"], tag
        assert out["value_sums"][0][:k].tolist() == cs["Wa"], tag
        assert int(r["best"]) == cs["best"], tag
        assert int(r["nodes"]) == cs["nodes_created"], tag
        assert int(r["sum_leaf_depth"]) == cs["sum_leaf_depth"] and int(r["max_leaf_depth"]) == cs["max_leaf_depth"], tag


@pytest.mark.parametrize("policy,sims,c,batch", [("first", 1600, 1.4, 32), ("last", 700, 1.4, 32), ("first", 250, 2.2, 9)])
def test_chess_whole_tree_hash_matches_oracle(policy, sims, c, batch):
    rng = np.random.default_rng(sims)
    states = []
    while len(states) < 48:
        s = zo.ch_init()
        for ply in range(int(rng.integers(0, 90))):
            mv = zo.ch_legal(s)
            if not mv:
                break
            s = zo.ch_play(s, mv[int(rng.integers(len(mv)))])
        if zo.ch_legal(s):
            states.append(s)
    n = len(states)
    ts = TreeSearch(_ffi.GAME_CHESS, n, sims)
    ts.set_roots(states_array(states))
    ts.run(sims, c, batch, _ffi.EVAL_CHESS_CRUDE, POL[policy])
    out, hashes = ts.results(), ts.tree_hash()
    opol = {"first": zo.POLICY_FIRST, "last": zo.POLICY_LAST}[policy]
    for i in range(n):
        o = zo.search(zo.GAME_CHESS, states[i], sims, c, batch, zo.EVAL_CHESS_CRUDE, opol)
        k = o.n_moves
        assert out["visits"][i][:k].tolist() == o.Na, i
        assert out["value_sums"][i][:k].tolist() == o.Wa, i
        assert int(out["result"][i]["best"]) == o.best, i
        assert int(hashes[i]) == o.tree_hash, i


def test_chess_split_phase_with_exact_evaluator_matches_fused():
    """external-evaluator kernels on chess: material from the packed planes (exact) == crude score
    wherever no mate occurs; compare with the oracle driven by the same function."""
    import torch
    vals = torch.tensor([1, 3, 3, 5, 9, 0, -1, -3, -3, -5, -9, 0], dtype=torch.float32, device="cuda")

    class Material:
        dtype = torch.bfloat16

        def __call__(self, planes, out):
            mat = (planes[:, :12].float().sum(dim=(2, 3)) * vals).sum(dim=1)
            white_to_move = planes[:, 12, 0, 0].float()
            out.copy_(mat * (2 * white_to_move - 1))
            return out

    def ext(states_u8):
        pv = {ord(k): v for k, v in {'P': 1, 'N': 3, 'B': 3, 'R': 5, 'Q': 9, 'p': -1, 'n': -3, 'b': -3, 'r': -5, 'q': -9}.items()}
        out = []
        for row in states_u8:
            out.append((1 - 2 * int(row[64])) * sum(pv.get(int(b), 0) for b in row[:64]))
        return np.array(out, dtype=np.float64)

    rng = np.random.default_rng(4)
    states = []
    while len(states) < 16:
        s = zo.ch_init()
        for ply in range(int(rng.integers(0, 60))):
            mv = zo.ch_legal(s)
            if not mv:
                break
            s = zo.ch_play(s, mv[int(rng.integers(len(mv)))])
        if zo.ch_legal(s):
            states.append(s)
    ts = TreeSearch(_ffi.GAME_CHESS, len(states), 400)
    ts.set_roots(states_array(states))
    ts.run_network(Material(), 400, 1.4, 32, _ffi.POLICY_FIRST)
    out, hashes = ts.results(), ts.tree_hash()
    for i, s in enumerate(states):
        o = zo.search(zo.GAME_CHESS, s, 400, 1.4, 32, zo.EVAL_EXTERNAL, zo.POLICY_FIRST, external=ext)
        assert out["visits"][i][:o.n_moves].tolist() == o.Na, i
        assert out["value_sums"][i][:o.n_moves].tolist() == o.Wa, i
        assert int(hashes[i]) == o.tree_hash, i


def test_warp_generator_on_crowded_and_odd_boards():
    """Random boards far outside normal play -- up to 30 pieces a side, several kings or none, pawns on the last rows -- so that
    every branch of the warp generator is taken (more than 16 pieces: no half-warp split of the emission; more than 24: the
    king's steps are tested by its own lane; more than 32: the serial fallback; no king -- the reference's phantom king at
    (-1,-1); several kings).  The warp generator,
    the serial generator and the oracle (the pinned restatement of chess_backend.cpp:184-360) must list the same moves in
    the same order."""
    from test_rules_odd_boards import odd_boards
    boards = [{"board": b.decode(), "turn": t, "fifty": 0, "flags": [0, 0, 0, 0]} for b, t in odd_boards()]
    arr = states_array(boards)
    mw, cw, fw = device_legal(arr, warp=True)
    ms, cs, fs = device_legal(arr, warp=False)
    assert cw.tolist() == cs.tolist() and fw.tolist() == fs.tolist()
    crowded = 0
    for i, rec in enumerate(boards):
        st = zo.ChState()
        st.board[:] = rec["board"].encode()
        st.turn = rec["turn"]
        want = [[m[0][0], m[0][1], m[0][2], m[0][3], float(m[1])] for m in zo.ch_legal(st)]
        assert as_lists(mw[i], cw[i]) == want, i
        assert as_lists(ms[i], cs[i]) == want, i
        own = sum(1 for ch in rec["board"] if ch != " " and ch.isupper() == (rec["turn"] == 0))
        crowded += own > 16
    assert crowded > 100
