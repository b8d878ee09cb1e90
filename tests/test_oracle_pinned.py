"""Pin the CPU oracle (oracle/zc_oracle.c) to the reference.

Sources of truth, in order: the reference's own test vectors (tests/test_cb.py in the
reference), golden vectors produced by running the unmodified reference
(tests/golden/make_golden.py), SURVEY.md App. B/D tables, and oracle/_ref when present.
"""
import os
import sys

import numpy as np
import pytest

from oracle import zc_oracle as zo
from conftest import load_golden, REPO

EVAL = {"c4_terminal": zo.EVAL_C4_TERMINAL, "c4_positional": zo.EVAL_C4_POSITIONAL, "chess_crude": zo.EVAL_CHESS_CRUDE}
POL = {"first": zo.POLICY_FIRST, "last": zo.POLICY_LAST}


def ch_state(rec):
    s = zo.ChState()
    for i, ch in enumerate(rec["board"]):
        s.board[i] = ord(ch)
    s.turn, s.fifty = rec["turn"], rec["fifty"]
    s.w_ck, s.w_cq, s.b_ck, s.b_cq = rec["flags"]
    return s


def mv_tuple(m):
    return ((m[0], m[1], m[2], m[3]), float(m[4]))


# ---------------------------------------------------------------- C4
def test_c4_order_table_matches_running_python_and_golden():
    g = load_golden("c4_order_py312.json")["order"]
    t = zo.python_c4_order()
    baked = zo.baked_c4_order()
    for mask in range(128):
        assert [int(x) for x in t[mask] if x != 255] == g[str(mask)]
    assert (baked == t).all(), "CPython set order changed: the baked 3.12 table no longer applies"


def test_c4_rules_against_reference_playouts():
    for trace in load_golden("c4_rules.json.gz"):
        s = zo.c4_init()
        for rec in trace:
            assert ["".join(r) for r in zo.c4_rows(s)] == rec["rows"]
            assert s.turn == rec["turn"]
            assert zo.c4_legal(s) == rec["legal"]
            assert zo.c4_check_win(s) == rec["win"]
            assert zo.c4_check_draw(s) == rec["draw"]
            assert zo.c4_to_tensor(s).astype(int).reshape(-1).tolist() == rec["tensor"]
            if "played" in rec:
                s = zo.c4_play(s, rec["played"])


def test_c4_search_matches_reference_bit_exact():
    cases = load_golden("c4_search.json")
    assert len(cases) > 40
    for cs in cases:
        root = zo.c4_from_rows(cs["rows"], cs["turn"])
        r = zo.search(zo.GAME_C4, root, cs["sims"], cs["c"], cs["batch"], EVAL[cs["evaluator"]], POL[cs["policy"]])
        tag = (cs["cols"], cs["evaluator"], cs["policy"], cs["sims"], cs["c"], cs["batch"])
        assert [m[0] for m in r.moves] == [m[0] for m in cs["moves"]], tag
        assert r.Na == cs["Na"], tag
        assert r.Wa == cs["Wa"], tag           # exact: dyadic evaluator values
        assert r.best == cs["best"], tag
        assert r.nodes_created == cs["nodes_created"], tag
        assert r.sum_leaf_depth == cs["sum_leaf_depth"] and r.max_leaf_depth == cs["max_leaf_depth"], tag


def test_c4_survey_known_answers():
    # SURVEY.md App. D rows 1-3
    r = zo.search(zo.GAME_C4, zo.c4_init(), 800, evaluator=zo.EVAL_C4_TERMINAL)
    assert [m[0] for m in r.moves] == [4, 0, 2, 3, 5, 6, 1]
    assert r.Na == [122, 129, 129, 129, 97, 97, 97] and r.moves[r.best] == (0, 0)
    r = zo.search(zo.GAME_C4, zo.c4_init(), 800, evaluator=zo.EVAL_C4_POSITIONAL)
    assert r.Na == [122, 97, 129, 129, 129, 97, 97]
    assert r.Wa == [3.109375, -1.15625, 3.203125, 5.21875, 1.1875, -1.15625, 0.359375]
    assert r.moves[r.best] == (2, 0)
    r = zo.search(zo.GAME_C4, zo.c4_from_moves([3, 3, 3, 3, 3, 3, 0, 1, 0, 1, 0, 1]), 800, evaluator=zo.EVAL_C4_TERMINAL)
    assert r.Na == [59, 385, 65, 65, 65, 161] and r.Wa == [-11, 63, -9, -7, -9, 11] and r.moves[r.best] == (0, 0)
    assert r.max_leaf_depth == 10


# ---------------------------------------------------------------- chess rules
def test_reference_test_cb_vectors():
    # reference tests/test_cb.py:39-51, 82-99, 105-116
    s = zo.ch_init()
    assert len(zo.ch_legal(s)) == 20
    assert not zo.ch_check_win(s) and not zo.ch_check_draw(s)
    assert zo.ch_play(s, zo.ch_legal(s)[0]).turn == 1
    assert zo.ch_to_tensor(s).shape == (17, 8, 8)
    for fen, win, draw in [
        ("rnb1kbnr/pppp1ppp/8/4p3/6Pq/5P2/PPPPP2P/RNBQKBNR w KQkq - 0 1", True, False),
        ("r1bqkbnr/ppp2Qpp/n2p4/4p3/2B1P3/8/PPPP1PPP/RNB1K1NR b KQkq - 0 1", True, False),
        ("7k/5Q2/6K1/8/8/8/8/8 b - - 0 1", False, True),
        ("8/8/8/8/8/8/2n5/2K4k w - - 0 1", False, True),
        ("8/8/8/1k6/8/8/4K3/5B2 w - - 0 1", False, True),
    ]:
        st = zo.ch_from_fen(fen)
        assert zo.ch_check_win(st) is win and zo.ch_check_draw(st) is draw


def test_reference_fifty_ply_counter():
    # reference tests/test_cb.py:53-79: counter reaches 50 after 50 knight-bounce plies
    s = zo.ch_init()
    wb, bb = {(7, 6, 5, 5), (5, 5, 7, 6)}, {(0, 6, 2, 5), (2, 5, 0, 6)}
    for ply in range(50):
        m = next(x for x in zo.ch_legal(s) if x[0] in (wb if s.turn == 0 else bb))
        s = zo.ch_play(s, m)
        assert (s.fifty >= 50) == (ply == 49)


def test_chess_fens_perft_and_playouts_against_reference():
    g = load_golden("chess_rules.json.gz")
    for name, rec in g["fens"].items():
        s = zo.ch_from_fen(rec["fen"])
        assert zo.ch_board_str(s) == rec["board"] and s.turn == rec["turn"] and s.fifty == rec["fifty"], name
        assert [s.w_ck, s.w_cq, s.b_ck, s.b_cq] == rec["flags"], name
        assert zo.ch_legal(s) == [mv_tuple(m) for m in rec["legal"]], name
        assert zo.ch_check_win(s) == rec["win"] and zo.ch_check_draw(s) == rec["draw"], name
        t = zo.ch_to_tensor(s).reshape(17, 64)
        assert [int(sum(1 << i for i in range(64) if t[p, i] == 1)) for p in range(17)] == rec["tensor_planes"], name
    for name, counts in g["perft"].items():
        s = zo.ch_from_fen(g["fens"][name]["fen"])
        assert [zo.ch_perft(s, d) for d in range(1, len(counts) + 1)] == counts, name
    n_draw = n_win = 0
    for trace in g["playouts"]:
        s = ch_state(trace[0])
        hist = [[], []]
        for rec in trace:
            assert zo.ch_board_str(s) == rec["board"] and s.turn == rec["turn"] and s.fifty == rec["fifty"]
            assert [s.w_ck, s.w_cq, s.b_ck, s.b_cq] == rec["flags"]
            assert zo.ch_legal(s) == [mv_tuple(m) for m in rec["legal"]]
            assert zo.ch_check_win(s) == rec["win"]
            assert zo.ch_check_draw(s, hist[0], hist[1]) == rec["draw"]
            n_draw += rec["draw"]
            n_win += rec["win"]
            if "played" in rec:
                m = mv_tuple(rec["played"])
                hist[s.turn].insert(0, m)
                s = zo.ch_play(s, m)
    assert n_draw > 10  # repetition / 50-ply / stalemate cases are present in the fixtures


def test_survey_perft_table():
    # SURVEY.md App. B (reference perft, differs from standard chess)
    assert [zo.ch_perft(zo.ch_init(), d) for d in (1, 2, 3, 4)] == [20, 400, 8902, 197281]
    kiwi = zo.ch_from_fen("r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq -")
    assert [zo.ch_perft(kiwi, d) for d in (1, 2, 3)] == [46, 1865, 86585]


@pytest.mark.slow
def test_survey_perft_depth5():
    assert zo.ch_perft(zo.ch_init(), 5) == 4865351


# ---------------------------------------------------------------- chess search
def test_chess_search_matches_reference_bit_exact():
    for cs in load_golden("chess_search.json"):
        root = ch_state(cs)
        r = zo.search(zo.GAME_CHESS, root, cs["sims"], cs["c"], cs["batch"], EVAL[cs["evaluator"]], POL[cs["policy"]])
        tag = (cs["name"], cs["policy"], cs["sims"], cs["batch"])
        assert r.moves == [mv_tuple(m) for m in cs["moves"]], tag
        assert r.Na == cs["Na"], tag
        assert r.Wa == cs["Wa"], tag
        assert r.best == cs["best"], tag
        assert r.nodes_created == cs["nodes_created"], tag
        assert r.sum_leaf_depth == cs["sum_leaf_depth"] and r.max_leaf_depth == cs["max_leaf_depth"], tag


# ---------------------------------------------------------------- oracle/_ref (when it travelled with the repo)
def _ref_modules():
    d = os.path.join(REPO, "oracle", "_ref")
    if not os.path.isdir(d) or not any(f.startswith("chess_backend") for f in os.listdir(d)):
        pytest.skip("oracle/_ref not built (make -C oracle ref needs /root/reference)")
    if d not in sys.path:
        sys.path.insert(0, d)
    import chess_backend
    import mcts
    return chess_backend, mcts


def test_oracle_vs_compiled_reference_random_positions():
    cb, _ = _ref_modules()
    rng = np.random.default_rng(5)
    for g in range(6):
        rs, s = cb.create_init_state(), zo.ch_init()
        for ply in range(120):
            rm = cb.get_legal_moves(rs)
            om = zo.ch_legal(s)
            assert [(tuple(m[0]), m[1]) for m in rm] == om
            assert cb.check_win(rs) == zo.ch_check_win(s)
            pv = {'P': 1, 'N': 3, 'B': 3, 'R': 5, 'Q': 9, 'p': -1, 'n': -3, 'b': -3, 'r': -5, 'q': -9}
            crude = 1000 if cb.check_win(rs) else (rs.turn * -2 + 1) * sum(pv.get(chr(p), 0) for p in rs.board)
            assert zo.eval_state(zo.EVAL_CHESS_CRUDE, s) == crude   # value_functions.py:49-55
            assert np.array_equal(cb.state_to_tensor(rs), zo.ch_to_tensor(s))
            if not rm:
                break
            k = int(rng.integers(len(rm)))
            rs, s = cb.play_move(rs, rm[k]), zo.ch_play(s, om[k])
            assert bytes(rs.board).decode() == zo.ch_board_str(s) and rs.fifty_move_rule_counter == s.fifty
