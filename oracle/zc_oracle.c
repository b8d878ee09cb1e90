/*
 * zc_oracle.c -- TEST INFRASTRUCTURE.  CPU restatement (plain C11) of the reference's
 * MCTS self-play hot path.  It is the checker for the CUDA engine; it is never linked,
 * imported or executed by the product (only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it).
 *
 * Every function cites the reference file:line it restates (paths relative to the
 * reference checkout, /root/reference in the dev container).  The style is deliberately
 * different from the product: mailbox boards and pointer-linked heap nodes here,
 * bitboards and struct-of-arrays pools in the CUDA engine, so that the two are
 * independent derivations of the same semantics.
 *
 * Parity pinning: this file is checked (tests/test_oracle_pinned.py) against
 *   - the reference's own test vectors (tests/test_cb.py:39-116),
 *   - the reference's perft / known-answer tables recorded in SURVEY.md App. B/D,
 *   - golden vectors produced by running the UNMODIFIED reference in the dev container
 *     (tests/golden/make_golden.py -> tests/golden/ JSON files), and
 *   - oracle/_ref (the reference C++ compiled from its own sources) when it is present.
 *
 * Build: make -C oracle port      (gcc -O2 -ffp-contract=off; fma() is called explicitly
 *                                  where the reference build fuses, see uct()).
 */
#include <ctype.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define ZO_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------ */
/*  Connect Four -- engine/games/connect4/c4_backend.py                                  */
/* ------------------------------------------------------------------------------------ */

enum { C4_ROWS = 6, C4_COLS = 7 };

/* cells: ' ' empty, 'X' first player (turn 0), 'O' second (c4_backend.py:9); row 0 = top */
typedef struct {
    char cell[C4_ROWS][C4_COLS];
    int32_t turn;
} zo_c4_state;

/* CPython set-iteration order of {(c,0) for c in legal columns} (c4_backend.py:49-50),
 * indexed by the bit mask of playable columns.  Default = CPython 3.12 (the table in
 * SURVEY.md App. C); tests regenerate it from the running interpreter and install it with
 * zo_c4_set_order().  255 terminates a row. */
static uint8_t c4_order[128][8] = {
#include "c4_order_py312.inc"
};

ZO_API void zo_c4_set_order(const uint8_t *table /* [128][8] */) { memcpy(c4_order, table, sizeof c4_order); }
ZO_API void zo_c4_get_order(uint8_t *table) { memcpy(table, c4_order, sizeof c4_order); }

/* c4_backend.py:11-12 */
ZO_API void zo_c4_init(zo_c4_state *s) {
    memset(s->cell, ' ', sizeof s->cell);
    s->turn = 0;
}

/* c4_backend.py:14-23 -- the disc lands in the lowest empty cell of the column; a full
 * column leaves the board unchanged but still flips the turn. */
ZO_API void zo_c4_play(const zo_c4_state *s, int col, zo_c4_state *out) {
    zo_c4_state n = *s;
    for (int r = C4_ROWS - 1; r >= 0; --r) {
        if (n.cell[r][col] == ' ') {
            n.cell[r][col] = s->turn == 0 ? 'X' : 'O';
            break;
        }
    }
    n.turn = 1 - s->turn;
    *out = n;
}

/* c4_backend.py:25-44 -- four in a row for the player who has just moved (1 - turn) */
ZO_API int zo_c4_check_win(const zo_c4_state *s) {
    const char t = (1 - s->turn) == 0 ? 'X' : 'O';
    static const int dr[4] = {0, 1, 1, -1}, dc[4] = {1, 0, 1, 1};
    for (int d = 0; d < 4; ++d)
        for (int r = 0; r < C4_ROWS; ++r)
            for (int c = 0; c < C4_COLS; ++c) {
                int r3 = r + 3 * dr[d], c3 = c + 3 * dc[d];
                if (r3 < 0 || r3 >= C4_ROWS || c3 >= C4_COLS) continue;
                int k = 0;
                while (k < 4 && s->cell[r + k * dr[d]][c + k * dc[d]] == t) ++k;
                if (k == 4) return 1;
            }
    return 0;
}

/* c4_backend.py:46-47 */
ZO_API int zo_c4_check_draw(const zo_c4_state *s) {
    for (int r = 0; r < C4_ROWS; ++r)
        for (int c = 0; c < C4_COLS; ++c)
            if (s->cell[r][c] == ' ') return 0;
    return 1;
}

/* c4_backend.py:49-50 -- columns whose TOP cell is empty, in CPython set order.
 * Wins are ignored: the tree keeps growing below a won position. */
ZO_API int zo_c4_legal(const zo_c4_state *s, int32_t *cols) {
    int mask = 0;
    for (int c = 0; c < C4_COLS; ++c)
        if (s->cell[0][c] == ' ') mask |= 1 << c;
    int n = 0;
    while (n < 7 && c4_order[mask][n] != 255) {
        cols[n] = c4_order[mask][n];
        ++n;
    }
    return n;
}

/* c4_backend.py:52-61 -- float32[2][6][7]: plane 0 = side to move's discs, plane 1 = opponent's */
ZO_API void zo_c4_to_tensor(const zo_c4_state *s, float *out) {
    const char cur = s->turn == 0 ? 'X' : 'O', opp = s->turn == 0 ? 'O' : 'X';
    for (int r = 0; r < C4_ROWS; ++r)
        for (int c = 0; c < C4_COLS; ++c) {
            out[r * C4_COLS + c] = s->cell[r][c] == cur ? 1.0f : 0.0f;
            out[42 + r * C4_COLS + c] = s->cell[r][c] == opp ? 1.0f : 0.0f;
        }
}

/* ------------------------------------------------------------------------------------ */
/*  Chess -- engine/games/chess/src/chess_backend.cpp, include/state.h                   */
/* ------------------------------------------------------------------------------------ */

/* state.h:9-23 without the two move-history deques (only check_draw reads them; they are
 * passed separately to zo_ch_check_draw). board[r*8+c], row 0 = rank 8, ASCII pieces. */
typedef struct {
    uint8_t board[64];
    uint8_t turn, fifty, w_ck, w_cq, b_ck, b_cq;
    uint8_t pad[2];
} zo_ch_state;

typedef struct {
    uint8_t fr, fc, tr, tc;
    float val; /* captured piece value, exactly representable (0,1,3,5,9) */
} zo_ch_move;

/* chess_backend.cpp:17-34 */
static const int8_t KNIGHT_D[8][2] = {{-2, -1}, {-2, 1}, {-1, -2}, {-1, 2}, {1, -2}, {1, 2}, {2, -1}, {2, 1}};
static const int8_t DIAG_D[4][2] = {{-1, -1}, {-1, 1}, {1, -1}, {1, 1}};
static const int8_t ORTH_D[4][2] = {{-1, 0}, {1, 0}, {0, -1}, {0, 1}};
static const int8_t ROYAL_D[8][2] = {{-1, -1}, {-1, 1}, {1, -1}, {1, 1}, {-1, 0}, {1, 0}, {0, -1}, {0, 1}};

static int on_board(int r, int c) { return r >= 0 && r < 8 && c >= 0 && c < 8; }
static int sq_empty(uint8_t p) { return p == ' ' || p == 0; }                  /* :40-43 */
static int sq_enemy(uint8_t p, int turn) {                                     /* :44-49 */
    if (sq_empty(p)) return 0;
    int white = isupper(p) != 0;
    return turn == 0 ? !white : white;
}
static int capture_value(uint8_t p) {                                          /* :50-64 */
    switch (toupper(p)) {
    case 'P': return 1;
    case 'N': case 'B': return 3;
    case 'R': return 5;
    case 'Q': return 9;
    case 'K': return 100;
    default: return 0;
    }
}

/* chess_backend.cpp:68-81 -- first square in index order holding that side's king */
static void locate_king(const zo_ch_state *s, int side, int *kr, int *kc) {
    const uint8_t k = side == 0 ? 'K' : 'k';
    for (int i = 0; i < 64; ++i)
        if (s->board[i] == k) { *kr = i / 8; *kc = i % 8; return; }
    *kr = *kc = -1;
}

/* chess_backend.cpp:85-144 -- is the king of side s->turn, standing on (kr,kc), attacked? */
static int king_in_check(const zo_ch_state *s, int kr, int kc) {
    const int t = s->turn;
    const int pr = t == 0 ? kr - 1 : kr + 1;
    const uint8_t pawn = t == 0 ? 'p' : 'P';
    for (int dc = -1; dc <= 1; dc += 2)
        if (on_board(pr, kc + dc) && s->board[pr * 8 + kc + dc] == pawn) return 1;
    const uint8_t kn = t ? 'N' : 'n';
    for (int i = 0; i < 8; ++i) {
        int r = kr + KNIGHT_D[i][0], c = kc + KNIGHT_D[i][1];
        if (on_board(r, c) && s->board[r * 8 + c] == kn) return 1;
    }
    const uint8_t rook = t ? 'R' : 'r', bishop = t ? 'B' : 'b', queen = t ? 'Q' : 'q';
    for (int pass = 0; pass < 2; ++pass) {
        const int8_t(*dirs)[2] = pass == 0 ? ORTH_D : DIAG_D;
        const uint8_t slider = pass == 0 ? rook : bishop;
        for (int i = 0; i < 4; ++i) {
            int r = kr + dirs[i][0], c = kc + dirs[i][1];
            while (on_board(r, c)) {
                uint8_t p = s->board[r * 8 + c];
                if (!sq_empty(p)) {
                    if (p == slider || p == queen) return 1;
                    break;
                }
                r += dirs[i][0];
                c += dirs[i][1];
            }
        }
    }
    const uint8_t ek = t ? 'K' : 'k';
    for (int i = 0; i < 8; ++i) {
        int r = kr + ROYAL_D[i][0], c = kc + ROYAL_D[i][1];
        if (on_board(r, c) && s->board[r * 8 + c] == ek) return 1;
    }
    return 0;
}

/* chess_backend.cpp:364-400 (board/flag part; histories are the caller's business) */
ZO_API void zo_ch_play(const zo_ch_state *s, const zo_ch_move *m, zo_ch_state *out) {
    zo_ch_state n = *s;
    n.turn = (uint8_t)(1 - s->turn);
    n.fifty = (uint8_t)(s->fifty + 1);
    const int from = m->fr * 8 + m->fc, to = m->tr * 8 + m->tc;
    const uint8_t pc = s->board[from], target = s->board[to];
    if (pc == 'P' || pc == 'p' || !sq_empty(target)) n.fifty = 0;
    if (pc == 'K' || (pc == 'R' && m->fc == 7)) n.w_ck = 0;
    if (pc == 'K' || (pc == 'R' && m->fc == 0)) n.w_cq = 0;
    if (pc == 'k' || (pc == 'r' && m->fc == 7)) n.b_ck = 0;
    if (pc == 'k' || (pc == 'r' && m->fc == 0)) n.b_cq = 0;
    const int df = (int)m->tc - (int)m->fc;
    if (pc == 'K' && df == 2)  { n.board[61] = 'R'; n.board[63] = ' '; }
    if (pc == 'k' && df == 2)  { n.board[5]  = 'r'; n.board[7]  = ' '; }
    if (pc == 'K' && df == -2) { n.board[59] = 'R'; n.board[56] = ' '; }
    if (pc == 'k' && df == -2) { n.board[3]  = 'r'; n.board[0]  = ' '; }
    n.board[to] = pc;
    n.board[from] = ' ';
    if (m->tr == 0 && pc == 'P') n.board[to] = 'Q';
    if (m->tr == 7 && pc == 'p') n.board[to] = 'q';
    *out = n;
}

static int push_move(zo_ch_move *mv, int n, int fr, int fc, int tr, int tc, int val) {
    mv[n].fr = (uint8_t)fr; mv[n].fc = (uint8_t)fc; mv[n].tr = (uint8_t)tr; mv[n].tc = (uint8_t)tc;
    mv[n].val = (float)val;
    return n + 1;
}

/* a non-pawn step/leap target is usable if empty, or an enemy piece other than the king
 * (chess_backend.cpp:261-266, 306-317, 331-337) */
static int step_target(const zo_ch_state *s, int t, int r, int c, int *val) {
    uint8_t p = s->board[r * 8 + c];
    if (sq_empty(p)) { *val = 0; return 1; }
    if (sq_enemy(p, t) && toupper(p) != 'K') { *val = capture_value(p); return 1; }
    return 0;
}

/* chess_backend.cpp:184-360.  moves[] must hold 256 entries.  Returns the count. */
ZO_API int zo_ch_legal(const zo_ch_state *s, zo_ch_move *legal) {
    const int t = s->turn;
    int heavy = 0, minor = 0;                                          /* :188-198 */
    for (int i = 0; i < 64; ++i) {
        int u = toupper(s->board[i]);
        if (u == 'P' || u == 'R' || u == 'Q') ++heavy;
        if (u == 'B' || u == 'N') ++minor;
    }
    if (heavy == 0 && minor <= 1) return 0;

    zo_ch_move pseudo[512];
    int np = 0;
    for (int i = 0; i < 64; ++i) {                                     /* :203-342 */
        const uint8_t pc = s->board[i];
        if (sq_empty(pc)) continue;
        const int white = isupper(pc) != 0;
        if ((t == 0) != white) continue;
        const int r = i / 8, c = i % 8, u = toupper(pc);
        int v;
        if (pc == 'P' || pc == 'p') {
            const int d = pc == 'P' ? -1 : 1, home = pc == 'P' ? 6 : 1;
            if (on_board(r + d, c) && sq_empty(s->board[(r + d) * 8 + c])) {
                np = push_move(pseudo, np, r, c, r + d, c, 0);
                if (r == home && on_board(r + 2 * d, c) && sq_empty(s->board[(r + 2 * d) * 8 + c]))
                    np = push_move(pseudo, np, r, c, r + 2 * d, c, 0);
            }
            for (int dc = -1; dc <= 1; dc += 2) {
                if (!on_board(r + d, c + dc)) continue;
                uint8_t o = s->board[(r + d) * 8 + c + dc];
                if (sq_enemy(o, t) && toupper(o) != 'K')
                    np = push_move(pseudo, np, r, c, r + d, c + dc, capture_value(o));
            }
        } else if (u == 'N') {
            for (int k = 0; k < 8; ++k) {
                int rr = r + KNIGHT_D[k][0], cc = c + KNIGHT_D[k][1];
                if (on_board(rr, cc) && step_target(s, t, rr, cc, &v)) np = push_move(pseudo, np, r, c, rr, cc, v);
            }
        } else if (u == 'B' || u == 'R' || u == 'Q') {
            const int8_t(*dirs)[2] = u == 'B' ? DIAG_D : u == 'R' ? ORTH_D : ROYAL_D;
            const int nd = u == 'Q' ? 8 : 4;
            for (int k = 0; k < nd; ++k) {
                int rr = r + dirs[k][0], cc = c + dirs[k][1];
                while (on_board(rr, cc)) {
                    uint8_t o = s->board[rr * 8 + cc];
                    if (sq_empty(o)) {
                        np = push_move(pseudo, np, r, c, rr, cc, 0);
                    } else {
                        if (sq_enemy(o, t) && toupper(o) != 'K') np = push_move(pseudo, np, r, c, rr, cc, capture_value(o));
                        break;
                    }
                    rr += dirs[k][0];
                    cc += dirs[k][1];
                }
            }
        } else if (u == 'K') {
            for (int k = 0; k < 8; ++k) {
                int rr = r + ROYAL_D[k][0], cc = c + ROYAL_D[k][1];
                if (on_board(rr, cc) && step_target(s, t, rr, cc, &v)) np = push_move(pseudo, np, r, c, rr, cc, v);
            }
        }
    }
    int nl = 0;                                                        /* :345-358 */
    for (int k = 0; k < np; ++k) {
        zo_ch_state after;
        zo_ch_play(s, &pseudo[k], &after);
        int kr, kc;
        locate_king(&after, t, &kr, &kc);
        after.turn = (uint8_t)t;
        if (!king_in_check(&after, kr, kc)) legal[nl++] = pseudo[k];
    }
    return nl;
}

/* chess_backend.cpp:404-412 */
ZO_API int zo_ch_check_win(const zo_ch_state *s) {
    zo_ch_move mv[256];
    if (zo_ch_legal(s, mv) > 0) return 0;
    int kr, kc;
    locate_king(s, s->turn, &kr, &kc);
    return king_in_check(s, kr, kc);
}

/* chess_backend.cpp:148-180 -- does some prefix of L consist of >= min_rep repeats of a
 * period >= min_len?  L is most-recent-first; entries compare by all five fields. */
static int repeated_prefix(const zo_ch_move *L, int n, int min_len, int min_rep) {
    if (n < min_len * min_rep) return 0;
    int *pi = (int *)calloc((size_t)n, sizeof(int));
    int j = 0, found = 0;
#define MV_EQ(a, b) (L[a].fr == L[b].fr && L[a].fc == L[b].fc && L[a].tr == L[b].tr && L[a].tc == L[b].tc && L[a].val == L[b].val)
    for (int i = 1; i < n; ++i) {
        while (j > 0 && !MV_EQ(i, j)) j = pi[j - 1];
        if (MV_EQ(i, j)) ++j;
        pi[i] = j;
    }
#undef MV_EQ
    for (int i = 0; i < n && !found; ++i) {
        int len = i + 1, p = len - pi[i];
        if (p >= min_len && len % p == 0 && len / p >= min_rep) found = 1;
    }
    free(pi);
    return found;
}

/* chess_backend.cpp:416-441.  hist_* are most-recent-first (push_front at :374). */
ZO_API int zo_ch_check_draw(const zo_ch_state *s, const zo_ch_move *hist_white, int nw,
                            const zo_ch_move *hist_black, int nb) {
    zo_ch_move mv[256];
    if (zo_ch_legal(s, mv) == 0) {
        int kr, kc;
        locate_king(s, s->turn, &kr, &kc);
        if (!king_in_check(s, kr, kc)) return 1;
    }
    if (s->fifty >= 50) return 1;
    if (repeated_prefix(hist_white, nw, 2, 3) && repeated_prefix(hist_black, nb, 2, 3)) return 1;
    return 0;
}

/* chess_backend.cpp:445-457 */
ZO_API void zo_ch_init(zo_ch_state *s) {
    memset(s, 0, sizeof *s);
    memcpy(s->board, "rnbqkbnrpppppppp", 16);
    memset(s->board + 16, ' ', 32);
    memcpy(s->board + 48, "PPPPPPPPRNBQKBNR", 16);
    s->w_ck = s->w_cq = s->b_ck = s->b_cq = 1;
}

/* chess_backend.cpp:525-556 -- en-passant and full-move fields ignored; a missing or
 * malformed half-move clock reads as 0 (operator>> failure semantics). */
ZO_API void zo_ch_from_fen(const char *fen, zo_ch_state *s) {
    memset(s, 0, sizeof *s);
    memset(s->board, ' ', 64);
    char pp[128] = "", ac[16] = "", cs[16] = "", ep[16] = "";
    int hm = 0;
    int got = sscanf(fen, "%127s %15s %15s %15s %d", pp, ac, cs, ep, &hm);
    if (got < 5) hm = 0;
    int idx = 0;
    for (const char *p = pp; *p && idx < 64; ++p) {
        if (*p == '/') continue;
        if (isdigit((unsigned char)*p)) idx += *p - '0';
        else s->board[idx++] = (uint8_t)*p;
    }
    s->turn = strcmp(ac, "w") == 0 ? 0 : 1;
    s->w_ck = strchr(cs, 'K') != NULL;
    s->w_cq = strchr(cs, 'Q') != NULL;
    s->b_ck = strchr(cs, 'k') != NULL;
    s->b_cq = strchr(cs, 'q') != NULL;
    s->fifty = (uint8_t)hm;
}

/* chess_backend.cpp:461-521 -- float32[17][8][8] */
ZO_API void zo_ch_to_tensor(const zo_ch_state *s, float *out) {
    static const char order[12] = {'P', 'N', 'B', 'R', 'Q', 'K', 'p', 'n', 'b', 'r', 'q', 'k'};
    memset(out, 0, 17 * 64 * sizeof(float));
    for (int i = 0; i < 64; ++i)
        for (int k = 0; k < 12; ++k)
            if (s->board[i] == (uint8_t)order[k]) { out[k * 64 + i] = 1.0f; break; }
    const float fill[5] = {s->turn == 0 ? 1.0f : 0.0f, s->w_ck ? 1.0f : 0.0f, s->w_cq ? 1.0f : 0.0f,
                           s->b_ck ? 1.0f : 0.0f, s->b_cq ? 1.0f : 0.0f};
    for (int k = 0; k < 5; ++k)
        for (int i = 0; i < 64; ++i) out[(12 + k) * 64 + i] = fill[k];
}

/* perft as SURVEY App. B counts it: number of legal moves at the last ply */
ZO_API uint64_t zo_ch_perft(const zo_ch_state *s, int depth) {
    zo_ch_move mv[256];
    int n = zo_ch_legal(s, mv);
    if (depth <= 1) return (uint64_t)n;
    uint64_t total = 0;
    for (int i = 0; i < n; ++i) {
        zo_ch_state c;
        zo_ch_play(s, &mv[i], &c);
        total += zo_ch_perft(&c, depth - 1);
    }
    return total;
}

/* ------------------------------------------------------------------------------------ */
/*  Evaluators -- engine/value_functions.py                                              */
/* ------------------------------------------------------------------------------------ */

enum {
    ZO_GAME_C4 = 0,
    ZO_GAME_CHESS = 1,
};
enum {
    ZO_EVAL_C4_TERMINAL = 0,   /* -1 if check_win(state) else 0: terminal branch of random_rollout, value_functions.py:41-43 */
    ZO_EVAL_C4_POSITIONAL = 1, /* -1 on a win, else sum over discs of w[col]*(+1 side to move, -1 opponent) / 64, w = 1,2,3,4,3,2,1 */
    ZO_EVAL_CHESS_CRUDE = 2,   /* crude_chess_score, value_functions.py:49-55 */
    ZO_EVAL_EXTERNAL = 3,      /* caller-supplied batch callback (the neural evaluator in tests) */
};
enum {
    ZO_POLICY_FIRST = 0, /* policy(moves) = moves[0]   */
    ZO_POLICY_LAST = 1,  /* policy(moves) = moves[-1]  */
};

static double eval_c4_terminal(const zo_c4_state *s) { return zo_c4_check_win(s) ? -1.0 : 0.0; }

static double eval_c4_positional(const zo_c4_state *s) {
    static const int w[7] = {1, 2, 3, 4, 3, 2, 1};
    if (zo_c4_check_win(s)) return -1.0;
    const char cur = s->turn == 0 ? 'X' : 'O';
    int acc = 0;
    for (int r = 0; r < C4_ROWS; ++r)
        for (int c = 0; c < C4_COLS; ++c)
            if (s->cell[r][c] != ' ') acc += s->cell[r][c] == cur ? w[c] : -w[c];
    return (double)acc / 64.0;
}

/* value_functions.py:49-55 -- note the sign quirk: a MATED side to move scores +1000 */
static double eval_chess_crude(const zo_ch_state *s) {
    if (zo_ch_check_win(s)) return 1000.0;
    int sum = 0;
    for (int i = 0; i < 64; ++i) {
        switch (s->board[i]) {
        case 'P': sum += 1; break;  case 'p': sum -= 1; break;
        case 'N': case 'B': sum += 3; break;  case 'n': case 'b': sum -= 3; break;
        case 'R': sum += 5; break;  case 'r': sum -= 5; break;
        case 'Q': sum += 9; break;  case 'q': sum -= 9; break;
        default: break;
        }
    }
    const int factor = s->turn * -2 + 1;
    return (double)(factor * sum);
}

ZO_API double zo_eval_state(int evaluator, const void *state) {
    switch (evaluator) {
    case ZO_EVAL_C4_TERMINAL: return eval_c4_terminal((const zo_c4_state *)state);
    case ZO_EVAL_C4_POSITIONAL: return eval_c4_positional((const zo_c4_state *)state);
    case ZO_EVAL_CHESS_CRUDE: return eval_chess_crude((const zo_ch_state *)state);
    default: return NAN;
    }
}

/* ------------------------------------------------------------------------------------ */
/*  Search -- engine/mcts/src/mcts.cpp                                                   */
/* ------------------------------------------------------------------------------------ */

#define ZO_STATE_BYTES 72 /* >= sizeof(zo_ch_state), sizeof(zo_c4_state) */

/* mcts.cpp:10-39 */
typedef struct zo_node {
    uint8_t state[ZO_STATE_BYTES];
    int n_moves;
    zo_ch_move *ch_moves; /* chess move list (backend order) */
    int32_t c4_moves[7];  /* C4 columns (backend order) */
    int *Na;
    double *Wa, *Qa;
    struct zo_node **child;
    int *untried;
    int n_untried;
    struct zo_node *parent;
    int parent_action;
    int N;
    int depth;
} zo_node;

typedef void (*zo_batch_eval_fn)(const uint8_t *states, int n, int state_bytes, double *out, void *user);

typedef struct {
    int game, evaluator, policy;
    zo_batch_eval_fn ext;
    void *ext_user;
    int64_t nodes_created, sum_leaf_depth, max_leaf_depth, reevaluated_leaves;
} zo_ctx;

static zo_node *node_new(zo_ctx *cx, const void *state, zo_node *parent, int action) {
    zo_node *n = (zo_node *)calloc(1, sizeof *n);
    if (cx->game == ZO_GAME_C4) {
        memcpy(n->state, state, sizeof(zo_c4_state));
        n->n_moves = zo_c4_legal((const zo_c4_state *)n->state, n->c4_moves);
    } else {
        zo_ch_move tmp[256];
        memcpy(n->state, state, sizeof(zo_ch_state));
        n->n_moves = zo_ch_legal((const zo_ch_state *)n->state, tmp);
        n->ch_moves = (zo_ch_move *)malloc(sizeof(zo_ch_move) * (size_t)(n->n_moves + 1));
        memcpy(n->ch_moves, tmp, sizeof(zo_ch_move) * (size_t)n->n_moves);
    }
    const size_t k = (size_t)n->n_moves + 1;
    n->Na = (int *)calloc(k, sizeof(int));
    n->Wa = (double *)calloc(k, sizeof(double));
    n->Qa = (double *)calloc(k, sizeof(double));
    n->child = (zo_node **)calloc(k, sizeof(zo_node *));
    n->untried = (int *)malloc(k * sizeof(int));
    for (int i = 0; i < n->n_moves; ++i) n->untried[i] = i;
    n->n_untried = n->n_moves;
    n->parent = parent;
    n->parent_action = action;
    n->depth = parent ? parent->depth + 1 : 0;
    cx->nodes_created++;
    return n;
}

static void node_free(zo_node *n) {
    if (!n) return;
    for (int i = 0; i < n->n_moves; ++i) node_free(n->child[i]);
    free(n->ch_moves); free(n->Na); free(n->Wa); free(n->Qa); free(n->child); free(n->untried);
    free(n);
}

/* mcts.cpp:41-45.  The reference build (-O3 -ffast-math on an FMA target) contracts
 * Qa + c*sqrt(log(N)/Na) into  fma(sqrt(log(N)/Na), c, Qa)  -- see `objdump -d` of
 * oracle/_ref/mcts*.so: call log; vdivsd; vsqrtsd; vfmadd213sd.  Written out explicitly. */
static double uct(const zo_node *n, int a, double c) {
    if (n->Na[a] == 0) return INFINITY;
    return fma(sqrt(log((double)n->N) / (double)n->Na[a]), c, n->Qa[a]);
}

/* mcts.cpp:47-63 */
static zo_node *select_leafward(zo_node *n, double c) {
    for (;;) {
        if (n->n_untried > 0) return n;
        int best = -1;
        double best_v = -1e100;
        for (int i = 0; i < n->n_moves; ++i) {
            if (!n->child[i]) continue;
            double v = uct(n, i, c);
            if (v > best_v) { best_v = v; best = i; }
        }
        if (best < 0) return n;
        n = n->child[best];
    }
}

/* mcts.cpp:65-78 with policy in {first, last} of the untried list */
static zo_node *expand(zo_ctx *cx, zo_node *n) {
    const int pick = cx->policy == ZO_POLICY_LAST ? n->n_untried - 1 : 0;
    const int mi = n->untried[pick];
    memmove(n->untried + pick, n->untried + pick + 1, sizeof(int) * (size_t)(n->n_untried - pick - 1));
    n->n_untried--;
    uint8_t next[ZO_STATE_BYTES] = {0};
    if (cx->game == ZO_GAME_C4) zo_c4_play((const zo_c4_state *)n->state, n->c4_moves[mi], (zo_c4_state *)next);
    else zo_ch_play((const zo_ch_state *)n->state, &n->ch_moves[mi], (zo_ch_state *)next);
    zo_node *ch = node_new(cx, next, n, mi);
    n->child[mi] = ch;
    return ch;
}

/* mcts.cpp:80-100 */
static void backprop(zo_node *n, double result) {
    while (n) {
        n->N += 1;
        zo_node *p = n->parent;
        if (!p) break;
        const int a = n->parent_action;
        p->Na[a] += 1;
        p->Wa[a] -= result;
        p->Qa[a] = p->Wa[a] / (double)p->Na[a];
        n = p;
        result = -result;
    }
}

typedef struct {
    int32_t n_moves;         /* root move count                                  */
    int32_t best;            /* index into the root move list (mcts.cpp:150-155) */
    int32_t Na[256];         /* per-child visit counts                           */
    double Wa[256];          /* per-child value sums                             */
    uint8_t moves[256][4];   /* chess: fr,fc,tr,tc   C4: col,0,0,0               */
    float move_val[256];
    int64_t nodes_created, sum_leaf_depth, max_leaf_depth, reevaluated_leaves;
    uint64_t tree_hash;      /* order-dependent hash over the whole tree, see hash_tree() */
    int32_t root_N;
} zo_search_result;

static uint64_t mix64(uint64_t h, uint64_t v) {
    h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    h *= 0xFF51AFD7ED558CCDull;
    h ^= h >> 33;
    return h;
}

/* Depth-first, children in move order.  Each node contributes (depth, #moves, #untried, N)
 * and each edge (index, Na, bits of Wa, child present).  The CUDA engine computes the same
 * hash over its pools (zc_tree_hash), so equality means the WHOLE tree matches bit for bit. */
static uint64_t hash_tree(const zo_node *n, uint64_t h) {
    h = mix64(h, ((uint64_t)(uint32_t)n->depth << 48) ^ ((uint64_t)(uint32_t)n->n_moves << 32) ^ ((uint64_t)(uint32_t)n->n_untried << 20) ^ (uint64_t)(uint32_t)n->N);
    for (int i = 0; i < n->n_moves; ++i) {
        uint64_t wb;
        double w = n->Wa[i] == 0.0 ? 0.0 : n->Wa[i]; /* fold -0.0 */
        memcpy(&wb, &w, 8);
        h = mix64(h, ((uint64_t)(uint32_t)i << 40) ^ ((uint64_t)(uint32_t)n->Na[i] << 1) ^ (n->child[i] ? 1u : 0u));
        h = mix64(h, wb);
        if (n->child[i]) h = hash_tree(n->child[i], h);
    }
    return h;
}

/* mcts.cpp:102-160.  `state` is a zo_c4_state or zo_ch_state according to `game`. */
ZO_API int zo_search(int game, const void *state, int simulations, double c, int batch_size, int evaluator,
                     int policy, zo_batch_eval_fn ext, void *ext_user, zo_search_result *res) {
    zo_ctx cx = {game, evaluator, policy, ext, ext_user, 0, 0, 0, 0};
    const int sbytes = game == ZO_GAME_C4 ? (int)sizeof(zo_c4_state) : (int)sizeof(zo_ch_state);
    if (batch_size < 1) batch_size = 1; /* a pending list of >= 1 flushes on every push, mcts.cpp:144 */
    zo_node *root = node_new(&cx, state, NULL, -1);
    zo_node **pending = (zo_node **)malloc(sizeof(zo_node *) * (size_t)batch_size);
    uint8_t *pstates = (uint8_t *)malloc((size_t)batch_size * (size_t)sbytes);
    double *vals = (double *)malloc(sizeof(double) * (size_t)batch_size);
    int np = 0;
    for (int i = 0; i <= simulations; ++i) {
        if (i < simulations) {
            zo_node *n = select_leafward(root, c);
            zo_node *leaf;
            if (n->n_untried > 0) leaf = expand(&cx, n);
            else { leaf = n; cx.reevaluated_leaves++; }
            cx.sum_leaf_depth += leaf->depth;
            if (leaf->depth > cx.max_leaf_depth) cx.max_leaf_depth = leaf->depth;
            pending[np++] = leaf;
        }
        if (np >= batch_size || (i == simulations && np > 0)) {         /* flush, :112-127 */
            if (evaluator == ZO_EVAL_EXTERNAL) {
                for (int k = 0; k < np; ++k) memcpy(pstates + (size_t)k * (size_t)sbytes, pending[k]->state, (size_t)sbytes);
                ext(pstates, np, sbytes, vals, ext_user);
            } else {
                for (int k = 0; k < np; ++k) vals[k] = zo_eval_state(evaluator, pending[k]->state);
            }
            for (int k = 0; k < np; ++k) backprop(pending[k], vals[k]);
            np = 0;
        }
    }
    memset(res, 0, sizeof *res);
    res->n_moves = root->n_moves;
    int best = -1, best_n = -1;                                       /* :150-155 */
    for (int i = 0; i < root->n_moves; ++i) {
        res->Na[i] = root->Na[i];
        res->Wa[i] = root->Wa[i];
        if (game == ZO_GAME_C4) { res->moves[i][0] = (uint8_t)root->c4_moves[i]; }
        else {
            res->moves[i][0] = root->ch_moves[i].fr; res->moves[i][1] = root->ch_moves[i].fc;
            res->moves[i][2] = root->ch_moves[i].tr; res->moves[i][3] = root->ch_moves[i].tc;
            res->move_val[i] = root->ch_moves[i].val;
        }
        if (root->child[i] && root->child[i]->N > best_n) { best_n = root->child[i]->N; best = i; }
    }
    res->best = best;
    res->root_N = root->N;
    res->nodes_created = cx.nodes_created;
    res->sum_leaf_depth = cx.sum_leaf_depth;
    res->max_leaf_depth = cx.max_leaf_depth;
    res->reevaluated_leaves = cx.reevaluated_leaves;
    res->tree_hash = hash_tree(root, 0x5A17C10E5EEDull);
    node_free(root);
    free(pending); free(pstates); free(vals);
    return 0;
}

/* ------------------------------------------------------------------------------------ */
/*  PUCT with stored priors and virtual loss -- NOT in the reference (mcts.cpp:41-63 is   */
/*  UCB1).  This is the checker of the library's opt-in ZC_SELECT_PUCT mode               */
/*  (zeroclone_b200/csrc/puct.cuh, include/zc_b200.h): the same definition written         */
/*  sequentially over heap nodes.  Parity is between this function and the CUDA kernels    */
/*  only -- there is no reference behaviour to pin it on ("parity unpinned" for this mode).*/
/* ------------------------------------------------------------------------------------ */

static int move_value_of(const zo_ctx *cx, const zo_node *n, int a) {
    return cx->game == ZO_GAME_C4 ? 0 : (int)n->ch_moves[a].val;   /* chess_backend.cpp:50-64; c4_backend.py:50 */
}

/* a* = argmax_a  Wa/Na + c * P(a) * sqrt(N + 1) / (1 + Na), lowest index on ties; every operation rounded on its own */
static int puct_pick(const zo_ctx *cx, const zo_node *n, double c, int prior_weight, const float *root_priors) {
    int wsum = 0;
    for (int a = 0; a < n->n_moves; ++a) wsum += 1 + prior_weight * move_value_of(cx, n, a);
    const double sqrtN = sqrt((double)(n->N + 1));
    int best = -1;
    double best_v = -INFINITY;
    for (int a = 0; a < n->n_moves; ++a) {
        const float P = (root_priors && !n->parent) ? root_priors[a] : (float)(1 + prior_weight * move_value_of(cx, n, a)) / (float)wsum;
        const double q = n->Na[a] ? n->Wa[a] / (double)n->Na[a] : 0.0;
        double u = c * (double)P;
        u = u * sqrtN;
        u = u / (double)(1 + n->Na[a]);
        const double v = q + u;
        if (v > best_v) { best_v = v; best = a; }
    }
    return best;
}

ZO_API int zo_search_puct(int game, const void *state, int simulations, double c, int batch_size, int evaluator,
                          double vloss, int prior_weight, const float *root_priors, zo_batch_eval_fn ext, void *ext_user,
                          zo_search_result *res) {
    zo_ctx cx = {game, evaluator, ZO_POLICY_FIRST, ext, ext_user, 0, 0, 0, 0};
    const int sbytes = game == ZO_GAME_C4 ? (int)sizeof(zo_c4_state) : (int)sizeof(zo_ch_state);
    if (batch_size < 1) batch_size = 1;
    zo_node *root = node_new(&cx, state, NULL, -1);
    zo_node **pending = (zo_node **)malloc(sizeof(zo_node *) * (size_t)batch_size);
    uint8_t *pstates = (uint8_t *)malloc((size_t)batch_size * (size_t)sbytes);
    double *vals = (double *)malloc(sizeof(double) * (size_t)batch_size);
    int np = 0;
    for (int i = 0; i <= simulations; ++i) {
        if (i < simulations) {
            zo_node *n = root;
            zo_node *leaf = NULL;
            for (;;) {
                if (n->n_moves == 0) { leaf = n; cx.reevaluated_leaves++; break; }      /* its own leaf, again */
                const int a = puct_pick(&cx, n, c, prior_weight, root_priors);
                n->N += 1;                                                            /* virtual loss */
                n->Na[a] += 1;
                n->Wa[a] -= vloss;
                if (!n->child[a]) {
                    uint8_t next[ZO_STATE_BYTES] = {0};
                    if (game == ZO_GAME_C4) zo_c4_play((const zo_c4_state *)n->state, n->c4_moves[a], (zo_c4_state *)next);
                    else zo_ch_play((const zo_ch_state *)n->state, &n->ch_moves[a], (zo_ch_state *)next);
                    n->child[a] = node_new(&cx, next, n, a);
                    n->n_untried--;
                    leaf = n->child[a];
                    break;
                }
                n = n->child[a];
            }
            leaf->N += 1;
            cx.sum_leaf_depth += leaf->depth;
            if (leaf->depth > cx.max_leaf_depth) cx.max_leaf_depth = leaf->depth;
            pending[np++] = leaf;
        }
        if (np >= batch_size || (i == simulations && np > 0)) {
            if (evaluator == ZO_EVAL_EXTERNAL) {
                for (int k = 0; k < np; ++k) memcpy(pstates + (size_t)k * (size_t)sbytes, pending[k]->state, (size_t)sbytes);
                ext(pstates, np, sbytes, vals, ext_user);
            } else {
                for (int k = 0; k < np; ++k) vals[k] = zo_eval_state(evaluator, pending[k]->state);
            }
            for (int k = 0; k < np; ++k) {                                            /* back up in pending order */
                double result = vals[k];
                for (zo_node *n = pending[k]; n->parent; n = n->parent) {
                    zo_node *p = n->parent;
                    const int a = n->parent_action;
                    p->Wa[a] += vloss;
                    p->Wa[a] -= result;
                    result = -result;
                }
            }
            np = 0;
        }
    }
    memset(res, 0, sizeof *res);
    res->n_moves = root->n_moves;
    int best = -1, best_n = -1;
    for (int i = 0; i < root->n_moves; ++i) {
        res->Na[i] = root->Na[i];
        res->Wa[i] = root->Wa[i];
        if (game == ZO_GAME_C4) { res->moves[i][0] = (uint8_t)root->c4_moves[i]; }
        else {
            res->moves[i][0] = root->ch_moves[i].fr; res->moves[i][1] = root->ch_moves[i].fc;
            res->moves[i][2] = root->ch_moves[i].tr; res->moves[i][3] = root->ch_moves[i].tc;
            res->move_val[i] = root->ch_moves[i].val;
        }
        if (root->child[i] && root->Na[i] > best_n) { best_n = root->Na[i]; best = i; }
    }
    res->best = best;
    res->root_N = root->N;
    res->nodes_created = cx.nodes_created;
    res->sum_leaf_depth = cx.sum_leaf_depth;
    res->max_leaf_depth = cx.max_leaf_depth;
    res->reevaluated_leaves = cx.reevaluated_leaves;
    res->tree_hash = hash_tree(root, 0x5A17C10E5EEDull);
    node_free(root);
    free(pending); free(pstates); free(vals);
    return 0;
}

ZO_API int zo_sizeof_c4_state(void) { return (int)sizeof(zo_c4_state); }
ZO_API int zo_sizeof_ch_state(void) { return (int)sizeof(zo_ch_state); }
ZO_API int zo_sizeof_search_result(void) { return (int)sizeof(zo_search_result); }
