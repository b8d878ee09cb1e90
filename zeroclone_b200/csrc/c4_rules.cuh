// Connect Four rules on bitboards -- host+device.
// Semantics follow the reference's engine/games/connect4/c4_backend.py (cited per function);
// the representation does not: the reference keeps a 6x7 list of lists, this keeps two u64.
//
// Bit layout: cell (row r, col c), row 0 = top  <->  bit c*7 + (5-r).  Bit c*7+6 is a guard
// that is never set, so shifted AND-chains cannot wrap between columns.
#pragma once
#include <stdint.h>

#include "zc_common.cuh"

namespace zc {
namespace c4 {

constexpr int COLS = 7, ROWS = 6;
constexpr uint64_t COL0 = 0x3Full;                       // six playable cells of column 0
constexpr uint64_t FULL = 0x0FDFBF7EFDFBFull;            // all 42 cells
constexpr uint64_t TOPS = 0x0810204081020ull;            // top cell (h = 5) of every column

// Search-tree state: discs of the side to move and of the opponent (turn-relative, so that
// play() is a swap).  c4_backend.py:4 keeps absolute 'X'/'O' plus turn.
struct State {
    uint64_t cur, opp;
};

// CPython set-iteration order of the legal (col,0) tuples, c4_backend.py:49-50.
// order[mask] packs up to seven column numbers, 4 bits each, first move in the low nibble.
#ifdef __CUDACC__
__constant__ uint32_t d_order[128];
#endif
extern uint32_t h_order[128];

ZC_HD uint32_t order_of(int mask) {
#ifdef __CUDA_ARCH__
    return d_order[mask];
#else
    return h_order[mask];
#endif
}

// c4_backend.py:49-50 -- a column is playable iff its TOP cell is empty (wins are ignored)
ZC_HD int legal_mask(const State& s) {
    // free top cells sit at bits 5, 12, ..., 47; (x >> 5) has them at 7c.  Multiplying by
    // sum_c 2^(42-6c) moves bit 7c to 42+c; no two partial products share a position, so no carries.
    const uint64_t free_top = (~(s.cur | s.opp) & TOPS) >> 5;
    return (int)((free_top * 0x0000041041041040ull) >> 42) & 0x7F;
}
ZC_HD int n_moves(const State& s) { return zc_popc32((unsigned)legal_mask(s)); }
// column of the idx-th move in backend (set-iteration) order
ZC_HD int move_col(int mask, int idx) { return (int)((order_of(mask) >> (4 * idx)) & 0xFu); }

// c4_backend.py:14-23 -- the disc falls to the lowest EMPTY cell of the column (also correct for
// hand-made boards with gaps); a full column changes nothing but the turn.
ZC_HD State play(const State& s, int col) {
    const uint64_t empty = ~(s.cur | s.opp) & (COL0 << (7 * col));
    const uint64_t cell = empty & (0 - empty);
    State n;
    n.cur = s.opp;
    n.opp = s.cur | cell;
    return n;
}

ZC_HD bool four_in_a_row(uint64_t b) {
    uint64_t m;
    m = b & (b >> 7);  if (m & (m >> 14)) return true;   // horizontal
    m = b & (b >> 1);  if (m & (m >> 2)) return true;    // vertical
    m = b & (b >> 6);  if (m & (m >> 12)) return true;   // diagonal
    m = b & (b >> 8);  if (m & (m >> 16)) return true;   // anti-diagonal
    return false;
}
// c4_backend.py:25-44 -- four in a row of the player who has just moved (= opp here)
ZC_HD bool check_win(const State& s) { return four_in_a_row(s.opp); }
// c4_backend.py:46-47
ZC_HD bool check_draw(const State& s) { return ((s.cur | s.opp) & FULL) == FULL; }

// ZC_EVAL_C4_TERMINAL: terminal branch of random_rollout (value_functions.py:41-43)
ZC_HD double eval_terminal(const State& s) { return check_win(s) ? -1.0 : 0.0; }

// ZC_EVAL_C4_POSITIONAL: -1 on a win, else sum_discs w[col] * (+1 own / -1 opponent) / 64.
// Dyadic, so every value sum in the tree is exact in fp64 whatever the order.
ZC_HD double eval_positional(const State& s) {
    if (check_win(s)) return -1.0;
    int acc = 0;
#pragma unroll
    for (int c = 0; c < COLS; ++c) {
        const int w = c < 4 ? c + 1 : 7 - c;
        const uint64_t col = COL0 << (7 * c);
        acc += w * (zc_popc64(s.cur & col) - zc_popc64(s.opp & col));
    }
    return (double)acc * (1.0 / 64.0);
}

// ZC_EVAL_C4_ROLLOUT: random_rollout (value_functions.py:35-45): uniformly random legal moves until
// check_win or check_draw; -1 if the side to move at `s` ends up the loser, +1 if the winner, 0 on a
// draw; an already-won state returns -1 at once.  Device RNG: statistically, not bitwise, the
// reference (which draws from CPython's global Mersenne Twister).
ZC_HD uint64_t c4_rng_next(uint64_t& z) {
    z += 0x9E3779B97F4A7C15ull;
    uint64_t x = z;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
ZC_HD double eval_rollout(const State& s0, uint64_t key) {
    State s = s0;
    int flips = 0;   // parity of plies played: the loser is the side to move when a win is found
    for (;;) {
        if (check_win(s)) return (flips & 1) ? 1.0 : -1.0;
        if (check_draw(s)) return 0.0;
        const int mask = legal_mask(s), n = zc_popc32((unsigned)mask);
        if (n == 0) return 0.0;   // unreachable for boards without gaps (a full top row is a full board)
        const int pick = (int)((c4_rng_next(key) >> 33) % (uint64_t)n);
        s = play(s, move_col(mask, pick));
        ++flips;
    }
}

}  // namespace c4
}  // namespace zc
