// Chess plugged into the generic search (search.cuh).
#pragma once
#include "chess_rules.cuh"
#include "search.cuh"
#include "../../include/zc_b200.h"

namespace zc {

struct ChessGame {
#ifndef ZC_CHESS_MINB
#define ZC_CHESS_MINB 4       // 16 warps per SM at 128 registers: +7 % over 8 blocks at 64 (spills, instruction-cache contention), A/B in profiles/
#endif
    static constexpr int kMinBlocks = ZC_CHESS_MINB;   // resident 128-thread blocks per SM the fused search is compiled for
    using State = chess::Board;
    static constexpr int SS = 2;             // four bit planes = 32 bytes
    static constexpr int FIRST_SLOTS = 32;   // header + state + first 29 edges in one warp load
    static constexpr int PLANE_ELEMS = 17 * 64;
    static constexpr bool kCheapSpine = false;  // a chain node needs a full move generation: expand level by level
    static constexpr bool kSmallCode = true;    // the descent's argmax butterfly stays rolled: the search is bound by instruction fetch
    static constexpr bool kPhaseSync = true;    // with -DZC_PHASE_SYNC (off: measured slower) the warps of a block start every batch together
    static constexpr int MOVE_SCRATCH = chess::MAX_PSEUDO;   // pseudo-legal staging (>= 218 legal), multiple of 8

    static constexpr int WARP_MOVES = chess::MAX_PSEUDO;   // entries of the per-warp staging list in shared memory
    struct Ctx {
        uint16_t* moves;   // this lane's move list (serial generator, cold paths): entry i at moves[i * stride], global memory
        int stride;        // 32 in the search: the lanes of a warp interleave their lists, lane-parallel accesses coalesce
        uint16_t* wmoves;  // the warp generator's list for ONE position: contiguous, in shared memory
    };
    ZC_D static Ctx make_ctx(const SearchParams& p, unsigned warp_slot, int lane, uint16_t* warp_moves) {
        Ctx c;
        c.moves = p.scratch + (size_t)warp_slot * 32 * MOVE_SCRATCH + (size_t)lane;
        c.stride = 32;
        c.wmoves = warp_moves;
        return c;
    }
    ZC_D static State state_from_lanes(const uint4& v) {
        const uint4 a = shfl4(v, 1), b = shfl4(v, 2);
        State s;
        s.p0 = ((uint64_t)a.y << 32) | a.x;
        s.p1 = ((uint64_t)a.w << 32) | a.z;
        s.p2 = ((uint64_t)b.y << 32) | b.x;
        s.p3 = ((uint64_t)b.w << 32) | b.z;
        return s;
    }
    ZC_HD static void store_state(uint4* dst, const State& s) {
        dst[0] = make_uint4((uint32_t)s.p0, (uint32_t)(s.p0 >> 32), (uint32_t)s.p1, (uint32_t)(s.p1 >> 32));
        dst[1] = make_uint4((uint32_t)s.p2, (uint32_t)(s.p2 >> 32), (uint32_t)s.p3, (uint32_t)(s.p3 >> 32));
    }
    ZC_HD static State load_state(const uint4* src) {
        const uint4 a = src[0], b = src[1];
        State s;
        s.p0 = ((uint64_t)a.y << 32) | a.x;
        s.p1 = ((uint64_t)a.w << 32) | a.z;
        s.p2 = ((uint64_t)b.y << 32) | b.x;
        s.p3 = ((uint64_t)b.w << 32) | b.z;
        return s;
    }
    ZC_D static State shfl_state(const State& s, int src) {
        State r;
        r.p0 = __shfl_sync(FULL_MASK, s.p0, src);
        r.p1 = __shfl_sync(FULL_MASK, s.p1, src);
        r.p2 = __shfl_sync(FULL_MASK, s.p2, src);
        r.p3 = __shfl_sync(FULL_MASK, s.p3, src);
        return r;
    }
    ZC_HD static uint64_t state_key(const State& s, uint32_t misc) {
        return (s.p0 * 0x9E3779B97F4A7C15ull) ^ (s.p1 * 0xBF58476D1CE4E5B9ull) ^ (s.p2 * 0x94D049BB133111EBull) ^ (s.p3 * 0xD1B54A32D192ED03ull) ^ misc;
    }
    ZC_HD static int move_slots(int k) { return (k + 7) >> 3; }
    // packed move i of the node at `node` (k moves): stored after the edges, 8 per slot
    ZC_HD static uint16_t move_at(const uint4* node, int k, int i) {
        const uint16_t* mv = reinterpret_cast<const uint16_t*>(node + 1 + SS + k);
        return mv[i];
    }
    // the value the reference attaches to move i: what it captures (chess_backend.cpp:50-64)
    ZC_HD static int move_value(const uint4* node, const State& st, int k, int i) {
        return chess::capture_value(chess::piece_at(st, chess::move_to(move_at(node, k, i))));
    }
    ZC_HD static State child(const State& parent, uint32_t pmisc, const uint4* pnode, int pk, int ei, uint32_t& cmisc) {
        const uint16_t m = move_at(pnode, pk, ei);
        return chess::play_generated(parent, pmisc, chess::move_from(m), chess::move_to(m), cmisc);   // the node's own generated move
    }
    // Policy.random (policy_functions.py:10-12) as mcts.cpp:65-78 applies it: every expansion draws uniformly from the node's
    // untried moves, i.e. the node's moves are expanded in a uniformly random order.  That order is fixed per node by giving
    // move i the key hash(node key, i) and expanding in ascending key order (ties by index) -- a uniformly random
    // permutation, the same distribution as the sequential draws (chi-square tests on first picks and pairs,
    // tests/test_gpu_parity_bench_sets.py).  The warp ranks the keys all against all (lane L owns moves L, L+32, ...) and
    // lane j receives the move of rank nexp + j through the warp's shared staging list.  Called by the whole warp.
    // (out of line: only Policy.random runs it)
    __device__ __noinline__ static int random_order(Ctx& gx, int k, int nexp, int m, int j, uint64_t key, int lane) {
        constexpr int Q = (ZC_MAX_MOVES + 31) / 32;
        uint32_t c[Q];
        int rank[Q];
        const int q_used = (k + 31) >> 5;
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const int i = lane + 32 * q;
            c[q] = i < k ? (((uint32_t)rng_mix(key ^ (0xD1B54A32D192ED03ull * (uint64_t)(i + 1))) & 0xFFFFFF00u) | (uint32_t)i) : 0xFFFFFFFFu;
            rank[q] = 0;
        }
#pragma unroll
        for (int qs = 0; qs < Q; ++qs) {
            if (qs < q_used) {
                for (int src = 0; src < 32; ++src) {
                    const uint32_t other = __shfl_sync(FULL_MASK, c[qs], src);
#pragma unroll
                    for (int q = 0; q < Q; ++q) rank[q] += other < c[q] ? 1 : 0;
                }
            }
        }
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const int r = rank[q] - nexp;
            if (lane + 32 * q < k && r >= 0 && r < m) gx.wmoves[r] = (uint16_t)(lane + 32 * q);
        }
        __syncwarp();
        const int mine = (j >= 0 && j < m) ? (int)gx.wmoves[j] : 0;
        __syncwarp();
        return mine;
    }
    // Policy.immediate_value (policy_functions.py:14-17) as mcts.cpp:65-78 applies it: every expansion of a node
    // picks uniformly among its untried moves whose capture value move[1] is >= (best untried value - freedom).
    // Which move the t-th expansion of a node picks is a function of the node (its key seeds the draws), so the
    // warp replays picks 0 .. nexp+m-1 (lane L owns moves L, L+32, ...) and lane j receives the move of
    // expansion nexp + j.  Called by the whole warp.
    // (out of line: only Policy.immediate_value runs it)
    __device__ __noinline__ static int immediate_value_order(const uint4* node, const State& st, int k, int nexp, int m, int j, float freedom,
                                          uint64_t key, int lane) {
        constexpr int Q = (ZC_MAX_MOVES + 31) / 32;
        int val[Q];
        uint32_t untried = 0;                       // bit q: move lane + 32q exists and has not been picked
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const int i = lane + 32 * q;
            val[q] = -1;
            if (i < k) {
                val[q] = chess::capture_value(chess::piece_at(st, chess::move_to(move_at(node, k, i))));
                untried |= 1u << q;
            }
        }
        int mine = 0;
        for (int t = 0; t < nexp + m; ++t) {
            int best = -1;
#pragma unroll
            for (int q = 0; q < Q; ++q)
                if (untried >> q & 1u) best = max(best, val[q]);
#pragma unroll
            for (int d = 16; d >= 1; d >>= 1) best = max(best, __shfl_xor_sync(FULL_MASK, best, d));
            const float threshold = (float)best - freedom;
            uint32_t cand = 0;
#pragma unroll
            for (int q = 0; q < Q; ++q)
                if ((untried >> q & 1u) && (float)val[q] >= threshold) cand |= 1u << q;
            int total;
            const int before = warp_excl_scan(__popc(cand), lane, total);
            const int r = (int)(pick_draw(key, t) % (uint64_t)max(total, 1));
            int pick = -1;
            if (r >= before && r < before + __popc(cand)) {
                uint32_t c = cand;
                for (int s = r - before; s > 0; --s) c &= c - 1;     // the (r - before)-th candidate of this lane
                const int q = __ffs((int)c) - 1;
                untried &= ~(1u << q);
                pick = lane + 32 * q;
            }
            const unsigned who = __ballot_sync(FULL_MASK, pick >= 0);
            const int chosen = __shfl_sync(FULL_MASK, pick, who ? __ffs((int)who) - 1 : 0);
            if (t - nexp == j) mine = chosen;
        }
        return mine;
    }
    // number of legal moves of a position, one thread, own storage (the test-only tree hash)
    ZC_D static int count_moves_serial(const State& s, uint32_t misc) {
        uint16_t mv[chess::MAX_PSEUDO];
        return chess::generate(s, (int)(misc & chess::MISC_TURN), mv);
    }
    static constexpr bool kLazyMoves = true;    // leaves are stubs; moves are generated when a node is first expanded
    // the move list of ONE position by the whole warp, left contiguous in the warp's shared-memory staging list
    // (out of line: three call sites, and the generator is the largest piece of code in the kernel)
#ifndef ZC_AB_MOVES_ATTR
#define ZC_AB_MOVES_ATTR __noinline__
#endif
    __device__ ZC_AB_MOVES_ATTR static int moves_warp(Ctx& gx, const State& s, uint32_t misc, int lane, bool any_only = false) {
        return chess::generate_warp(s, (int)(misc & chess::MISC_TURN), gx.wmoves, 1, lane, any_only);
    }
    // ... and packed into a node's move slots by the whole warp (8 moves per slot)
    ZC_D static void store_moves_warp(Ctx& gx, uint4* dst, int k, int lane) {
        const uint4* src = reinterpret_cast<const uint4*>(gx.wmoves);
#pragma unroll 1
        for (int i = lane; i < move_slots(k); i += 32) dst[i] = src[i];
    }
    // crude_chess_score of freshly created children (value_functions.py:49-55), one child per active lane, without
    // their move lists: check_win = no legal move AND king attacked (chess_backend.cpp:404-412), so only a child
    // whose side to move is in check needs to know whether it has a move -- the warp generates those lists in turn.
    ZC_D static double eval_stubs(Ctx& gx, const State& s, uint32_t misc, bool act, int lane) {
        const int turn = (int)(misc & 1u);
        const bool chk = act && chess::in_check(s, turn);
        double v = act ? (double)((turn ? -1 : 1) * chess::material(s)) : 0.0;
        unsigned todo = __ballot_sync(FULL_MASK, chk);
        while (todo) {
            const int owner = __ffs((int)todo) - 1;
            todo &= todo - 1;
            const State os = shfl_state(s, owner);
            const uint32_t om = __shfl_sync(FULL_MASK, misc, owner);
            const int k = moves_warp(gx, os, om, lane, true);      // in check: only "any legal move?" is needed
            if (lane == owner && k == 0) v = 1000.0;
        }
        return v;
    }
    ZC_HD static int count_moves(Ctx& gx, const State& s, uint32_t misc) {
        return chess::generate(s, (int)(misc & chess::MISC_TURN), gx.moves, gx.stride);
    }
    ZC_HD static void store_moves(Ctx& gx, uint4* dst, int k) {
        for (int i = 0; i < move_slots(k); ++i) {
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {      // entries beyond k are whatever the staging holds; never read back
                const uint32_t lo = gx.moves[(size_t)(8 * i + 2 * e) * gx.stride], hi = gx.moves[(size_t)(8 * i + 2 * e + 1) * gx.stride];
                w[e] = lo | (hi << 16);
            }
            dst[i] = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
    ZC_HD static double eval(const State& s, uint32_t misc, int, uint64_t) { return chess::crude_score(s, (int)(misc & 1u), 0); }
    ZC_HD static double eval_child(const State& s, uint32_t misc, int k, int, uint64_t) { return chess::crude_score(s, (int)(misc & 1u), k); }

    // chess_backend.cpp:461-521 -- 17 planes x 64 cells; planes 0-11 'PNBRQKpnbrqk', 12 white to move,
    // 13-16 castling flags.  Cell order r*8+c = bit order of the boards.
    ZC_HD static uint64_t plane_bits(const State& s, uint32_t misc, int pl) {
        if (pl < 12) {
            const uint64_t t = chess::of_type(s, pl % 6 + 1);
            return t & (pl < 6 ? ~s.p3 : s.p3) & chess::occupied(s);
        }
        const bool on = pl == 12 ? !(misc & chess::MISC_TURN) : (misc >> (pl - 12)) & 1u;
        return on ? ~0ull : 0ull;
    }
    // By the whole warp: lane i holds leaf i (row row0 + i); rows are written one after the other, lane q storing groups
    // q, q + 32, ... of four cells (17 planes x 16 groups per row), so every store instruction covers a contiguous run.
    ZC_D static void pack_planes(void* planes, int dtype, size_t row0, int n_valid, int n_rows, const State& s, uint32_t misc, int lane) {
        constexpr int GROUPS = PLANE_ELEMS / 4;                 // 272
        const uint32_t one16 = dtype == 2 ? 0x3C00u : 0x3F80u;
        for (int i = 0; i < n_rows; ++i) {
            const State ls = shfl_state(s, i);
            const uint32_t lm = __shfl_sync(FULL_MASK, misc, i);
            const bool valid = i < n_valid;
#pragma unroll 1
            for (int g = lane; g < GROUPS; g += 32) {
                const int pl = g >> 4, q = g & 15;
                const uint64_t bb = valid ? plane_bits(ls, lm, pl) : 0ull;
                const uint32_t bits = (uint32_t)(bb >> (4 * q)) & 0xFu;
                const size_t o = (row0 + (size_t)i) * GROUPS + (size_t)g;
                if (dtype == 1) {
                    reinterpret_cast<float4*>(planes)[o] =
                        make_float4(bits & 1 ? 1.f : 0.f, bits & 2 ? 1.f : 0.f, bits & 4 ? 1.f : 0.f, bits & 8 ? 1.f : 0.f);
                } else {
                    uint2 h;
                    h.x = (bits & 1 ? one16 : 0u) | (bits & 2 ? one16 << 16 : 0u);
                    h.y = (bits & 4 ? one16 : 0u) | (bits & 8 ? one16 << 16 : 0u);
                    reinterpret_cast<uint2*>(planes)[o] = h;
                }
            }
        }
    }
};

}  // namespace zc
