// Connect Four plugged into the generic search (search.cuh).
#pragma once
#include "c4_rules.cuh"
#include "search.cuh"
#include "../../include/zc_b200.h"

namespace zc {

struct C4Game {
#ifndef ZC_C4_MINB
#define ZC_C4_MINB 7
#endif
    static constexpr int kMinBlocks = ZC_C4_MINB;   // resident 128-thread blocks per SM the fused search is compiled for
    using State = c4::State;
    // never reached: the host maps immediate_value to random for Connect Four (all move values are 0)
    ZC_D static int immediate_value_order(const uint4*, const State&, int, int, int, int j, float, uint64_t, int) { return j; }
    static constexpr int SS = 1;            // state slots per node
    static constexpr int FIRST_SLOTS = 9;   // header + state + up to 7 edges: the whole node in one warp load
    static constexpr int MAX_K = 7;         // moves per node
    static_assert(MAX_K <= KEYED_PERM_MAX, "Policy.random on the spine path orders a node's moves with keyed_perm");
    static constexpr int PLANE_ELEMS = 84;  // 2 x 6 x 7 (c4_backend.py:52-61)
    static constexpr int MOVE_SCRATCH = 0;
    static constexpr bool kCheapSpine = true;   // play + legal mask are a handful of instructions
    static constexpr bool kSmallCode = false;
    static constexpr bool kPhaseSync = false;   // the whole search fits the instruction cache (100 % hit rate measured)
    struct Ctx {};
    static constexpr int WARP_MOVES = 8;    // no staging list: moves are implied by the legal mask
    ZC_D static Ctx make_ctx(const SearchParams&, unsigned, int, uint16_t*) { return Ctx(); }
    ZC_D static int random_order(Ctx&, int, int, int, int j, uint64_t, int) { return j; }   // never reached: Connect Four orders with keyed_perm

    ZC_D static State state_from_lanes(const uint4& v) {
        const uint4 s = shfl4(v, 1);
        State st;
        st.cur = ((uint64_t)s.y << 32) | s.x;
        st.opp = ((uint64_t)s.w << 32) | s.z;
        return st;
    }
    ZC_HD static void store_state(uint4* dst, const State& s) {
        *dst = make_uint4((uint32_t)s.cur, (uint32_t)(s.cur >> 32), (uint32_t)s.opp, (uint32_t)(s.opp >> 32));
    }
    ZC_HD static State load_state(const uint4* src) {
        const uint4 s = *src;
        State st;
        st.cur = ((uint64_t)s.y << 32) | s.x;
        st.opp = ((uint64_t)s.w << 32) | s.z;
        return st;
    }
    ZC_D static State shfl_state(const State& s, int src) {
        State r;
        r.cur = __shfl_sync(FULL_MASK, s.cur, src);
        r.opp = __shfl_sync(FULL_MASK, s.opp, src);
        return r;
    }
    ZC_HD static uint64_t state_key(const State& s, uint32_t) { return s.cur * 0x9E3779B97F4A7C15ull ^ (s.opp + 0xD1B54A32D192ED03ull) * 0xBF58476D1CE4E5B9ull; }
    ZC_HD static int move_slots(int) { return 0; }                    // moves are implied by the legal mask
    ZC_HD static int move_value(const uint4*, const State&, int, int) { return 0; }   // c4_backend.py:50: every move is (col, 0)
    ZC_HD static void store_moves(Ctx&, uint4*, int) {}
    // state after the ei-th move (backend order) of `parent`
    ZC_HD static State child(const State& parent, uint32_t, const uint4*, int, int ei, uint32_t& cmisc) {
        cmisc = 0;
        return c4::play(parent, c4::move_col(c4::legal_mask(parent), ei));
    }
    ZC_HD static int count_moves(Ctx&, const State& s, uint32_t) { return c4::n_moves(s); }
    ZC_D static int count_moves_serial(const State& s, uint32_t) { return c4::n_moves(s); }
    ZC_HD static double eval(const State& s, uint32_t, int evaluator, uint64_t key) {
        if (evaluator == ZC_EVAL_C4_POSITIONAL) return c4::eval_positional(s);
        if (evaluator == ZC_EVAL_C4_ROLLOUT) return c4::eval_rollout(s, key);
        return c4::eval_terminal(s);
    }
    ZC_HD static double eval_child(const State& s, uint32_t m, int, int evaluator, uint64_t key) { return eval(s, m, evaluator, key); }

    // leaves -> network input rows, by the whole warp: plane 0 = side to move, plane 1 = opponent, [r][c] with r = 0 the
    // top row (c4_backend.py:52-61).  Lane i holds leaf i (row row0 + i); rows are written one after the other with lane q
    // storing the row's q-th group of four cells, so every store instruction covers one contiguous run of the batch.
    ZC_D static void pack_planes(void* planes, int dtype, size_t row0, int n_valid, int n_rows, const State& s, uint32_t, int lane) {
        constexpr int GROUPS = PLANE_ELEMS / 4;                 // 21 groups of four cells per row
        const uint32_t one16 = dtype == 2 ? 0x3C00u : 0x3F80u;   // f16 : bf16
        for (int i = 0; i < n_rows; ++i) {
            const State ls = shfl_state(s, i);
            if (lane < GROUPS) {
                uint32_t bits = 0;
                if (i < n_valid) {
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const int idx = lane * 4 + t, pl = idx / 42, cell = idx % 42, r = cell / 7, c = cell % 7;
                        const uint64_t bb = pl == 0 ? ls.cur : ls.opp;
                        bits |= (uint32_t)((bb >> (c * 7 + (5 - r))) & 1ull) << t;
                    }
                }
                const size_t o = (row0 + (size_t)i) * GROUPS + (size_t)lane;
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
