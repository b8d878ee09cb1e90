// Node arena layout in HBM.
//
// The reference (engine/mcts/src/mcts.cpp:10-39) keeps one heap `Node` per tree node with six
// std::vectors.  Here every tree owns one contiguous arena of 16-byte slots and a node is a run
// of slots written once and never moved:
//
//   slot 0            header   {N, k | n_expanded<<16, parent slot, parent_edge | misc<<8 | depth<<16}
//   slot 1..SS        game state (C4: {cur, opp}; chess: four bit planes of the piece codes)
//   slot 1+SS+i       edge i   {Wa (fp64), Na (i32), child slot (u32)}      i = 0..k-1, backend move order
//   then (chess)      k packed moves, 8 per slot
//
// so one coalesced warp load (lane L reads slot L) fetches header, state and the first edges of a
// node, and a child link is a 32-bit slot index inside the tree's arena.  `untried`
// (mcts.cpp:17) is not stored: moves are expanded in an order that is a pure function of
// (policy, k, j), so "untried" is {order(j) : j >= n_expanded}.  `Qa` (mcts.cpp:15) is Wa/Na
// recomputed with the same IEEE division the reference used when it stored it (mcts.cpp:93).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "zc_common.cuh"

namespace zc {

struct __align__(16) TreeCtl {
    uint32_t top;        // next free slot
    uint32_t nodes;      // nodes in the tree
    uint32_t sims_done;
    int32_t status;      // 0 or ZC_ECAPACITY
    unsigned long long sum_leaf_depth;
    uint32_t max_leaf_depth;
    uint32_t reevaluated;
    unsigned long long sum_path_children;
    uint32_t root_turn;  // C4: whose discs `cur` are at the root (for readout)
    uint32_t tree_id;    // RNG stream id (set at set_roots)
};

// split-phase (external evaluator) hand-over between select and backprop
struct __align__(16) Pending {
    int32_t B;           // leaves in this batch (0 = nothing pending)
    int32_t D;           // deepest path level
    uint32_t info[32];   // per leaf, see LEAF_* below
};

// Games with costly move generation (chess) create a leaf as a STUB: header + state only, k = K_UNKNOWN.  The
// reference lists a node's moves when it creates the node (mcts.cpp:74), but a leaf's list is never looked at
// unless the search later expands below it -- about one node in thirty.  A stub is MATERIALISED (moves
// generated, a full node written at the arena top, the parent's child link redirected) the first time the
// search needs its moves; a stub with no moves becomes a move-less node in place.  Results are unchanged.
constexpr uint32_t K_UNKNOWN = 0xFFFFu;

constexpr uint32_t LEAF_LEVEL_MASK = 0xFFFFu;  // level of the leaf's parent (self leaf: its own level)
constexpr int LEAF_EDGE_SHIFT = 16;            // 8 bits: edge index at the parent
constexpr uint32_t LEAF_PATH = 1u << 30;       // the leaf is itself the next node of the chain
constexpr uint32_t LEAF_SELF = 1u << 31;       // a move-less node evaluated again (mcts.cpp:138-141)

ZC_HD uint32_t hdr_k(const uint4& h) { return h.y & 0xFFFFu; }
ZC_HD uint32_t hdr_nexp(const uint4& h) { return h.y >> 16; }
ZC_HD uint32_t hdr_parent_edge(const uint4& h) { return h.w & 0xFFu; }
ZC_HD uint32_t hdr_misc(const uint4& h) { return (h.w >> 8) & 0xFFu; }
ZC_HD uint32_t hdr_depth(const uint4& h) { return h.w >> 16; }
ZC_HD uint4 make_hdr(uint32_t N, uint32_t k, uint32_t nexp, uint32_t parent, uint32_t pedge, uint32_t misc,
                     uint32_t depth) {
    uint4 h;
    h.x = N;
    h.y = k | (nexp << 16);
    h.z = parent;
    h.w = pedge | (misc << 8) | (depth << 16);
    return h;
}

ZC_HD double edge_W(const uint4& e) {
    unsigned long long b = ((unsigned long long)e.y << 32) | e.x;
#ifdef __CUDA_ARCH__
    return __longlong_as_double((long long)b);
#else
    double d;
    __builtin_memcpy(&d, &b, 8);
    return d;
#endif
}
ZC_HD void edge_set_W(uint4& e, double w) {
    unsigned long long b;
#ifdef __CUDA_ARCH__
    b = (unsigned long long)__double_as_longlong(w);
#else
    __builtin_memcpy(&b, &w, 8);
#endif
    e.x = (uint32_t)b;
    e.y = (uint32_t)(b >> 32);
}

// same mixing function as oracle/zc_oracle.c:mix64 (the tree hash must agree bit for bit)
ZC_HD unsigned long long mix64(unsigned long long h, unsigned long long v) {
    h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    h *= 0xFF51AFD7ED558CCDull;
    h ^= h >> 33;
    return h;
}

}  // namespace zc
