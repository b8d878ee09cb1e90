// PUCT selection with stored priors and virtual loss -- an OPT-IN search mode next to the reference's UCB1
// (search.cuh).  The reference has no such mode (engine/mcts/src/mcts.cpp:41-63 is UCB1 over frozen batches); BASELINE.json's
// north star names it (priors in the node pools, one warp per tree, warp-shuffle argmax, virtual-loss backup), so it is
// provided as ZC_SELECT_PUCT with its own CPU restatement in oracle/zc_oracle.c (zo_search_puct) and the same
// whole-tree-hash parity tests.  It never touches the bit-exact UCB1 path.
//
// Definition (one simulation; `vl` = virtual loss, `c` = c_puct):
//   descend from the root:  at a node with k > 0 moves, visits N and per-edge (Na, Wa, P):
//       a* = argmax_a  Q(a) + c * P(a) * sqrt(N + 1) / (1 + Na(a)),   Q(a) = Wa/Na (0 if Na = 0),  lowest index on ties
//       N += 1, Na(a*) += 1, Wa(a*) -= vl                    (virtual loss: later simulations of the batch look elsewhere)
//       if the edge has no child: create it (one new node per simulation, as in the reference) -- it is the leaf; else go down
//     a node without moves is its own leaf (evaluated again, mcts.cpp:59,138-141);  leaf.N += 1
//   a batch selects `batch_size` leaves this way, one after the other, evaluates them together, then backs them up in
//   order:  for every edge of the leaf's path, from the leaf upwards:  Wa += vl;  Wa -= result;  result = -result.
// Priors are stored with the node (one float per edge, after the edges / moves) when the node gets its move list:
//   P(a) = w(a) / sum_b w(b),  w(a) = 1 + prior_weight * move_value(a)   (integers; move_value = the capture value the
//   reference attaches to a move, chess_backend.cpp:50-64 -- what Policy.immediate_value looks at; 0 for Connect Four,
//   so uniform priors).  zc_search_set_root_priors overwrites the root's (an external policy head, exploration noise).
//
// Mapping: one warp per tree.  A descent step is one coalesced warp load of the node (lane L = slot L), lane = edge for the
// score, butterfly-shuffle argmax; lane 0 applies the virtual loss.  Every simulation's path is kept (64 entries per
// simulation of the batch) so the backup is lane = path level: all levels of one simulation in one step, simulations
// in order (the same fp64 operation sequence as the oracle).  Chess leaves are stubs, materialised on first descent.
#pragma once
#include "search.cuh"

namespace zc {

constexpr int PUCT_MAX_DEPTH = 63;                   // deepest leaf a PUCT search may create (path segment = 64 entries)
constexpr int PUCT_PATH_STRIDE = PUCT_MAX_DEPTH + 1;
constexpr uint32_t PUCT_PATH_CAP = 32u * PUCT_PATH_STRIDE;

// PUCT score of one edge, every operation rounded on its own (the oracle is compiled without contraction)
ZC_D double puct_score(double W, int Na, float P, double sqrtN, double c) {
    const double q = Na ? __ddiv_rn(W, (double)Na) : 0.0;
    const double u = __ddiv_rn(__dmul_rn(__dmul_rn(c, (double)P), sqrtN), (double)(1 + Na));
    return __dadd_rn(q, u);
}

struct PuctLeaf {
    uint32_t slot;
    uint32_t misc;
    int len;        // path entries of this simulation
    int k;          // moves of the leaf (games whose nodes are complete at creation), else 0
    bool self;      // a move-less node evaluated again
    bool fresh;     // created by this simulation
};

// One simulation's descent.  Returns false on an error (tree flagged).  All lanes return the same values.
template <class G>
ZC_D bool puct_descend(const SearchParams& p, typename G::Ctx& gx, uint4* __restrict__ arena, uint2* __restrict__ seg, TreeCtl& ctl,
                       int lane, PuctLeaf& leaf, typename G::State& leaf_st) {
    uint32_t node = 0;
    int depth = 0;
    for (;;) {
        uint4* np = arena + node;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (lane < G::FIRST_SLOTS) v = np[lane];
        const uint4 hdr = shfl4(v, 0);
        typename G::State st = G::state_from_lanes(v);
        const uint32_t misc = hdr_misc(hdr);
        int k = (int)hdr_k(hdr);
        if constexpr (!G::kCheapSpine) {
            if (k == (int)K_UNKNOWN) {                   // a leaf of an earlier simulation: its moves are needed now
                if (!materialize<G>(p, gx, arena, ctl, lane, node, st, misc, k, true)) return false;
                continue;                                // reload it from where it lives now
            }
        }
        if (k == 0) {                                    // no moves: the node is its own leaf
            leaf.slot = node; leaf.misc = misc; leaf.len = depth; leaf.k = 0; leaf.self = true; leaf.fresh = false;
            leaf_st = st;
            if (lane == 0) np[0].x = hdr.x + 1u;
            __syncwarp();
            ctl.reevaluated += 1;
            return true;
        }
        const double sqrtN = __dsqrt_rn((double)(hdr.x + 1u));
        const float* pri = node_priors<G>(np, k);
        double best = -CUDART_INF;
        int best_e = 0x7FFFFFFF;
        uint32_t best_child = 0;
        for (int slot = lane; slot < 1 + G::SS + k; slot += 32) {
            const int e = slot - (1 + G::SS);
            if (e < 0) continue;
            const uint4 ev = slot < G::FIRST_SLOTS ? v : np[slot];
            const double s = puct_score(edge_W(ev), (int)ev.z, pri[e], sqrtN, p.c);
            if (s > best) { best = s; best_e = e; best_child = ev.w; }
        }
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            const double ob = __shfl_xor_sync(FULL_MASK, best, d);
            const int oe = __shfl_xor_sync(FULL_MASK, best_e, d);
            const uint32_t oc = __shfl_xor_sync(FULL_MASK, best_child, d);
            if (ob > best || (ob == best && oe < best_e)) { best = ob; best_e = oe; best_child = oc; }
        }
        if (depth >= PUCT_MAX_DEPTH) { ctl.status = -4; return false; }
        uint32_t child = best_child;
        typename G::State cs = st;
        uint32_t cmisc = 0;
        int ck = 0, total = 0;
        if (child == 0) {                                // one new node per simulation (mcts.cpp:65-78)
            cs = G::child(st, misc, np, k, best_e, cmisc);
            if constexpr (G::kCheapSpine) {              // complete at creation: header, state, zeroed edges, priors
                ck = G::count_moves(gx, cs, cmisc);
                total = node_slots<G>(ck, true);
            } else {
                total = 1 + G::SS;                       // stub
            }
            if ((uint64_t)ctl.top + (uint64_t)total > p.arena_slots) { ctl.status = -4; return false; }
            child = ctl.top;
            for (int t = lane; t < total; t += 32) arena[child + t] = make_uint4(0, 0, 0, 0);
            __syncwarp();
            if (lane == 0) {
                arena[child] = make_hdr(1, G::kCheapSpine ? (uint32_t)ck : K_UNKNOWN, 0, node, (uint32_t)best_e, cmisc, (uint32_t)(depth + 1));
                G::store_state(arena + child + 1, cs);
            }
            __syncwarp();
            if constexpr (G::kCheapSpine) write_priors<G>(arena + child, cs, ck, p.prior_weight, lane);
            ctl.top += (uint32_t)total;
            ctl.nodes += 1;
        }
        if (lane == 0) {                                 // virtual loss on the chosen edge, the node's visit, the path entry
            uint4* ep = np + 1 + G::SS + best_e;
            uint4 e = *ep;
            edge_set_W(e, __dsub_rn(edge_W(e), p.vloss));
            e.z += 1u;
            if (best_child == 0) {
                e.w = child;
                np[0] = make_uint4(hdr.x + 1u, hdr.y + (1u << 16), hdr.z, hdr.w);      // N += 1, one more child
            } else {
                np[0].x = hdr.x + 1u;
            }
            *ep = e;
            seg[depth] = make_uint2(node, (uint32_t)best_e);
        }
        __syncwarp();
        ++depth;
        if (best_child == 0) {
            leaf.slot = child; leaf.misc = cmisc; leaf.len = depth; leaf.k = ck; leaf.self = false; leaf.fresh = true;
            leaf_st = cs;
            return true;
        }
        node = child;
    }
}

// select `B` leaves one after the other; lane i ends up holding leaf i
template <class G, bool kBuiltinEval>
ZC_D bool puct_select_batch(const SearchParams& p, typename G::Ctx& gx, uint4* __restrict__ arena, uint2* __restrict__ path, TreeCtl& ctl,
                            int B, int lane, uint32_t& my_info, typename G::State& my_st, uint32_t& my_misc, double& my_value) {
    my_info = 0;
    my_value = 0.0;
    my_misc = 0;
    int my_k = 0;
    bool my_self = false, act = false;
    for (int i = 0; i < B; ++i) {
        PuctLeaf leaf;
        typename G::State st;
        if (!puct_descend<G>(p, gx, arena, path + (size_t)i * PUCT_PATH_STRIDE, ctl, lane, leaf, st)) return false;
        ctl.sum_leaf_depth += (unsigned)leaf.len;
        ctl.max_leaf_depth = max(ctl.max_leaf_depth, (uint32_t)leaf.len);
        if (lane == i) {
            my_info = (uint32_t)leaf.len | (leaf.self ? LEAF_SELF : 0u) | (1u << 29);      // bit 29: this lane holds a leaf
            my_st = st;
            my_misc = leaf.misc;
            my_k = leaf.k;
            my_self = leaf.self;
            act = true;
        }
    }
    if constexpr (kBuiltinEval) {
        const uint64_t tkey = p.seed ^ ((uint64_t)ctl.tree_id << 32);
        const uint64_t rkey = rng_mix(tkey ^ G::state_key(my_st, my_misc) ^ ((uint64_t)ctl.sims_done << 40) ^ ((uint64_t)lane << 8));
        if constexpr (G::kCheapSpine) {
            if (act) my_value = my_self ? G::eval(my_st, my_misc, p.evaluator, rkey) : G::eval_child(my_st, my_misc, my_k, p.evaluator, rkey);
        } else {
            const double v = G::eval_stubs(gx, my_st, my_misc, act && !my_self, lane);     // whole warp: mate tests take turns
            if (act) my_value = my_self ? G::eval(my_st, my_misc, p.evaluator, rkey) : v;
        }
    }
    return true;
}

// back the batch up in order; lane l owns path level l of the simulation being applied
template <class G>
ZC_D void puct_backup_batch(const SearchParams& p, uint4* __restrict__ arena, const uint2* __restrict__ path, int B, int lane, uint32_t info,
                            double value) {
    for (int i = 0; i < B; ++i) {
        const int len = (int)(__shfl_sync(FULL_MASK, info, i) & LEAF_LEVEL_MASK);
        const double v = __shfl_sync(FULL_MASK, value, i);
        for (int base = 0; base < len; base += 32) {
            const int l = base + lane;
            if (l < len) {
                const uint2 pe = path[(size_t)i * PUCT_PATH_STRIDE + l];
                uint4* ep = arena + pe.x + 1 + G::SS + pe.y;
                uint4 e = *ep;
                const double r = ((len - 1 - l) & 1) ? -v : v;          // result = -result per level, from the leaf upwards
                edge_set_W(e, __dsub_rn(__dadd_rn(edge_W(e), p.vloss), r));
                *ep = e;
            }
        }
        __syncwarp();
    }
}

// ---- kernels: the same three shapes as search.cuh
template <class G, int MINB = G::kMinBlocks>
__global__ void __launch_bounds__(SEARCH_BLOCK, MINB) k_search_fused_puct(SearchParams p) {
    __shared__ __align__(16) uint16_t warp_moves[SEARCH_BLOCK / 32][G::WARP_MOVES];   // chess: the warp generator's move list
    const int lane = threadIdx.x & 31;
    for (;;) {
        int tree = 0;
        if (lane == 0) tree = (int)atomicAdd(p.work_counter, 1u);
        tree = __shfl_sync(FULL_MASK, tree, 0);
        if (tree >= p.n_trees) return;
        uint4* arena = p.arena + (uint64_t)tree * p.arena_slots;
        uint2* path = p.path + (uint64_t)tree * p.path_cap;
        TreeCtl ctl = p.ctl[tree];
        typename G::Ctx gx = G::make_ctx(p, (blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5, lane, warp_moves[threadIdx.x >> 5]);
        for (int done = 0; done < p.simulations && ctl.status == 0;) {
            const int B = min(p.batch_size, p.simulations - done);
            uint32_t info, misc;
            typename G::State st;
            double value;
            if (!puct_select_batch<G, true>(p, gx, arena, path, ctl, B, lane, info, st, misc, value)) break;
            puct_backup_batch<G>(p, arena, path, B, lane, info, value);
            done += B;
            ctl.sims_done += (uint32_t)B;
        }
        if (lane == 0) p.ctl[tree] = ctl;
    }
}

template <class G>
__global__ void __launch_bounds__(SEARCH_BLOCK) k_select_puct(SearchParams p, int sims_left) {
    __shared__ __align__(16) uint16_t warp_moves[SEARCH_BLOCK / 32][G::WARP_MOVES];   // chess: the warp generator's move list
    const int lane = threadIdx.x & 31;
    const int tree = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    if (tree >= p.n_trees) return;
    uint4* arena = p.arena + (uint64_t)tree * p.arena_slots;
    uint2* path = p.path + (uint64_t)tree * p.path_cap;
    TreeCtl ctl = p.ctl[tree];
    Pending* pd = p.pending + tree;
    const int B = ctl.status == 0 ? min(p.batch_size, sims_left) : 0;
    uint32_t info = 0, misc = 0;
    typename G::State st = typename G::State();
    double value = 0.0;
    bool ok = B > 0;
    typename G::Ctx gx = G::make_ctx(p, (unsigned)tree, lane, warp_moves[threadIdx.x >> 5]);
    if (ok) ok = puct_select_batch<G, false>(p, gx, arena, path, ctl, B, lane, info, st, misc, value);
    if (lane == 0) {
        pd->B = ok ? B : 0;
        pd->D = 0;
        p.ctl[tree] = ctl;
    }
    pd->info[lane] = ok ? info : 0u;
    G::pack_planes(p.planes, p.plane_dtype, (size_t)tree * (size_t)p.batch_size, ok ? B : 0, p.batch_size, st, misc, lane);
}

template <class G>
__global__ void __launch_bounds__(SEARCH_BLOCK) k_backprop_puct(SearchParams p) {
    const int lane = threadIdx.x & 31;
    const int tree = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    if (tree >= p.n_trees) return;
    Pending* pd = p.pending + tree;
    const int B = pd->B;
    if (B == 0) return;
    uint4* arena = p.arena + (uint64_t)tree * p.arena_slots;
    const uint2* path = p.path + (uint64_t)tree * p.path_cap;
    const uint32_t info = pd->info[lane];
    double v = 0.0;
    if (lane < B) v = (double)p.values[(size_t)tree * (size_t)p.batch_size + (size_t)lane];
    puct_backup_batch<G>(p, arena, path, B, lane, info, v);
    if (lane == 0) {
        p.ctl[tree].sims_done += (uint32_t)B;
        pd->B = 0;
    }
}

}  // namespace zc
