// Batched MCTS: one warp owns one tree.
//
// What is reproduced (engine/mcts/src/mcts.cpp, digest in SURVEY.md App. A): UCB1 selection
// with statistics FROZEN for a batch of `batch_size` simulations (mcts.cpp:129-149), one new
// node per simulation, evaluation of the batch, then backprop in pending order (:112-127).
//
// How it is restructured for the GPU.  With frozen statistics the `batch_size` descents of a
// batch are not independent walks: every simulation of the batch re-walks the same argmax path
// until it meets the first node that still has untried moves, and a freshly created child has
// Na = 0 => UCT = +inf (mcts.cpp:43) => it wins the argmax at its parent.  So one batch is
//     ONE frozen descent  root -> P                                   (warp-wide argmax per level)
//   + a CHAIN: expand the untried moves of P (one lane per child), then step into P's
//     lowest-index fresh child and expand ITS moves, ... until batch_size leaves exist;
//     a move-less node ends the chain: it is evaluated again by every remaining simulation
//   + one backprop pass: lane l owns path level l and applies the leaves' values to its edge
//     strictly in pending order (same fp64 operation sequence as mcts.cpp:80-100).
// This is exact, not an approximation: tests compare whole-tree hashes with the oracle.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>

#include "tree.cuh"

namespace zc {

constexpr unsigned FULL_MASK = 0xFFFFFFFFu;

struct SearchParams {
    uint4* arena;            // [n_trees][arena_slots]
    TreeCtl* ctl;            // [n_trees]
    uint2* path;             // [n_trees][path_cap]  {node slot, edge index} per level
    Pending* pending;        // [n_trees]  (split-phase only)
    const double* log_tab;   // log_tab[n] = glibc log((double)n), filled on the host (mcts.cpp:44)
    unsigned int* work_counter;
    uint64_t arena_slots;
    uint32_t path_cap;
    int n_trees;
    int simulations;         // per tree, this call
    int batch_size;          // 1..32
    int evaluator;
    int policy;
    float policy_freedom;    // Policy.immediate_value only
    int prior_weight;        // PUCT mode (puct.cuh): priors ~ 1 + prior_weight * move value
    double vloss;            // PUCT mode: virtual loss
    double c;
    uint64_t seed;
    // split-phase leaf packing
    uint16_t* scratch;       // chess: [warp slot][32 lanes][MOVE_SCRATCH] packed moves
    void* planes;
    int plane_dtype;
    const float* values;
};

ZC_D uint4 shfl4(const uint4& v, int src) {
    uint4 r;
    r.x = __shfl_sync(FULL_MASK, v.x, src);
    r.y = __shfl_sync(FULL_MASK, v.y, src);
    r.z = __shfl_sync(FULL_MASK, v.z, src);
    r.w = __shfl_sync(FULL_MASK, v.w, src);
    return r;
}
ZC_D int warp_excl_scan(int v, int lane, int& total) {
    int x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int y = __shfl_up_sync(FULL_MASK, x, d);
        if (lane >= d) x += y;
    }
    total = __shfl_sync(FULL_MASK, x, 31);
    return x - v;
}

// ---- counter-based RNG (splitmix64 finaliser): every random draw is a pure function of a key
ZC_HD uint64_t rng_mix(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// Keyed uniformly random ordering of [0,k), k <= 8: element j of a Fisher-Yates shuffle whose draws are the
// mixed-radix digits of one hash of the key.  Used for Policy.random (policy_functions.py:10-12) on nodes with few moves (Connect
// Four: k <= 7): expanding moves in the order perm(0), perm(1), ... draws each next move uniformly from the untried
// ones, which is what random.choice on the untried list does: pick t takes the (digit t)-th of the remaining elements (chi-square tests on first picks and on pairs,
// tests/test_gpu_parity_bench_sets.py).  The remaining elements live as 4-bit entries of one 32-bit word.
// Wider nodes (chess) order their moves by per-move hash keys instead: ChessGame::random_order.
constexpr int KEYED_PERM_MAX = 8;
ZC_HD uint64_t pick_draw(uint64_t key, int t) { return rng_mix(key ^ (0x9E3779B97F4A7C15ull * (uint64_t)(t + 1))); }
// (out of line: only Policy.random runs it, and inlined it costs the deterministic policies registers)
#ifdef __CUDACC__
__host__ __device__ __noinline__
#else
inline
#endif
int keyed_perm(int k, int j, uint64_t key) {
    if (k <= 1) return 0;
    // one hash per node; the draws are the mixed-radix digits (radices k, k-1, ...) of its high 32 bits: k! <= 8! = 40320, so
    // every draw is uniform up to a relative bias below 2^-16 (Connect Four's 7 moves: 2^-19) -- and the divisions are 32-bit
    uint32_t w = (uint32_t)(rng_mix(key) >> 32);
    uint32_t rest = 0x76543210u;                                   // the moves not yet drawn, four bits each
    int pick = 0;
    for (int t = 0; t <= j; ++t) {
        const uint32_t radix = (uint32_t)(k - t);
        const int d = (int)(w % radix);
        w /= radix;
        pick = (int)((rest >> (4 * d)) & 0xFu);
        const uint32_t low = d ? rest & ((1u << (4 * d)) - 1u) : 0u;
        rest = d < 7 ? low | ((rest >> (4 * d + 4)) << (4 * d)) : low;      // drop entry d
    }
    return pick;
}

// j-th move to be expanded at a node with k moves (mcts.cpp:65-78).  first / last element of the
// untried list keep `untried` an interval; random is a keyed permutation; all are functions of j.
ZC_D int expansion_order(int policy, int k, int j, uint64_t node_key) {
    if (policy == 0) return j;               // ZC_POLICY_FIRST
    if (policy == 1) return k - 1 - j;       // ZC_POLICY_LAST
    return keyed_perm(k, j, node_key);       // ZC_POLICY_RANDOM (games with k <= KEYED_PERM_MAX)
}

// UCB1, mcts.cpp:41-45, with the operation sequence of the reference build
// (log; divide; sqrt; FUSED multiply-add -- see oracle/zc_oracle.c:uct).
ZC_D double uct(double W, int Na, double logN, double c) {
    if (Na == 0) return CUDART_INF;
    const double na = (double)Na;
    const double q = __ddiv_rn(W, na);                       // Qa as stored at mcts.cpp:93
    return __fma_rn(__dsqrt_rn(__ddiv_rn(logN, na)), c, q);
}

// ---------------------------------------------------------------------------------------------
// select + expand one batch for one tree.  Leaves end up one per lane (lane i = pending[i]).
// Returns false if the arena overflowed (tree is then flagged and abandoned).
// ---------------------------------------------------------------------------------------------
template <class G>
struct Leaf {
    uint32_t info;          // LEAF_* encoding, 0 on lanes >= B
    double value;           // built-in evaluators only
    typename G::State st;   // leaf state (for plane packing)
    uint32_t misc;
};

// Per-warp plan of one batch, in shared memory: where each chain level's leaves start in the
// pending list and which leaf becomes the next chain node.  Level g is path level d0+g.
struct WarpPlan {
    int off[34];     // off[g] = index of the first leaf whose parent is chain level g; off[G] = B
    int pleaf[34];   // pleaf[g] = lane of the leaf (child of level g) that is chain level g+1
};

// ---- frozen descent (mcts.cpp:47-63): root -> first node with untried moves (or without moves)
template <class G>
ZC_D bool descend(const SearchParams& p, const uint4* __restrict__ arena, uint2* __restrict__ path, TreeCtl& ctl, int lane,
                  uint32_t& node_out, int& depth_out, uint4& hdr, typename G::State& st) {
    uint32_t node = 0;
    int depth = 0;
    for (;;) {
        const uint4* np = arena + node;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (lane < G::FIRST_SLOTS) v = np[lane];
        hdr = shfl4(v, 0);
        st = G::state_from_lanes(v);
        const int k = (int)hdr_k(hdr), nexp = (int)hdr_nexp(hdr);
        if (nexp < k || k == 0) break;
        const double logN = p.log_tab[hdr.x];
        double best = -CUDART_INF;
        int best_e = 0x7FFFFFFF;
        uint32_t best_child = 0;
        // lane L looks at edges L - (1+SS), +32, ...: the first round comes out of the warp load above, wide
        // (chess) nodes take further rounds; ascending e per lane, so strict > keeps the lowest index
        for (int slot = lane; slot < 1 + G::SS + k; slot += 32) {
            const int e = slot - (1 + G::SS);
            if (e < 0) continue;
            const uint4 ev = slot < G::FIRST_SLOTS ? v : np[slot];
            const double u = uct(edge_W(ev), (int)ev.z, logN, p.c);
            if (u > best) { best = u; best_e = e; best_child = ev.w; }
        }
#pragma unroll (G::kSmallCode ? 1 : 5)      // rolled for chess: its fused search is bound by its instruction footprint (profiles/r2_ab_experiments.md)
        for (int d = 16; d >= 1; d >>= 1) {
            const double ob = __shfl_xor_sync(FULL_MASK, best, d);
            const int oe = __shfl_xor_sync(FULL_MASK, best_e, d);
            const uint32_t oc = __shfl_xor_sync(FULL_MASK, best_child, d);
            if (ob > best || (ob == best && oe < best_e)) { best = ob; best_e = oe; best_child = oc; }
        }
        if (lane == 0) path[depth] = make_uint2(node, (uint32_t)best_e);
        ctl.sum_path_children += (unsigned)k;
        node = best_child;
        ++depth;
        if ((uint32_t)depth + 36u >= p.path_cap) { ctl.status = -4; return false; }
    }
    node_out = node;
    depth_out = depth;
    return true;
}

// lowest move index among the moves order(nexp) .. order(nexp+m-1) expanded at a node, and which j gives it
ZC_D void lowest_fresh(int policy, int k, int nexp, int m, uint64_t nkey, int lane, int& e_star, int& j_star) {
    if (policy == 0) { e_star = nexp; j_star = 0; return; }                   // first: ascending
    if (policy == 1) { e_star = k - nexp - m; j_star = m - 1; return; }       // last: descending
    int e = lane < m ? keyed_perm(k, nexp + lane, nkey) : 0x7FFFFFFF;
    int mn = e;
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) mn = min(mn, __shfl_xor_sync(FULL_MASK, mn, d));
    e_star = mn;
    j_star = __ffs((int)__ballot_sync(FULL_MASK, e == mn)) - 1;
}

// ---------------------------------------------------------------------------------------------
// Chain of expansions, SPINE-FIRST variant for games whose rules are a few instructions
// (G::kCheapSpine): the chain nodes ("spine": P, its lowest fresh child, that child's lowest
// child, ...) are computed by the whole warp redundantly, then ALL batch_size leaves are created
// in ONE lane-parallel step (lane i = pending[i]) instead of one step per chain level.
// ---------------------------------------------------------------------------------------------
template <class G, bool kBuiltinEval, bool KEYED = true>
ZC_D bool expand_spine(const SearchParams& p, typename G::Ctx& gx, uint4* __restrict__ arena, uint2* __restrict__ path,
                       TreeCtl& ctl, int B, int lane, uint32_t P, const uint4& hdr, const typename G::State& Pst,
                       int d0, WarpPlan& wp, int& D_out, Leaf<G>& leaf) {
    const uint64_t tkey = p.seed ^ ((uint64_t)ctl.tree_id << 32);
    const int policy = KEYED ? p.policy : (p.policy & 1);      // KEYED = false: first / last only, the randomised orders are compiled out
    const bool keyed = policy >= 2;                            // only the randomised policies order moves by a node key,
    const bool drawn = p.evaluator == ZC_EVAL_C4_ROLLOUT;      // only the rollout evaluator draws random numbers
    typename G::State Sg = Pst;
    uint32_t Mg = hdr_misc(hdr);
    const int k0 = (int)hdr_k(hdr);
    int kg = k0, nexpg = (int)hdr_nexp(hdr);
    int made = 0, g = 0, spine_lane = -1;
    // latched per lane while the plan is built
    typename G::State myP = Pst;
    uint32_t myPm = Mg;
    int myk = 0, mynexp = 0, myj = 0, myg = 0, my_parent_lane = -1, my_child_nexp = 0, my_spine_e = -1;
    bool is_leaf = false, is_self = false, my_path = false;
    int P_new_nexp = nexpg, P_e = -1, n_self = 0;
    unsigned long long depth_sum = 0;
    uint32_t max_depth = 0;
    for (;;) {
        if (lane == 0) wp.off[g] = made;
        if (kg == 0) {                       // move-less node: select() returns it again and again (mcts.cpp:59,138-141)
            if (lane >= made && lane < B) { is_self = true; myP = Sg; myPm = Mg; myg = g; }
            n_self = B - made;
            depth_sum += (unsigned long long)n_self * (unsigned)(d0 + g);
            max_depth = max(max_depth, (uint32_t)(d0 + g));
            made = B;
            break;
        }
        const int m = min(kg - nexpg, B - made);
        if (lane >= made && lane < made + m) {
            is_leaf = true; myP = Sg; myPm = Mg; myk = kg; mynexp = nexpg; myj = lane - made; myg = g; my_parent_lane = spine_lane;
        }
        if (g == 0) P_new_nexp = nexpg + m;
        else if (lane == spine_lane) my_child_nexp = m;
        depth_sum += (unsigned long long)m * (unsigned)(d0 + g + 1);
        max_depth = max(max_depth, (uint32_t)(d0 + g + 1));
        made += m;
        if (made >= B) break;
        // the node is fully expanded and simulations remain: every fresh child has UCT = +inf, the lowest
        // move index among them wins (mcts.cpp:43,57)
        int e_star, j_star;
        lowest_fresh(policy, kg, nexpg, m, keyed ? rng_mix(tkey ^ G::state_key(Sg, Mg)) : 0ull, lane, e_star, j_star);
        const int next_lane = made - m + j_star;
        if (g == 0) P_e = e_star;
        else if (lane == spine_lane) my_spine_e = e_star;
        if (lane == next_lane) my_path = true;
        if (lane == 0) wp.pleaf[g] = next_lane;
        ctl.sum_path_children += (unsigned)kg;
        uint32_t cm;
        Sg = G::child(Sg, Mg, nullptr, kg, e_star, cm);
        Mg = cm;
        kg = G::count_moves(gx, Sg, Mg);
        nexpg = 0;
        spine_lane = next_lane;
        ++g;
        if ((uint32_t)(d0 + g) + 2u >= p.path_cap) { ctl.status = -4; return false; }
    }
    const int Gc = g + 1;
    if (lane == 0) wp.off[Gc] = B;

    // ---- all leaves at once
    int ei = 0, ck = 0;
    typename G::State cs = myP;
    uint32_t cmisc = myPm;
    double val = 0.0;
    if (is_leaf) {
        ei = expansion_order(policy, myk, mynexp + myj, keyed ? rng_mix(tkey ^ G::state_key(myP, myPm)) : 0ull);
        cs = G::child(myP, myPm, nullptr, myk, ei, cmisc);
        ck = G::count_moves(gx, cs, cmisc);
        if (kBuiltinEval)
            val = G::eval_child(cs, cmisc, ck, p.evaluator, drawn ? rng_mix(tkey ^ G::state_key(cs, cmisc) ^ ((uint64_t)ctl.sims_done << 40) ^ 0x51ull) : 0ull);
    } else if (is_self) {
        if (kBuiltinEval)
            val = G::eval(myP, myPm, p.evaluator, drawn ? rng_mix(tkey ^ G::state_key(myP, myPm) ^ ((uint64_t)ctl.sims_done << 40) ^ ((uint64_t)lane << 8)) : 0ull);
    }
    const int csize = is_leaf ? 1 + G::SS + ck + G::move_slots(ck) : 0;
    int total;
    const int off = warp_excl_scan(csize, lane, total);
    if ((uint64_t)ctl.top + (uint64_t)total > p.arena_slots) { ctl.status = -4; return false; }
    const uint32_t base = ctl.top;
    for (int t = lane; t < total; t += 32) arena[base + t] = make_uint4(0, 0, 0, 0);   // edges start at Na=0, Wa=0, no child
    __syncwarp();
    const uint32_t my_slot = base + (uint32_t)off;
    uint32_t pslot = __shfl_sync(FULL_MASK, my_slot, my_parent_lane < 0 ? 0 : my_parent_lane);
    if (my_parent_lane < 0) pslot = P;
    leaf.info = 0;
    leaf.value = val;
    leaf.st = cs;
    leaf.misc = cmisc;
    if (is_leaf) {
        arena[my_slot] = make_hdr(0, (uint32_t)ck, (uint32_t)my_child_nexp, pslot, (uint32_t)ei, cmisc, (uint32_t)(d0 + myg + 1));
        G::store_state(arena + my_slot + 1, cs);
        G::store_moves(gx, arena + my_slot + 1 + G::SS + ck, ck);
        arena[pslot + 1 + G::SS + ei].w = my_slot;                     // children[move_idx] = child (mcts.cpp:76)
        leaf.info = (uint32_t)(d0 + myg) | ((uint32_t)ei << LEAF_EDGE_SHIFT) | (my_path ? LEAF_PATH : 0u);
        if (my_path) path[d0 + myg + 1] = make_uint2(my_slot, my_spine_e >= 0 ? (uint32_t)my_spine_e : 0xFFFFFFFFu);
    } else if (is_self) {
        leaf.info = LEAF_SELF | (uint32_t)(d0 + myg);
        leaf.st = myP;
        leaf.misc = myPm;
    }
    if (lane == 0) {
        arena[P].y = (uint32_t)k0 | ((uint32_t)P_new_nexp << 16);      // untried.erase (mcts.cpp:72)
        path[d0] = make_uint2(P, P_e >= 0 ? (uint32_t)P_e : 0xFFFFFFFFu);
    }
    ctl.top += (uint32_t)total;
    ctl.nodes += (uint32_t)(B - n_self);
    ctl.reevaluated += (uint32_t)n_self;
    ctl.sum_leaf_depth += depth_sum;
    ctl.max_leaf_depth = max(ctl.max_leaf_depth, max_depth);
    __syncwarp();
    D_out = d0 + Gc - 1;
    return true;
}

// ---------------------------------------------------------------------------------------------
// Chain of expansions, STEPWISE + LAZY variant (chess): one lane-parallel step per chain level, because the
// next chain node's move list must exist before its children can be made.  Children are created as STUBS
// (tree.cuh) -- play the move, evaluate, write header + state -- and a node's move list is generated, by the
// whole warp (one piece per lane), only when the search expands below it: the node the descent ends at, and
// each node the chain steps into.  One or two warp-wide generations per batch instead of one per simulation.
// ---------------------------------------------------------------------------------------------
// ---- stored priors (PUCT mode only, puct.cuh): one float per edge after the node's edges / packed moves
ZC_HD int prior_slots(int k) { return (k + 3) >> 2; }
// slots of a complete node (header, state, edges, packed moves, and the priors in PUCT mode)
template <class G>
ZC_HD int node_slots(int k, bool with_priors) { return 1 + G::SS + k + G::move_slots(k) + (with_priors ? prior_slots(k) : 0); }
template <class G>
ZC_HD float* node_priors(uint4* node, int k) { return reinterpret_cast<float*>(node + 1 + G::SS + k + G::move_slots(k)); }
template <class G>
ZC_HD const float* node_priors(const uint4* node, int k) { return reinterpret_cast<const float*>(node + 1 + G::SS + k + G::move_slots(k)); }
// P(a) = w(a) / sum_b w(b), w(a) = 1 + prior_weight * move_value(a): integers, one IEEE division per edge.  Whole warp;
// the node's state and moves are in place.
template <class G>
ZC_D void write_priors(uint4* node, const typename G::State& st, int k, int prior_weight, int lane) {
    int wsum = 0;
    for (int base = 0; base < k; base += 32) {
        const int a = base + lane;
        int w = a < k ? 1 + prior_weight * G::move_value(node, st, k, a) : 0;
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) w += __shfl_xor_sync(FULL_MASK, w, d);
        wsum += w;
    }
    float* pri = node_priors<G>(node, k);
    for (int a = lane; a < k; a += 32)
        pri[a] = __fdiv_rn((float)(1 + prior_weight * G::move_value(node, st, k, a)), (float)wsum);
    __syncwarp();
}

// Turn the stub at `node` into a complete node; returns its (possibly new) slot in `node` and its k.
#ifndef ZC_AB_MAT_ATTR
#define ZC_AB_MAT_ATTR __noinline__
#endif
template <class G>
__device__ ZC_AB_MAT_ATTR bool materialize(const SearchParams& p, typename G::Ctx& gx, uint4* __restrict__ arena, TreeCtl& ctl, int lane,
                      uint32_t& node, const typename G::State& st, uint32_t misc, int& k_out, bool with_priors = false) {
    const uint4 sh = arena[node];
    const int k = G::moves_warp(gx, st, misc, lane);
    k_out = k;
    if (k == 0) {                                            // move-less node: same footprint as the stub
        if (p.n_trees < 0) zc_layout_pad<ZC_PAD_MAT>();          // never true (zc_common.cuh: code layout)
        if (lane == 0) arena[node].y = 0u;
        __syncwarp();
        return true;
    }
    const uint32_t need = (uint32_t)node_slots<G>(k, with_priors);
    if ((uint64_t)ctl.top + need > p.arena_slots) { ctl.status = -4; return false; }
    const uint32_t base = ctl.top;
#pragma unroll 1
    for (int t = lane; t < k; t += 32) arena[base + 1 + G::SS + t] = make_uint4(0, 0, 0, 0);   // Na = 0, Wa = 0, no child
    G::store_moves_warp(gx, arena + base + 1 + G::SS + k, k, lane);
    if (lane == 0) {
        arena[base] = make_hdr(sh.x, (uint32_t)k, 0, sh.z, hdr_parent_edge(sh), misc, hdr_depth(sh));
        G::store_state(arena + base + 1, st);
        arena[sh.z + 1 + G::SS + hdr_parent_edge(sh)].w = base;        // the parent's children[move_idx] follows the node
    }
    __syncwarp();
    if (with_priors) write_priors<G>(arena + base, st, k, p.prior_weight, lane);
    ctl.top += need;
    node = base;
    return true;
}

template <class G, bool kBuiltinEval, bool KEYED = true>
ZC_D bool expand_lazy(const SearchParams& p, typename G::Ctx& gx, uint4* __restrict__ arena, uint2* __restrict__ path,
                      TreeCtl& ctl, int B, int lane, uint32_t P, const uint4& hdr, const typename G::State& st,
                      int d0, WarpPlan& wp, int& D_out, Leaf<G>& leaf) {
    const uint64_t tkey = p.seed ^ ((uint64_t)ctl.tree_id << 32);
    int Pk = (int)hdr_k(hdr), Pnexp = (int)hdr_nexp(hdr);
    typename G::State Pst = st;
    uint32_t Pmisc = hdr_misc(hdr);
    int D = d0, made = 0, g = 0;
    leaf.info = 0;
    leaf.value = 0.0;
    leaf.st = st;
    leaf.misc = Pmisc;
    if (Pk == (int)K_UNKNOWN) {                              // the descent ended at a leaf of an earlier batch
        if (!materialize<G>(p, gx, arena, ctl, lane, P, Pst, Pmisc, Pk)) return false;
        Pnexp = 0;
    }
    while (made < B) {
        if (lane == 0) wp.off[g] = made;
        if (Pk == 0) {                       // move-less node: select() returns it again and again (:59)
            if (p.n_trees < 0) zc_layout_pad<ZC_PAD_MAIN>();     // never true (zc_common.cuh: code layout)
            if (lane >= made && lane < B) {
                leaf.info = LEAF_SELF | (uint32_t)D;
                leaf.st = Pst;
                leaf.misc = Pmisc;
                if (kBuiltinEval)
                    leaf.value = G::eval(Pst, Pmisc, p.evaluator, rng_mix(tkey ^ G::state_key(Pst, Pmisc) ^ ((uint64_t)ctl.sims_done << 40) ^ ((uint64_t)lane << 8)));
            }
            ctl.reevaluated += (uint32_t)(B - made);
            ctl.sum_leaf_depth += (unsigned long long)(B - made) * (unsigned)D;
            ctl.max_leaf_depth = max(ctl.max_leaf_depth, (uint32_t)D);
            made = B;
            break;
        }
        const int m = min(Pk - Pnexp, B - made);
        const int j = lane - made;
        const bool act = j >= 0 && j < m;
        const int policy = KEYED ? p.policy : (p.policy & 1);     // KEYED = false: first / last only, the randomised orders are compiled out
        const uint64_t nkey = policy >= 2 ? rng_mix(tkey ^ G::state_key(Pst, Pmisc)) : 0ull;      // only the randomised policies are keyed
        int ei = 0x7FFFFFFF;
        typename G::State cs = Pst;
        uint32_t cmisc = 0;
        int ei_iv = 0;
        // random: the node's keyed uniformly random order; immediate_value: the warp replays the node's pick process (uniform
        // among the untried moves within policy_freedom of the best untried capture value)
        if constexpr (KEYED) {
            if (policy == 2) ei_iv = G::random_order(gx, Pk, Pnexp, m, j, nkey, lane);
            else if (policy == 3) ei_iv = G::immediate_value_order(arena + P, Pst, Pk, Pnexp, m, j, p.policy_freedom, nkey, lane);
        }
        if (act) {
            ei = policy >= 2 ? ei_iv : expansion_order(policy, Pk, Pnexp + j, nkey);
            cs = G::child(Pst, Pmisc, arena + P, Pk, ei, cmisc);
        }
        const int total = m * (1 + G::SS);
        if ((uint64_t)ctl.top + (uint64_t)total > p.arena_slots) { ctl.status = -4; return false; }
        uint32_t my_slot = 0;
        if (act) {
            my_slot = ctl.top + (uint32_t)(j * (1 + G::SS));
            arena[my_slot] = make_hdr(0, K_UNKNOWN, 0, P, (uint32_t)ei, cmisc, (uint32_t)(D + 1));
            G::store_state(arena + my_slot + 1, cs);
            arena[P + 1 + G::SS + ei].w = my_slot;                     // children[move_idx] = child (:76)
            leaf.info = (uint32_t)D | ((uint32_t)ei << LEAF_EDGE_SHIFT);
            leaf.st = cs;
            leaf.misc = cmisc;
        }
        if (kBuiltinEval) {
            const double v = G::eval_stubs(gx, cs, cmisc, act, lane);
            if (act) leaf.value = v;
        }
        Pnexp += m;
        if (lane == 0) arena[P].y = (uint32_t)Pk | ((uint32_t)Pnexp << 16);   // untried.erase (:72)
        __syncwarp();
        ctl.top += (uint32_t)total;
        ctl.nodes += (uint32_t)m;
        ctl.sum_leaf_depth += (unsigned long long)m * (unsigned)(D + 1);
        ctl.max_leaf_depth = max(ctl.max_leaf_depth, (uint32_t)(D + 1));
        made += m;
        if (made >= B) break;
        // P is fully expanded and simulations remain: UCT = +inf for every fresh child, the
        // lowest move index among them wins (mcts.cpp:43,57).
        const int min_e = (int)__reduce_min_sync(FULL_MASK, (unsigned)ei);     // ei >= 0
        const unsigned who = __ballot_sync(FULL_MASK, act && ei == min_e);
        const int src = __ffs((int)who) - 1;
        if (lane == src) leaf.info |= LEAF_PATH;
        if (lane == 0) {
            path[D] = make_uint2(P, (uint32_t)min_e);
            wp.pleaf[g] = src;
        }
        ctl.sum_path_children += (unsigned)Pk;
        P = __shfl_sync(FULL_MASK, my_slot, src);
        Pst = G::shfl_state(cs, src);
        Pmisc = __shfl_sync(FULL_MASK, cmisc, src);
        if (!materialize<G>(p, gx, arena, ctl, lane, P, Pst, Pmisc, Pk)) return false;   // its moves are needed now
        Pnexp = 0;
        ++D;
        ++g;
        if ((uint32_t)D + 2u >= p.path_cap) { ctl.status = -4; return false; }
    }
    if (lane == 0) {
        path[D] = make_uint2(P, 0xFFFFFFFFu);
        wp.off[g + 1] = B;
    }
    __syncwarp();
    D_out = D;
    return true;
}

// select + expand one batch for one tree.  Leaves end up one per lane (lane i = pending[i]).
// Returns false if the arena overflowed (tree is then flagged and abandoned).
template <class G, bool kBuiltinEval, bool KEYED = true>
ZC_D bool select_expand(const SearchParams& p, typename G::Ctx& gx, uint4* __restrict__ arena, uint2* __restrict__ path,
                        TreeCtl& ctl, int B, int lane, WarpPlan& wp, int& d0_out, int& D_out, Leaf<G>& leaf) {
    uint32_t P;
    int d0;
    uint4 hdr;
    typename G::State st;
    if (!descend<G>(p, arena, path, ctl, lane, P, d0, hdr, st)) return false;
    d0_out = d0;
    if constexpr (G::kCheapSpine) return expand_spine<G, kBuiltinEval, KEYED>(p, gx, arena, path, ctl, B, lane, P, hdr, st, d0, wp, D_out, leaf);
    else return expand_lazy<G, kBuiltinEval, KEYED>(p, gx, arena, path, ctl, B, lane, P, hdr, st, d0, wp, D_out, leaf);
}

// leaves that are not on the chain: their node and their edge see exactly one backprop
template <class G>
ZC_D void backprop_offchain_leaf(uint4* __restrict__ arena, const uint2* __restrict__ path, int B, int lane, uint32_t info,
                                 double value) {
    if (lane < B && !(info & (LEAF_SELF | LEAF_PATH))) {
        const int li = (int)(info & LEAF_LEVEL_MASK);
        const int e = (int)((info >> LEAF_EDGE_SHIFT) & 0xFFu);
        const uint32_t parent = path[li].x;
        uint4* ep = arena + parent + 1 + G::SS + e;
        const uint32_t child = ep->w;
        uint4 nv;
        edge_set_W(nv, 0.0 - value);      // Wa -= result, from Wa = 0
        nv.z = 1;
        nv.w = child;
        *ep = nv;
        arena[child].x = 1;               // node->N += 1
    }
}

// ---------------------------------------------------------------------------------------------
// backprop of one batch in pending order (mcts.cpp:80-100, 120-124).
// lane i holds leaf i (info, value); lane l then owns path level l and applies the values one by
// one, in order: the same fp64 operation sequence as the reference, for ANY values (network).
// ---------------------------------------------------------------------------------------------
template <class G>
ZC_D void backprop_batch(uint4* __restrict__ arena, const uint2* __restrict__ path, int B, int D, int lane,
                         uint32_t info, double value) {
    backprop_offchain_leaf<G>(arena, path, B, lane, info, value);
    for (int base = 0; base <= D; base += 32) {
        const int l = base + lane;
        const bool on = l <= D;
        uint32_t node = 0, a = 0, N = 0;
        uint4 edge = make_uint4(0, 0, 0, 0);
        uint4* ep = nullptr;
        if (on) {
            const uint2 pe = path[l];
            node = pe.x;
            a = pe.y;
            N = arena[node].x;
            if (l < D) {
                ep = arena + node + 1 + G::SS + a;
                edge = *ep;
            }
        }
        double W = edge_W(edge);
        int Na = (int)edge.z;
        for (int i = 0; i < B; ++i) {
            const uint32_t inf = __shfl_sync(FULL_MASK, info, i);
            const double v = __shfl_sync(FULL_MASK, value, i);
            const int li = (int)(inf & LEAF_LEVEL_MASK);
            const bool self = (inf & LEAF_SELF) != 0, onpath = (inf & LEAF_PATH) != 0;
            const int leaf_depth = self ? li : li + 1;
            if (li >= l || (onpath && li + 1 == l)) ++N;                      // node->N += 1 along the parent chain
            if (l < D && (li > l || (li == l && onpath))) {
                ++Na;                                                         // parent->Na[a] += 1
                W = W - (((leaf_depth - l - 1) & 1) ? -v : v);                // parent->Wa[a] -= result; result = -result
            }
        }
        if (on) {
            arena[node].x = N;
            if (l < D) {
                edge_set_W(edge, W);
                edge.z = (uint32_t)Na;
                *ep = edge;
            }
        }
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// The same backprop in closed form, for evaluators whose values are integers or dyadic rationals
// (every built-in one): all partial sums are exact in fp64, so the order of the additions cannot
// change a bit and the 32-step ordered loop collapses into one warp suffix scan.
//   leaves sorted by level  =>  the leaves below path edge l are a SUFFIX of the pending list
//   (start = wp.off[g+1], g = l - d0) plus the one leaf that is chain node l+1 (wp.pleaf[g]);
//   Wa[l] -= (-1)^(l+1) * sum_{i in suffix} (-1)^{depth_i} v_i  +  v_pleaf
// ---------------------------------------------------------------------------------------------
template <class G>
ZC_D void backprop_exact(uint4* __restrict__ arena, const uint2* __restrict__ path, int B, int d0, int D, int lane,
                         uint32_t info, double value, const WarpPlan& wp) {
    backprop_offchain_leaf<G>(arena, path, B, lane, info, value);
    const int li = (int)(info & LEAF_LEVEL_MASK);
    const int leaf_depth = (info & LEAF_SELF) ? li : li + 1;
    double S = lane < B ? ((leaf_depth & 1) ? -value : value) : 0.0;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {                       // inclusive suffix sums
        const double t = __shfl_down_sync(FULL_MASK, S, d);
        if (lane + d < 32) S += t;
    }
    for (int base = 0; base <= D; base += 32) {
        const int l = base + lane;
        const bool on = l <= D;
        const int g = l - d0;
        int startN = 0, startE = 0, pl = -1;
        if (on && g >= 0) {
            startN = wp.off[g];
            if (l < D) { startE = wp.off[g + 1]; pl = wp.pleaf[g]; }
        }
        double suffix = __shfl_sync(FULL_MASK, S, startE & 31);
        if (startE >= 32) suffix = 0.0;
        const double vp = __shfl_sync(FULL_MASK, value, pl < 0 ? 0 : pl);
        if (on) {
            const uint2 pe = path[l];
            const uint32_t node = pe.x;
            arena[node].x += (uint32_t)(B - startN) + (g > 0 ? 1u : 0u);          // node->N
            if (l < D) {
                uint4* ep = arena + node + 1 + G::SS + pe.y;
                uint4 edge = *ep;
                double W = edge_W(edge);
                W = W - (((l + 1) & 1) ? -suffix : suffix);
                if (g >= 0) W = W - vp;
                edge_set_W(edge, W);
                edge.z += (uint32_t)(B - startE) + (g >= 0 ? 1u : 0u);            // parent->Na[a]
                *ep = edge;
            }
        }
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------
#ifndef ZC_SEARCH_BLOCK
#define ZC_SEARCH_BLOCK 128
#endif
constexpr int SEARCH_BLOCK = ZC_SEARCH_BLOCK;

// Fused persistent search with a built-in evaluator: the whole simulation loop of get_move
// (mcts.cpp:129-149) for every tree.  Warps pull tree indices from a global counter.
// KEYED = false serves the deterministic expansion orders (first / last) without the code of the randomised ones: the hot code
// of a kernel competes for the SM's instruction cache, and what is not there cannot push it apart.
template <class G, bool KEYED = true, int MINB = G::kMinBlocks>
__global__ void __launch_bounds__(SEARCH_BLOCK, MINB) k_search_fused(SearchParams p) {
    __shared__ __align__(16) uint16_t warp_moves[SEARCH_BLOCK / 32][G::WARP_MOVES];   // chess: the warp generator's move list
    __shared__ WarpPlan plans[SEARCH_BLOCK / 32];
    WarpPlan& wp = plans[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
#ifdef ZC_PHASE_SYNC
    // The warps of a block start every batch together (G::kPhaseSync): a warp streams through the whole hot code once per
    // batch, so warps in step fetch each instruction line once per block instead of once per warp.
    if constexpr (G::kPhaseSync) {
        const int n_batches = (p.simulations + p.batch_size - 1) / p.batch_size;
        for (;;) {
            int tree = 0;
            if (lane == 0) tree = (int)atomicAdd(p.work_counter, 1u);
            tree = __shfl_sync(FULL_MASK, tree, 0);
            const bool have = tree < p.n_trees;
            if (!__syncthreads_or(have)) return;
            uint4* arena = p.arena + (uint64_t)(have ? tree : 0) * p.arena_slots;
            uint2* path = p.path + (uint64_t)(have ? tree : 0) * p.path_cap;
            TreeCtl ctl = p.ctl[have ? tree : 0];
            typename G::Ctx gx = G::make_ctx(p, (blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5, lane, warp_moves[threadIdx.x >> 5]);
            bool live = have && ctl.status == 0;
            int done = 0;
            for (int b = 0; b < n_batches; ++b) {
                __syncthreads();
                if (live) {
                    const int B = min(p.batch_size, p.simulations - done);
                    int d0, D;
                    Leaf<G> leaf;
                    if (!select_expand<G, true, KEYED>(p, gx, arena, path, ctl, B, lane, wp, d0, D, leaf)) { live = false; continue; }
                    backprop_exact<G>(arena, path, B, d0, D, lane, leaf.info, leaf.value, wp);
                    done += B;
                    ctl.sims_done += (uint32_t)B;
                    live = ctl.status == 0;
                }
            }
            if (have && lane == 0) p.ctl[tree] = ctl;
        }
    }
#endif
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
            int d0, D;
            Leaf<G> leaf;
            if (!select_expand<G, true, KEYED>(p, gx, arena, path, ctl, B, lane, wp, d0, D, leaf)) break;
            backprop_exact<G>(arena, path, B, d0, D, lane, leaf.info, leaf.value, wp);   // built-in values are exactly summable
            done += B;
            ctl.sims_done += (uint32_t)B;
        }
        if (lane == 0) p.ctl[tree] = ctl;
    }
}

// Split phase 1 (external evaluator): select+expand one batch per tree, pack leaf planes.
template <class G>
__global__ void __launch_bounds__(SEARCH_BLOCK) k_select(SearchParams p, int sims_left) {
    __shared__ __align__(16) uint16_t warp_moves[SEARCH_BLOCK / 32][G::WARP_MOVES];   // chess: the warp generator's move list
    __shared__ WarpPlan plans[SEARCH_BLOCK / 32];
    WarpPlan& wp = plans[threadIdx.x >> 5];
    const int lane = threadIdx.x & 31;
    const int tree = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5);
    if (tree >= p.n_trees) return;
    uint4* arena = p.arena + (uint64_t)tree * p.arena_slots;
    uint2* path = p.path + (uint64_t)tree * p.path_cap;
    TreeCtl ctl = p.ctl[tree];
    Pending* pd = p.pending + tree;
    const int B = ctl.status == 0 ? min(p.batch_size, sims_left) : 0;
    int D = 0, d0 = 0;
    Leaf<G> leaf;
    leaf.info = 0;
    bool ok = B > 0;
    typename G::Ctx gx = G::make_ctx(p, (unsigned)tree, lane, warp_moves[threadIdx.x >> 5]);
    if (ok) ok = select_expand<G, false>(p, gx, arena, path, ctl, B, lane, wp, d0, D, leaf);
    if (lane == 0) {
        pd->B = ok ? B : 0;
        pd->D = D;
        p.ctl[tree] = ctl;
    }
    pd->info[lane] = ok ? leaf.info : 0u;
    G::pack_planes(p.planes, p.plane_dtype, (size_t)tree * (size_t)p.batch_size, ok ? B : 0, p.batch_size, leaf.st, leaf.misc, lane);
}

// Split phase 2: backprop the batch with the caller's values.
template <class G>
__global__ void __launch_bounds__(SEARCH_BLOCK) k_backprop(SearchParams p) {
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
    backprop_batch<G>(arena, path, B, pd->D, lane, info, v);
    if (lane == 0) {
        p.ctl[tree].sims_done += (uint32_t)B;
        pd->B = 0;
    }
}

// Whole-tree hash, one thread per tree, stackless depth-first walk (children in move order).
// Must equal oracle/zc_oracle.c:hash_tree.
template <class G>
__global__ void k_tree_hash(const uint4* __restrict__ arena_all, uint64_t arena_slots, int n_trees,
                            unsigned long long* __restrict__ out) {
    const int tree = blockIdx.x * blockDim.x + threadIdx.x;
    if (tree >= n_trees) return;
    const uint4* arena = arena_all + (uint64_t)tree * arena_slots;
    unsigned long long h = 0x5A17C10E5EEDull;
    uint32_t node = 0;
    uint32_t i = 0;
    bool entering = true;
    for (;;) {
        const uint4 hd = arena[node];
        uint32_t k = hdr_k(hd);
        if (k == K_UNKNOWN) {
            // a stub stands for the leaf the reference holds with its move list and all-zero edges: hash it as that
            const uint32_t kk = (uint32_t)G::count_moves_serial(G::load_state(arena + node + 1), hdr_misc(hd));
            h = mix64(h, ((unsigned long long)hdr_depth(hd) << 48) ^ ((unsigned long long)kk << 32) ^
                             ((unsigned long long)kk << 20) ^ (unsigned long long)hd.x);
            for (uint32_t e = 0; e < kk; ++e) {
                h = mix64(h, (unsigned long long)e << 40);
                h = mix64(h, 0ull);
            }
            if (node == 0) break;
            i = hdr_parent_edge(hd) + 1;
            node = hd.z;
            entering = false;
            continue;
        }
        if (entering) {
            h = mix64(h, ((unsigned long long)hdr_depth(hd) << 48) ^ ((unsigned long long)k << 32) ^
                             ((unsigned long long)(k - hdr_nexp(hd)) << 20) ^ (unsigned long long)hd.x);
            i = 0;
            entering = false;
        }
        bool descended = false;
        while (i < k) {
            const uint4 e = arena[node + 1 + G::SS + i];
            double w = edge_W(e);
            if (w == 0.0) w = 0.0;   // fold -0.0
            h = mix64(h, ((unsigned long long)i << 40) ^ ((unsigned long long)e.z << 1) ^ (e.w ? 1ull : 0ull));
            h = mix64(h, (unsigned long long)__double_as_longlong(w));
            if (e.w) {
                node = e.w;
                entering = true;
                descended = true;
                break;
            }
            ++i;
        }
        if (descended) continue;
        if (node == 0) break;
        i = hdr_parent_edge(hd) + 1;
        node = hd.z;
    }
    out[tree] = h;
}

}  // namespace zc
