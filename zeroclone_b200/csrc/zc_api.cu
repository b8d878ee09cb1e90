// libzc_b200.so -- C-ABI (include/zc_b200.h) over the CUDA search kernels.
// Host side: handle bookkeeping, launches, copies.  No CPU implementation of the search exists.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/zc_b200.h"
#include "c4_game.cuh"
#include "chess_game.cuh"
#include "search.cuh"
#include "puct.cuh"

namespace zc {
namespace c4 {
// CPython 3.12 set order; replaced at import by the Python host
uint32_t h_order[128];
static const uint8_t kDefaultOrder[128][8] = {
#include "c4_order_py312.inc"
};
}  // namespace c4
}  // namespace zc

using namespace zc;

static thread_local std::string g_err;
static std::string tower_fault_note();   // tower_api.inl: names a tower kernel's timed-out wait, if any
static int fail(int code, const std::string& msg) {
    g_err = code == ZC_ECUDA ? msg + tower_fault_note() : msg;
    return code;
}
#define CUDA_TRY(expr)                                                                         \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess)                                                                 \
            return fail(ZC_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));         \
    } while (0)

struct zc_search {
    int game = 0, device = 0, max_trees = 0, max_sims = 0, n_trees = 0;
    uint64_t arena_slots = 0;
    uint32_t path_cap = 0;
    uint4* arena = nullptr;
    TreeCtl* ctl = nullptr;
    uint2* path = nullptr;
    Pending* pending = nullptr;
    double* log_tab = nullptr;
    unsigned int* work_counter = nullptr;
    uint16_t* scratch = nullptr;        // chess move staging, one buffer per lane
    void* roots_dev = nullptr;          // staging for host roots
    zc_root_result* res_dev = nullptr;
    zc_root_result* res_dev2[2] = {nullptr, nullptr};   // results_begin/end: double-buffered device results
    zc_root_result* res_host = nullptr;  // pinned staging of the per-tree results
    cudaStream_t copy_stream = nullptr;  // results_begin: the D2H copy runs beside the next step's kernels
    cudaEvent_t ev_res_ready = nullptr, ev_copy_done[2] = {nullptr, nullptr};
    int res_flip = 0, res_pending = -1, res_pending_n = 0;
    int32_t* visits_dev = nullptr;
    double* wsum_dev = nullptr;
    zc_chess_move* moves_dev = nullptr;
    int res_stride = 0;
    unsigned long long* hash_dev = nullptr;
    int32_t* adv_res_dev = nullptr;     // zc_search_advance outputs
    zc_chess_move* adv_mv_dev = nullptr;
    int* adv_kmp_dev = nullptr;
    int adv_kmp_cap = 0;
    int32_t* adv_err_dev = nullptr;     // {flags, first offending tree}
    int64_t bytes = 0;
    int64_t launches = 0;
    int fused_grid = 0, fused_grid_det = 0;      // grids of k_search_fused<G, true> / <G, false>
    int order_version = 0;
    // split-phase state
    int sp_left = 0, sp_batch = 0, sp_policy = 0, sp_selected = 0;
    double sp_c = 1.4;
    uint64_t sp_seed = 0;
    double policy_freedom = 0.0;   // Policy.immediate_value
    // selection rule (zc_search_set_mode): the reference's UCB1, or PUCT with stored priors and virtual loss (puct.cuh)
    int select_mode = ZC_SELECT_UCB1, prior_weight = 0, fused_grid_puct = 0;
    double vloss = 1.0;
};

// ------------------------------------------------------------------------------- C4 move order
static bool g_order_ready = false;
static int g_order_version = 1;
static void pack_order(const uint8_t* table) {
    for (int m = 0; m < 128; ++m) {
        uint32_t w = 0;
        for (int i = 0; i < 7 && table[m * 8 + i] != 255; ++i) w |= (uint32_t)(table[m * 8 + i] & 0xF) << (4 * i);
        c4::h_order[m] = w;
    }
    g_order_ready = true;
}
static void ensure_order() {
    if (!g_order_ready) pack_order(&c4::kDefaultOrder[0][0]);
}
static int upload_order() {
    ensure_order();
    CUDA_TRY(cudaMemcpyToSymbol(c4::d_order, c4::h_order, sizeof(c4::h_order)));
    return ZC_OK;
}

extern "C" int zc_c4_set_move_order(const uint8_t* table) {
    if (!table) return fail(ZC_EINVAL, "table is NULL");
    for (int m = 0; m < 128; ++m) {   // every row must be a permutation of the mask's columns
        int seen = 0, n = 0;
        for (; n < 7 && table[m * 8 + n] != 255; ++n) {
            if (table[m * 8 + n] > 6) return fail(ZC_EINVAL, "move order: column out of range");
            seen |= 1 << table[m * 8 + n];
        }
        if (seen != m || n != __builtin_popcount(m)) return fail(ZC_EINVAL, "move order: row is not a permutation of its mask");
    }
    pack_order(table);
    ++g_order_version;   // handles re-upload to their device at the next set_roots
    return ZC_OK;
}
extern "C" int zc_c4_get_move_order(uint8_t* table) {
    if (!table) return fail(ZC_EINVAL, "table is NULL");
    ensure_order();
    memset(table, 255, 128 * 8);
    for (int m = 0; m < 128; ++m)
        for (int i = 0; i < __builtin_popcount(m); ++i) table[m * 8 + i] = (uint8_t)((c4::h_order[m] >> (4 * i)) & 0xF);
    return ZC_OK;
}

extern "C" int zc_abi_version(void) { return ZC_ABI_VERSION; }
extern "C" const char* zc_last_error(void) { return g_err.c_str(); }
extern "C" int zc_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

// ------------------------------------------------------------------------------- kernels local to the API
__global__ void k_set_roots_c4(const zc_c4_state* __restrict__ roots, uint4* __restrict__ arena_all, uint64_t arena_slots,
                               TreeCtl* __restrict__ ctl, Pending* __restrict__ pending, int n, int with_priors) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const zc_c4_state r = roots[t];
    c4::State s;
    s.cur = r.turn == 0 ? r.x : r.o;
    s.opp = r.turn == 0 ? r.o : r.x;
    const int k = c4::n_moves(s);
    uint4* arena = arena_all + (uint64_t)t * arena_slots;
    arena[0] = make_hdr(0, (uint32_t)k, 0, 0, 0, 0, 0);
    C4Game::store_state(arena + 1, s);
    for (int i = 0; i < k; ++i) arena[2 + i] = make_uint4(0, 0, 0, 0);
    if (with_priors) {                                   // PUCT mode: uniform priors (every Connect Four move has value 0)
        for (int i = 0; i < prior_slots(k); ++i) arena[2 + k + i] = make_uint4(0, 0, 0, 0);
        float* pri = node_priors<C4Game>(arena, k);
        for (int i = 0; i < k; ++i) pri[i] = __fdiv_rn(1.0f, (float)k);
    }
    TreeCtl c;
    memset(&c, 0, sizeof c);
    c.top = (uint32_t)node_slots<C4Game>(k, with_priors != 0);
    c.nodes = 1;
    c.root_turn = (uint32_t)r.turn;
    c.tree_id = (uint32_t)t;
    ctl[t] = c;
    pending[t].B = 0;
}

__global__ void k_results_c4(const uint4* __restrict__ arena_all, uint64_t arena_slots, const TreeCtl* __restrict__ ctl,
                             int n, zc_root_result* __restrict__ res, int32_t* __restrict__ visits,
                             double* __restrict__ wsum, zc_chess_move* __restrict__ moves, int stride) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const uint4* arena = arena_all + (uint64_t)t * arena_slots;
    const uint4 hd = arena[0];
    const int k = (int)hdr_k(hd);
    const c4::State s = C4Game::load_state(arena + 1);
    const int mask = c4::legal_mask(s);
    int best = -1, best_n = -1;
    for (int i = 0; i < k; ++i) {
        const uint4 e = arena[2 + i];
        if (e.w && (int)e.z > best_n) { best_n = (int)e.z; best = i; }      // mcts.cpp:150-155
        if (i < stride) {
            if (visits) visits[(size_t)t * stride + i] = (int)e.z;
            if (wsum) wsum[(size_t)t * stride + i] = edge_W(e);
            if (moves) {
                zc_chess_move mv;
                mv.fr = (uint8_t)c4::move_col(mask, i);
                mv.fc = mv.tr = mv.tc = 0;
                mv.value = 0.f;
                moves[(size_t)t * stride + i] = mv;
            }
        }
    }
    const TreeCtl c = ctl[t];
    zc_root_result r;
    memset(&r, 0, sizeof r);
    r.n_moves = k;
    r.best = best;
    r.root_visits = (int)hd.x;
    r.status = c.status;
    r.best_move[0] = best >= 0 ? (uint8_t)c4::move_col(mask, best) : 255;
    r.nodes = (int)c.nodes;
    r.sum_leaf_depth = (int64_t)c.sum_leaf_depth;
    r.max_leaf_depth = (int)c.max_leaf_depth;
    r.reevaluated_leaves = (int)c.reevaluated;
    res[t] = r;
}


__global__ void k_set_roots_chess(const zc_chess_state* __restrict__ roots, uint4* __restrict__ arena_all,
                                  uint64_t arena_slots, TreeCtl* __restrict__ ctl, Pending* __restrict__ pending,
                                  uint16_t* __restrict__ scratch, int n, int with_priors, int prior_weight) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const zc_chess_state r = roots[t];
    chess::Board b = {0, 0, 0, 0};
    for (int i = 0; i < 64; ++i) chess::put_piece(b, i, chess::code_of_char(r.board[i]));
    const uint32_t misc = (r.turn ? chess::MISC_TURN : 0u) | (r.w_ck ? chess::MISC_WCK : 0u) | (r.w_cq ? chess::MISC_WCQ : 0u) |
                          (r.b_ck ? chess::MISC_BCK : 0u) | (r.b_cq ? chess::MISC_BCQ : 0u);
    ChessGame::Ctx gx;
    gx.moves = scratch + (size_t)t * 32 * ChessGame::MOVE_SCRATCH;
    gx.stride = 1;
    gx.wmoves = nullptr;      // the warp generator is not used here
    const int k = ChessGame::count_moves(gx, b, misc);
    uint4* arena = arena_all + (uint64_t)t * arena_slots;
    arena[0] = make_hdr(0, (uint32_t)k, 0, 0, 0, misc, 0);
    ChessGame::store_state(arena + 1, b);
    for (int i = 0; i < k; ++i) arena[3 + i] = make_uint4(0, 0, 0, 0);
    ChessGame::store_moves(gx, arena + 3 + k, k);
    if (with_priors) {                                   // PUCT mode: P(a) = (1 + prior_weight * capture value) / sum
        for (int i = 0; i < prior_slots(k); ++i) arena[3 + k + ChessGame::move_slots(k) + i] = make_uint4(0, 0, 0, 0);
        int wsum = 0;
        for (int i = 0; i < k; ++i) wsum += 1 + prior_weight * ChessGame::move_value(arena, b, k, i);
        float* pri = node_priors<ChessGame>(arena, k);
        for (int i = 0; i < k; ++i) pri[i] = __fdiv_rn((float)(1 + prior_weight * ChessGame::move_value(arena, b, k, i)), (float)wsum);
    }
    TreeCtl c;
    memset(&c, 0, sizeof c);
    c.top = (uint32_t)node_slots<ChessGame>(k, with_priors != 0);
    c.nodes = 1;
    c.root_turn = r.turn;
    c.tree_id = (uint32_t)t;
    ctl[t] = c;
    pending[t].B = 0;
}

__device__ __host__ inline zc_chess_move decode_move(const chess::Board& b, uint16_t m) {
    zc_chess_move mv;
    const int f = chess::move_from(m), t = chess::move_to(m);
    mv.fr = (uint8_t)(f >> 3);
    mv.fc = (uint8_t)(f & 7);
    mv.tr = (uint8_t)(t >> 3);
    mv.tc = (uint8_t)(t & 7);
    mv.value = (float)chess::capture_value(chess::piece_at(b, t));
    return mv;
}

__global__ void k_results_chess(const uint4* __restrict__ arena_all, uint64_t arena_slots, const TreeCtl* __restrict__ ctl,
                                int n, zc_root_result* __restrict__ res, int32_t* __restrict__ visits,
                                double* __restrict__ wsum, zc_chess_move* __restrict__ moves, int stride) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const uint4* arena = arena_all + (uint64_t)t * arena_slots;
    const uint4 hd = arena[0];
    const int k = (int)hdr_k(hd);
    const chess::Board b = ChessGame::load_state(arena + 1);
    int best = -1, best_n = -1;
    for (int i = 0; i < k; ++i) {
        const uint4 e = arena[3 + i];
        if (e.w && (int)e.z > best_n) { best_n = (int)e.z; best = i; }      // mcts.cpp:150-155
        if (i < stride) {
            if (visits) visits[(size_t)t * stride + i] = (int)e.z;
            if (wsum) wsum[(size_t)t * stride + i] = edge_W(e);
            if (moves) moves[(size_t)t * stride + i] = decode_move(b, ChessGame::move_at(arena, k, i));
        }
    }
    const TreeCtl c = ctl[t];
    zc_root_result r;
    memset(&r, 0, sizeof r);
    r.n_moves = k;
    r.best = best;
    r.root_visits = (int)hd.x;
    r.status = c.status;
    if (best >= 0) {
        const zc_chess_move mv = decode_move(b, ChessGame::move_at(arena, k, best));
        r.best_move[0] = mv.fr; r.best_move[1] = mv.fc; r.best_move[2] = mv.tr; r.best_move[3] = mv.tc;
        r.best_move_value = mv.value;
    } else {
        r.best_move[0] = r.best_move[1] = r.best_move[2] = r.best_move[3] = 255;
    }
    r.nodes = (int)c.nodes;
    r.sum_leaf_depth = (int64_t)c.sum_leaf_depth;
    r.max_leaf_depth = (int)c.max_leaf_depth;
    r.reevaluated_leaves = (int)c.reevaluated;
    res[t] = r;
}

#define ZC_DISPATCH(game, ...)                 \
    do {                                       \
        if ((game) == ZC_GAME_C4) {            \
            using G = C4Game;                  \
            __VA_ARGS__;                       \
        } else {                               \
            using G = ChessGame;               \
            __VA_ARGS__;                       \
        }                                      \
    } while (0)

// ------------------------------------------------------------------------------- handle
static int check_handle(const zc_search* h) {
    if (!h) return fail(ZC_EINVAL, "handle is NULL");
    return ZC_OK;
}

extern "C" int zc_search_create(int game, int device, int max_trees, int max_sims, int64_t arena_slots_per_tree,
                                zc_search** out) {
    if (!out) return fail(ZC_EINVAL, "out is NULL");
    *out = nullptr;
    if (game != ZC_GAME_C4 && game != ZC_GAME_CHESS) return fail(ZC_EINVAL, "unknown game");
    if (max_trees < 1 || max_sims < 1) return fail(ZC_EINVAL, "max_trees and max_sims must be >= 1");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(ZC_ENODEVICE, "no CUDA device: libzc_b200 has no CPU path");
    }
    if (device < 0 || device >= ndev) return fail(ZC_EINVAL, "device out of range");
    CUDA_TRY(cudaSetDevice(device));
    zc_search* h = new zc_search();
    h->game = game;
    h->device = device;
    h->max_trees = max_trees;
    h->max_sims = max_sims;
    // chess: stubs are 3 slots and about one node in thirty becomes a complete node (~37 slots): 4.9 slots per
    // node measured on configs[4]; 24 leaves a factor 5, and an arena that does overflow is reported, not truncated
    // (Connect Four: header + state + 7 edges + 2 prior slots, the exact worst case in either selection mode)
    const int per_node = game == ZC_GAME_C4 ? node_slots<C4Game>(7, true) : 24;
    h->arena_slots = arena_slots_per_tree > 0 ? (uint64_t)arena_slots_per_tree
                                              : (uint64_t)(max_sims + 1) * per_node + (game == ZC_GAME_C4 ? 0 : 320);   // + a maximal chess root
    h->arena_slots = (h->arena_slots + 1) & ~1ull;   // keep every tree's arena 32-byte aligned
    h->path_cap = std::max((uint32_t)max_sims + 40u, PUCT_PATH_CAP);   // PUCT keeps 64 path entries per simulation of a batch
    auto alloc = [&](void** p, size_t bytes) -> cudaError_t {
        h->bytes += (int64_t)bytes;
        return cudaMalloc(p, bytes);
    };
    const size_t arena_bytes = ((size_t)max_trees * h->arena_slots + 64) * sizeof(uint4);   // +64: speculative node loads
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = alloc((void**)&h->arena, arena_bytes);
    if (e == cudaSuccess) e = alloc((void**)&h->ctl, sizeof(TreeCtl) * max_trees);
    if (e == cudaSuccess) e = alloc((void**)&h->path, sizeof(uint2) * (size_t)max_trees * h->path_cap);
    if (e == cudaSuccess) e = alloc((void**)&h->pending, sizeof(Pending) * max_trees);
    if (e == cudaSuccess) e = alloc((void**)&h->log_tab, sizeof(double) * ((size_t)max_sims + 4));
    if (e == cudaSuccess) e = alloc((void**)&h->work_counter, sizeof(unsigned int));
    if (e == cudaSuccess) e = alloc((void**)&h->roots_dev, sizeof(zc_chess_state) * max_trees);
    if (e == cudaSuccess) e = alloc((void**)&h->res_dev, sizeof(zc_root_result) * max_trees);
    if (e == cudaSuccess) e = cudaMallocHost((void**)&h->res_host, sizeof(zc_root_result) * max_trees);
    if (e == cudaSuccess) e = alloc((void**)&h->hash_dev, sizeof(unsigned long long) * max_trees);
    if (e != cudaSuccess) {
        std::string msg = std::string("cudaMalloc: ") + cudaGetErrorString(e);
        cudaGetLastError();      // an allocation failure is not sticky, but it would be reported by the next launch check
        zc_search_destroy(h);
        return fail(ZC_ECUDA, msg);
    }
    // log table from the HOST libm: the reference calls glibc's log (mcts.cpp:44), which is not
    // correctly rounded, so the device must use the very same values to stay bit-exact.
    std::vector<double> lt((size_t)max_sims + 4);
    lt[0] = -INFINITY;
    for (size_t i = 1; i < lt.size(); ++i) {
        volatile double x = (double)i;
        lt[i] = std::log(x);
    }
    int occ = 0, occ_det = 0, sms = 0;      // resident blocks per SM of the two instantiations (randomised / deterministic orders)
    e = cudaMemcpy(h->log_tab, lt.data(), lt.size() * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess)
        e = game == ZC_GAME_C4 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_search_fused<C4Game, true>, SEARCH_BLOCK, 0)
                               : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_search_fused<ChessGame, true>, SEARCH_BLOCK, 0);
    if (e == cudaSuccess)
        e = game == ZC_GAME_C4 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_det, k_search_fused<C4Game, false>, SEARCH_BLOCK, 0)
                               : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_det, k_search_fused<ChessGame, false>, SEARCH_BLOCK, 0);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->ev_res_ready, cudaEventDisableTiming);
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
        e = cudaEventCreateWithFlags(&h->ev_copy_done[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = alloc((void**)&h->res_dev2[i], sizeof(zc_root_result) * max_trees);
    }
    // struct padding of zc_root_result is never written by the readout kernels: give it a defined value once
    if (e == cudaSuccess) e = cudaMemset(h->res_dev, 0, sizeof(zc_root_result) * max_trees);
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) e = cudaMemset(h->res_dev2[i], 0, sizeof(zc_root_result) * max_trees);
    if (e != cudaSuccess) {     // the handle is not handed out: free what it holds
        std::string msg = std::string("zc_search_create: ") + cudaGetErrorString(e);
        cudaGetLastError();
        zc_search_destroy(h);
        return fail(ZC_ECUDA, msg);
    }
    h->fused_grid = occ * sms;
    h->fused_grid_det = occ_det * sms;
    if (game == ZC_GAME_CHESS) {
        const size_t warps = (size_t)std::max(max_trees, std::max(h->fused_grid, h->fused_grid_det) * (SEARCH_BLOCK / 32));
        const size_t sb = warps * 32 * ChessGame::MOVE_SCRATCH * sizeof(uint16_t);
        h->bytes += (int64_t)sb;
        cudaError_t se = cudaMalloc((void**)&h->scratch, sb);
        if (se != cudaSuccess) {
            cudaGetLastError();
            zc_search_destroy(h);
            return fail(ZC_ECUDA, std::string("cudaMalloc(scratch): ") + cudaGetErrorString(se));
        }
    }
    *out = h;
    return ZC_OK;
}

extern "C" int zc_search_destroy(zc_search* h) {
    if (!h) return ZC_OK;
    cudaSetDevice(h->device);
    if (h->copy_stream) cudaStreamSynchronize(h->copy_stream);      // a readout copy may still be in flight
    cudaFree(h->arena);
    cudaFree(h->ctl);
    cudaFree(h->path);
    cudaFree(h->pending);
    cudaFree(h->log_tab);
    cudaFree(h->work_counter);
    cudaFree(h->scratch);
    cudaFree(h->roots_dev);
    cudaFree(h->res_dev);
    cudaFree(h->res_dev2[0]);
    cudaFree(h->res_dev2[1]);
    cudaFreeHost(h->res_host);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    if (h->ev_res_ready) cudaEventDestroy(h->ev_res_ready);
    if (h->ev_copy_done[0]) cudaEventDestroy(h->ev_copy_done[0]);
    if (h->ev_copy_done[1]) cudaEventDestroy(h->ev_copy_done[1]);
    cudaFree(h->adv_err_dev);
    cudaFree(h->visits_dev);
    cudaFree(h->wsum_dev);
    cudaFree(h->moves_dev);
    cudaFree(h->hash_dev);
    cudaFree(h->adv_res_dev);
    cudaFree(h->adv_mv_dev);
    cudaFree(h->adv_kmp_dev);
    delete h;
    return ZC_OK;
}

extern "C" int64_t zc_search_device_bytes(const zc_search* h) { return h ? h->bytes : 0; }

static int set_roots_common(zc_search* h, const void* dev_states, int n, cudaStream_t st) {
    h->n_trees = n;
    h->sp_left = 0;
    if (h->order_version != g_order_version) {
        if (int rc = upload_order()) return rc;
        h->order_version = g_order_version;
    }
    if (h->game == ZC_GAME_C4)
        k_set_roots_c4<<<(n + 127) / 128, 128, 0, st>>>((const zc_c4_state*)dev_states, h->arena, h->arena_slots, h->ctl,
                                                      h->pending, n, h->select_mode == ZC_SELECT_PUCT);
    else
        k_set_roots_chess<<<(n + 63) / 64, 64, 0, st>>>((const zc_chess_state*)dev_states, h->arena, h->arena_slots, h->ctl,
                                                      h->pending, h->scratch, n, h->select_mode == ZC_SELECT_PUCT, h->prior_weight);
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    return ZC_OK;
}

extern "C" int zc_search_set_roots(zc_search* h, const void* host_states, int n, void* stream) {
    if (int rc = check_handle(h)) return rc;
    if (!host_states || n < 1 || n > h->max_trees) return fail(ZC_EINVAL, "set_roots: bad states or n");
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t sz = h->game == ZC_GAME_C4 ? sizeof(zc_c4_state) : sizeof(zc_chess_state);
    CUDA_TRY(cudaMemcpyAsync(h->roots_dev, host_states, sz * n, cudaMemcpyHostToDevice, st));
    // stream-ordered, no host synchronisation: cudaMemcpyAsync has staged a pageable source before it returns;
    // a pinned source must stay unchanged until the stream has passed this point
    return set_roots_common(h, h->roots_dev, n, st);
}

extern "C" int zc_search_set_roots_dev(zc_search* h, const void* dev_states, int n, void* stream) {
    if (int rc = check_handle(h)) return rc;
    if (!dev_states || n < 1 || n > h->max_trees) return fail(ZC_EINVAL, "set_roots_dev: bad states or n");
    CUDA_TRY(cudaSetDevice(h->device));
    return set_roots_common(h, dev_states, n, (cudaStream_t)stream);
}

static SearchParams make_params(zc_search* h, int sims, double c, int batch, int evaluator, int policy, uint64_t seed) {
    SearchParams p;
    memset(&p, 0, sizeof p);
    p.arena = h->arena;
    p.ctl = h->ctl;
    p.path = h->path;
    p.pending = h->pending;
    p.log_tab = h->log_tab;
    p.work_counter = h->work_counter;
    p.arena_slots = h->arena_slots;
    p.path_cap = h->path_cap;
    p.n_trees = h->n_trees;
    p.simulations = sims;
    p.batch_size = batch;
    p.evaluator = evaluator;
    // Connect Four moves all carry value 0 (c4_backend.py:50): immediate_value picks uniformly = random
    p.policy = (policy == ZC_POLICY_IMMEDIATE_VALUE && h->game == ZC_GAME_C4) ? ZC_POLICY_RANDOM : policy;
    p.policy_freedom = (float)h->policy_freedom;
    p.prior_weight = h->prior_weight;
    p.vloss = h->vloss;
    p.c = c;
    p.seed = seed;
    p.scratch = h->scratch;
    return p;
}

static int check_search_args(zc_search* h, int sims, int batch, int policy) {
    if (h->n_trees < 1) return fail(ZC_ESTATE, "no roots set");
    if (sims < 0 || sims > h->max_sims) return fail(ZC_EINVAL, "simulations exceeds max_sims of the handle");
    if (batch < 1 || batch > 32) return fail(ZC_EINVAL, "batch_size must be in 1..32");
    if (policy != ZC_POLICY_FIRST && policy != ZC_POLICY_LAST && policy != ZC_POLICY_RANDOM && policy != ZC_POLICY_IMMEDIATE_VALUE)
        return fail(ZC_EINVAL, "unsupported policy");
    return ZC_OK;
}

extern "C" int zc_search_run(zc_search* h, int simulations, double c, int batch_size, int evaluator, int policy,
                             uint64_t seed, void* stream) {
    if (int rc = check_handle(h)) return rc;
    if (int rc = check_search_args(h, simulations, batch_size, policy)) return rc;
    const bool ev_ok = h->game == ZC_GAME_C4 ? (evaluator == ZC_EVAL_C4_TERMINAL || evaluator == ZC_EVAL_C4_POSITIONAL || evaluator == ZC_EVAL_C4_ROLLOUT)
                                             : evaluator == ZC_EVAL_CHESS_CRUDE;
    if (!ev_ok) return fail(ZC_EINVAL, "evaluator is not a built-in evaluator of this game");
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    SearchParams p = make_params(h, simulations, c, batch_size, evaluator, policy, seed);
    CUDA_TRY(cudaMemsetAsync(h->work_counter, 0, sizeof(unsigned int), st));
    const int blocks_needed = (h->n_trees * 32 + SEARCH_BLOCK - 1) / SEARCH_BLOCK;
    if (h->select_mode == ZC_SELECT_PUCT) {
        if (evaluator == ZC_EVAL_C4_ROLLOUT) return fail(ZC_EINVAL, "PUCT mode: the rollout evaluator is not supported");
        if (!h->fused_grid_puct) {
            int occ = 0, sms = 0;
            ZC_DISPATCH(h->game, CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_search_fused_puct<G>, SEARCH_BLOCK, 0)));
            CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device));
            h->fused_grid_puct = occ * sms;
        }
        const int grid = blocks_needed < h->fused_grid_puct ? blocks_needed : h->fused_grid_puct;
        ZC_DISPATCH(h->game, k_search_fused_puct<G><<<grid, SEARCH_BLOCK, 0, st>>>(p));
        h->launches++;
        CUDA_TRY(cudaGetLastError());
        return ZC_OK;
    }
    if (policy >= ZC_POLICY_RANDOM) {
        const int grid = blocks_needed < h->fused_grid ? blocks_needed : h->fused_grid;
        ZC_DISPATCH(h->game, (k_search_fused<G, true><<<grid, SEARCH_BLOCK, 0, st>>>(p)));
    } else {                                 // first / last: the instantiation without the randomised orders' code
        const int grid = blocks_needed < h->fused_grid_det ? blocks_needed : h->fused_grid_det;
        ZC_DISPATCH(h->game, (k_search_fused<G, false><<<grid, SEARCH_BLOCK, 0, st>>>(p)));
    }
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    return ZC_OK;
}

extern "C" int zc_search_set_policy_freedom(zc_search* h, double policy_freedom) {
    if (int rc = check_handle(h)) return rc;
    if (!(policy_freedom >= 0.0)) return fail(ZC_EINVAL, "policy_freedom must be >= 0");
    h->policy_freedom = policy_freedom;
    return ZC_OK;
}

// The selection rule of the handle.  Call before zc_search_set_roots*: PUCT roots carry their priors.
extern "C" int zc_search_set_mode(zc_search* h, int select_mode, double virtual_loss, int prior_weight) {
    if (int rc = check_handle(h)) return rc;
    if (select_mode != ZC_SELECT_UCB1 && select_mode != ZC_SELECT_PUCT) return fail(ZC_EINVAL, "unknown selection mode");
    if (!(virtual_loss >= 0.0) || prior_weight < 0 || prior_weight > 1000) return fail(ZC_EINVAL, "bad virtual_loss or prior_weight");
    h->select_mode = select_mode;
    h->vloss = virtual_loss;
    h->prior_weight = prior_weight;
    h->n_trees = 0;          // roots have to be set again in the new mode
    h->sp_left = 0;
    return ZC_OK;
}

template <class G>
__global__ void k_set_root_priors(uint4* __restrict__ arena_all, uint64_t arena_slots, int n, const float* __restrict__ priors, int stride) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    uint4* root = arena_all + (uint64_t)t * arena_slots;
    const int k = (int)hdr_k(root[0]);
    float* pri = node_priors<G>(root, k);
    for (int a = 0; a < k && a < stride; ++a) pri[a] = priors[(size_t)t * stride + a];
}

extern "C" int zc_search_set_root_priors(zc_search* h, const float* host_priors, int stride, void* stream) {
    if (int rc = check_handle(h)) return rc;
    if (h->select_mode != ZC_SELECT_PUCT) return fail(ZC_ESTATE, "set_root_priors: the handle is not in PUCT mode");
    if (h->n_trees < 1) return fail(ZC_ESTATE, "no roots set");
    if (!host_priors || stride < 1 || stride > ZC_MAX_MOVES) return fail(ZC_EINVAL, "set_root_priors: bad priors or stride");
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    float* dev = nullptr;
    const size_t bytes = sizeof(float) * (size_t)h->n_trees * stride;
    CUDA_TRY(cudaMalloc((void**)&dev, bytes));
    cudaError_t e = cudaMemcpyAsync(dev, host_priors, bytes, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) {
        ZC_DISPATCH(h->game, k_set_root_priors<G><<<(h->n_trees + 127) / 128, 128, 0, st>>>(h->arena, h->arena_slots, h->n_trees, dev, stride));
        h->launches++;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(dev);
    if (e != cudaSuccess) return fail(ZC_ECUDA, std::string("set_root_priors: ") + cudaGetErrorString(e));
    return ZC_OK;
}

extern "C" int zc_search_begin(zc_search* h, int simulations, double c, int batch_size, int policy, uint64_t seed) {
    if (int rc = check_handle(h)) return rc;
    if (int rc = check_search_args(h, simulations, batch_size, policy)) return rc;
    h->sp_left = simulations;
    h->sp_batch = batch_size;
    h->sp_policy = policy;
    h->sp_c = c;
    h->sp_seed = seed;
    h->sp_selected = 0;
    return ZC_OK;
}
extern "C" int zc_search_pending(const zc_search* h) { return h ? h->sp_left : 0; }

extern "C" int zc_search_select(zc_search* h, void* dev_planes, int plane_dtype, void* stream) {
    if (int rc = check_handle(h)) return rc;
    if (h->sp_left <= 0) return fail(ZC_ESTATE, "select: no simulations pending (call zc_search_begin)");
    if (h->sp_selected) return fail(ZC_ESTATE, "select: previous batch not backpropagated");
    if (!dev_planes || plane_dtype < 0 || plane_dtype > 2) return fail(ZC_EINVAL, "select: bad planes or dtype");
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    SearchParams p = make_params(h, h->sp_left, h->sp_c, h->sp_batch, ZC_EVAL_EXTERNAL, h->sp_policy, h->sp_seed);
    p.planes = dev_planes;
    p.plane_dtype = plane_dtype;
    const int grid = (h->n_trees * 32 + SEARCH_BLOCK - 1) / SEARCH_BLOCK;
    if (h->select_mode == ZC_SELECT_PUCT) ZC_DISPATCH(h->game, k_select_puct<G><<<grid, SEARCH_BLOCK, 0, st>>>(p, h->sp_left));
    else ZC_DISPATCH(h->game, k_select<G><<<grid, SEARCH_BLOCK, 0, st>>>(p, h->sp_left));
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    h->sp_selected = h->sp_left < h->sp_batch ? h->sp_left : h->sp_batch;
    return ZC_OK;
}

extern "C" int zc_search_backprop(zc_search* h, const float* dev_values, void* stream) {
    if (int rc = check_handle(h)) return rc;
    if (!h->sp_selected) return fail(ZC_ESTATE, "backprop: nothing selected");
    if (!dev_values) return fail(ZC_EINVAL, "backprop: values is NULL");
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    SearchParams p = make_params(h, h->sp_left, h->sp_c, h->sp_batch, ZC_EVAL_EXTERNAL, h->sp_policy, h->sp_seed);
    p.values = dev_values;
    const int grid = (h->n_trees * 32 + SEARCH_BLOCK - 1) / SEARCH_BLOCK;
    if (h->select_mode == ZC_SELECT_PUCT) ZC_DISPATCH(h->game, k_backprop_puct<G><<<grid, SEARCH_BLOCK, 0, st>>>(p));
    else ZC_DISPATCH(h->game, k_backprop<G><<<grid, SEARCH_BLOCK, 0, st>>>(p));
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    h->sp_left -= h->sp_selected;
    h->sp_selected = 0;
    return ZC_OK;
}

static void launch_results_kernel(zc_search* h, zc_root_result* dst, int32_t* visits, double* wsum, zc_chess_move* moves, int stride,
                                  cudaStream_t st) {
    const int n = h->n_trees;
    if (h->game == ZC_GAME_C4)
        k_results_c4<<<(n + 127) / 128, 128, 0, st>>>(h->arena, h->arena_slots, h->ctl, n, dst, visits, wsum, moves, stride);
    else
        k_results_chess<<<(n + 127) / 128, 128, 0, st>>>(h->arena, h->arena_slots, h->ctl, n, dst, visits, wsum, moves, stride);
    h->launches++;
}

extern "C" int zc_search_results(zc_search* h, zc_root_result* results, int32_t* visits, double* value_sums,
                                 zc_chess_move* moves, int stride, void* stream) {
    if (int rc = check_handle(h)) return rc;
    if (!results) return fail(ZC_EINVAL, "results is NULL");
    if (h->n_trees < 1) return fail(ZC_ESTATE, "no roots set");
    if ((visits || value_sums || moves) && (stride < 1 || stride > ZC_MAX_MOVES)) return fail(ZC_EINVAL, "bad stride");
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int n = h->n_trees;
    if (stride > h->res_stride && (visits || value_sums || moves)) {
        cudaFree(h->visits_dev);
        cudaFree(h->wsum_dev);
        cudaFree(h->moves_dev);
        h->visits_dev = nullptr; h->wsum_dev = nullptr; h->moves_dev = nullptr;
        CUDA_TRY(cudaMalloc((void**)&h->visits_dev, sizeof(int32_t) * (size_t)h->max_trees * stride));
        CUDA_TRY(cudaMalloc((void**)&h->wsum_dev, sizeof(double) * (size_t)h->max_trees * stride));
        CUDA_TRY(cudaMalloc((void**)&h->moves_dev, sizeof(zc_chess_move) * (size_t)h->max_trees * stride));
        h->res_stride = stride;
    }
    if (visits || value_sums || moves) {
        CUDA_TRY(cudaMemsetAsync(h->visits_dev, 0, sizeof(int32_t) * (size_t)n * stride, st));
        CUDA_TRY(cudaMemsetAsync(h->wsum_dev, 0, sizeof(double) * (size_t)n * stride, st));
        CUDA_TRY(cudaMemsetAsync(h->moves_dev, 0, sizeof(zc_chess_move) * (size_t)n * stride, st));
    }
    launch_results_kernel(h, h->res_dev, visits ? h->visits_dev : nullptr, value_sums ? h->wsum_dev : nullptr,
                          moves ? h->moves_dev : nullptr, stride, st);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(h->res_host, h->res_dev, sizeof(zc_root_result) * n, cudaMemcpyDeviceToHost, st));
    if (visits) CUDA_TRY(cudaMemcpyAsync(visits, h->visits_dev, sizeof(int32_t) * (size_t)n * stride, cudaMemcpyDeviceToHost, st));
    if (value_sums) CUDA_TRY(cudaMemcpyAsync(value_sums, h->wsum_dev, sizeof(double) * (size_t)n * stride, cudaMemcpyDeviceToHost, st));
    if (moves) CUDA_TRY(cudaMemcpyAsync(moves, h->moves_dev, sizeof(zc_chess_move) * (size_t)n * stride, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    memcpy(results, h->res_host, sizeof(zc_root_result) * (size_t)n);
    for (int i = 0; i < n; ++i)
        if (results[i].status != 0) return fail(ZC_ECAPACITY, "tree " + std::to_string(i) + " outgrew its arena");
    return ZC_OK;
}

// Root readout without stalling the host: the per-tree results (mcts.cpp:150-159) are computed on `stream`, copied
// to pinned host memory on the handle's copy stream, and collected by zc_search_results_end.  The next
// set_roots / run may be enqueued at once; device results are double-buffered.
extern "C" int zc_search_results_begin(zc_search* h, void* stream) {
    if (int rc = check_handle(h)) return rc;
    if (h->n_trees < 1) return fail(ZC_ESTATE, "no roots set");
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int f = h->res_flip;
    CUDA_TRY(cudaStreamWaitEvent(st, h->ev_copy_done[f], 0));      // the copy that last read this buffer (two begins ago)
    launch_results_kernel(h, h->res_dev2[f], nullptr, nullptr, nullptr, 0, st);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(h->ev_res_ready, st));
    CUDA_TRY(cudaStreamWaitEvent(h->copy_stream, h->ev_res_ready, 0));
    CUDA_TRY(cudaMemcpyAsync(h->res_host, h->res_dev2[f], sizeof(zc_root_result) * (size_t)h->n_trees, cudaMemcpyDeviceToHost, h->copy_stream));
    CUDA_TRY(cudaEventRecord(h->ev_copy_done[f], h->copy_stream));
    h->res_pending = f;
    h->res_pending_n = h->n_trees;
    h->res_flip ^= 1;
    return ZC_OK;
}

extern "C" int zc_search_results_end(zc_search* h, zc_root_result* results) {
    if (int rc = check_handle(h)) return rc;
    if (!results) return fail(ZC_EINVAL, "results is NULL");
    if (h->res_pending < 0) return fail(ZC_ESTATE, "results_end without results_begin");
    CUDA_TRY(cudaSetDevice(h->device));
    CUDA_TRY(cudaEventSynchronize(h->ev_copy_done[h->res_pending]));
    const int n = h->res_pending_n;
    h->res_pending = -1;
    memcpy(results, h->res_host, sizeof(zc_root_result) * (size_t)n);
    for (int i = 0; i < n; ++i)
        if (results[i].status != 0) return fail(ZC_ECAPACITY, "tree " + std::to_string(i) + " outgrew its arena");
    return ZC_OK;
}

extern "C" int zc_search_tree_hash(zc_search* h, uint64_t* host_hashes, void* stream) {
    if (int rc = check_handle(h)) return rc;
    if (!host_hashes || h->n_trees < 1) return fail(ZC_EINVAL, "tree_hash: bad arguments");
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    ZC_DISPATCH(h->game, k_tree_hash<G><<<(h->n_trees + 63) / 64, 64, 0, st>>>(h->arena, h->arena_slots, h->n_trees, h->hash_dev));
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(host_hashes, h->hash_dev, sizeof(uint64_t) * h->n_trees, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    return ZC_OK;
}

extern "C" int zc_search_read_tree(zc_search* h, int tree, void* host_slots, int64_t max_slots, int64_t* used,
                                   int32_t* state_slots_out, void* stream) {
    if (int rc = check_handle(h)) return rc;
    if (tree < 0 || tree >= h->n_trees || !host_slots || !used || max_slots < 1) return fail(ZC_EINVAL, "read_tree: bad argument");
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    TreeCtl c;
    CUDA_TRY(cudaMemcpyAsync(&c, h->ctl + tree, sizeof c, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    *used = (int64_t)c.top;
    const int64_t n = *used < max_slots ? *used : max_slots;
    CUDA_TRY(cudaMemcpyAsync(host_slots, h->arena + (uint64_t)tree * h->arena_slots, sizeof(uint4) * (size_t)n,
                             cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (state_slots_out) *state_slots_out = h->game == ZC_GAME_C4 ? C4Game::SS : ChessGame::SS;
    return ZC_OK;
}

extern "C" int zc_search_get_counters(zc_search* h, zc_search_counters* out, void* stream) {
    if (int rc = check_handle(h)) return rc;
    if (!out) return fail(ZC_EINVAL, "out is NULL");
    memset(out, 0, sizeof *out);
    out->kernel_launches = h->launches;
    if (h->n_trees < 1) return ZC_OK;
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    std::vector<TreeCtl> c((size_t)h->n_trees);
    CUDA_TRY(cudaMemcpyAsync(c.data(), h->ctl, sizeof(TreeCtl) * c.size(), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    for (const TreeCtl& t : c) {
        out->simulations += t.sims_done;
        out->nodes += t.nodes;
        out->sum_leaf_depth += (int64_t)t.sum_leaf_depth;
        out->sum_path_children += (int64_t)t.sum_path_children;
        out->arena_slots_used += t.top;
    }
    return ZC_OK;
}

#include "rules_api.inl"
#include "tower_api.inl"
