// Game-rule entry points of the C-ABI (included by zc_api.cu, same translation unit so that the
// __constant__ move-order table is shared).
//   * single-state helpers: host execution of the host+device rule code (root bookkeeping);
//   * *_batch / perft: the same rule code as device kernels, for parity tests and bulk stepping.

// ------------------------------------------------------------------------------------------ C4 host
static c4::State c4_rel(const zc_c4_state* s) {
    c4::State r;
    r.cur = s->turn == 0 ? s->x : s->o;
    r.opp = s->turn == 0 ? s->o : s->x;
    return r;
}

extern "C" int zc_c4_init_state(zc_c4_state* out) {
    if (!out) return fail(ZC_EINVAL, "out is NULL");
    memset(out, 0, sizeof *out);
    return ZC_OK;
}
extern "C" int zc_c4_legal_moves(const zc_c4_state* s, int32_t* cols) {
    if (!s || !cols) return fail(ZC_EINVAL, "NULL argument");
    ensure_order();
    const c4::State r = c4_rel(s);
    const int mask = c4::legal_mask(r), n = zc_popc32((unsigned)mask);
    for (int i = 0; i < n; ++i) cols[i] = c4::move_col(mask, i);
    return n;
}
extern "C" int zc_c4_play_move(const zc_c4_state* s, int col, zc_c4_state* out) {
    if (!s || !out || col < 0 || col > 6) return fail(ZC_EINVAL, "bad argument");
    const c4::State n = c4::play(c4_rel(s), col);   // n.cur = old opponent, n.opp = mover incl. the new disc
    zc_c4_state o;
    memset(&o, 0, sizeof o);
    o.turn = 1 - s->turn;
    o.x = s->turn == 0 ? n.opp : n.cur;
    o.o = s->turn == 0 ? n.cur : n.opp;
    *out = o;
    return ZC_OK;
}
extern "C" int zc_c4_check_win(const zc_c4_state* s) {
    if (!s) return fail(ZC_EINVAL, "NULL argument");
    return c4::check_win(c4_rel(s)) ? 1 : 0;
}
extern "C" int zc_c4_check_draw(const zc_c4_state* s) {
    if (!s) return fail(ZC_EINVAL, "NULL argument");
    return c4::check_draw(c4_rel(s)) ? 1 : 0;
}
extern "C" int zc_c4_to_tensor(const zc_c4_state* s, float* out) {
    if (!s || !out) return fail(ZC_EINVAL, "NULL argument");
    const c4::State r = c4_rel(s);
    for (int row = 0; row < 6; ++row)
        for (int c = 0; c < 7; ++c) {
            const int b = c * 7 + (5 - row);
            out[row * 7 + c] = (float)((r.cur >> b) & 1ull);
            out[42 + row * 7 + c] = (float)((r.opp >> b) & 1ull);
        }
    return ZC_OK;
}

// ------------------------------------------------------------------------------------------ chess host
__host__ __device__ static inline chess::Board board_of(const zc_chess_state& s) {
    chess::Board b = {0, 0, 0, 0};
    for (int i = 0; i < 64; ++i) chess::put_piece(b, i, chess::code_of_char(s.board[i]));
    return b;
}
__host__ __device__ static inline uint32_t misc_of(const zc_chess_state& s) {
    return (s.turn ? chess::MISC_TURN : 0u) | (s.w_ck ? chess::MISC_WCK : 0u) | (s.w_cq ? chess::MISC_WCQ : 0u) |
           (s.b_ck ? chess::MISC_BCK : 0u) | (s.b_cq ? chess::MISC_BCQ : 0u);
}

extern "C" int zc_chess_init_state(zc_chess_state* out) {
    if (!out) return fail(ZC_EINVAL, "out is NULL");
    memset(out, 0, sizeof *out);
    memcpy(out->board, "rnbqkbnrpppppppp                                PPPPPPPPRNBQKBNR", 64);
    out->w_ck = out->w_cq = out->b_ck = out->b_cq = 1;
    return ZC_OK;
}

extern "C" int zc_chess_from_fen(const char* fen, zc_chess_state* out) {
    if (!fen || !out) return fail(ZC_EINVAL, "NULL argument");
    memset(out, 0, sizeof *out);
    memset(out->board, ' ', 64);
    // six whitespace-separated fields; en passant and the full-move number are ignored, a
    // missing half-move clock reads as 0 (chess_backend.cpp:525-556)
    std::string f[6];
    int nf = 0;
    for (const char* p = fen; *p && nf < 6;) {
        while (*p == ' ' || *p == '\t' || *p == '\n') ++p;
        if (!*p) break;
        const char* q = p;
        while (*q && *q != ' ' && *q != '\t' && *q != '\n') ++q;
        f[nf++] = std::string(p, q);
        p = q;
    }
    int idx = 0;
    for (char ch : f[0]) {
        if (ch == '/') continue;
        if (ch >= '0' && ch <= '9') {
            idx += ch - '0';
        } else if (idx < 64) {
            out->board[idx++] = (uint8_t)ch;
        }
        if (idx > 64) return fail(ZC_EINVAL, "FEN: more than 64 squares");
    }
    out->turn = f[1] == "w" ? 0 : 1;
    out->w_ck = f[2].find('K') != std::string::npos;
    out->w_cq = f[2].find('Q') != std::string::npos;
    out->b_ck = f[2].find('k') != std::string::npos;
    out->b_cq = f[2].find('q') != std::string::npos;
    int hm = 0;
    if (nf >= 5) {
        char* end = nullptr;
        const long v = strtol(f[4].c_str(), &end, 10);
        if (end != f[4].c_str()) hm = (int)v;
    }
    out->fifty_move_rule_counter = (uint8_t)hm;
    return ZC_OK;
}

extern "C" int zc_chess_legal_moves(const zc_chess_state* s, zc_chess_move* out) {
    if (!s || !out) return fail(ZC_EINVAL, "NULL argument");
    const chess::Board b = board_of(*s);
    uint16_t mv[chess::MAX_PSEUDO];
    const int n = chess::generate(b, s->turn ? 1 : 0, mv);
    for (int i = 0; i < n; ++i) out[i] = decode_move(b, mv[i]);
    return n;
}

extern "C" int zc_chess_play_move(const zc_chess_state* s, const zc_chess_move* m, zc_chess_state* out) {
    if (!s || !m || !out) return fail(ZC_EINVAL, "NULL argument");
    if (m->fr > 7 || m->fc > 7 || m->tr > 7 || m->tc > 7) return fail(ZC_EINVAL, "move off the board");
    const chess::Board b = board_of(*s);
    const int from = m->fr * 8 + m->fc, to = m->tr * 8 + m->tc;
    uint32_t nm;
    const chess::Board nb = chess::play(b, misc_of(*s), from, to, nm);
    zc_chess_state o;
    memset(&o, 0, sizeof o);
    for (int i = 0; i < 64; ++i) o.board[i] = chess::char_of_code(chess::piece_at(nb, i));
    o.turn = (uint8_t)(nm & chess::MISC_TURN);
    o.w_ck = (nm & chess::MISC_WCK) != 0;
    o.w_cq = (nm & chess::MISC_WCQ) != 0;
    o.b_ck = (nm & chess::MISC_BCK) != 0;
    o.b_cq = (nm & chess::MISC_BCQ) != 0;
    o.fifty_move_rule_counter = chess::resets_fifty(b, from, to) ? 0 : (uint8_t)(s->fifty_move_rule_counter + 1);
    *out = o;
    return ZC_OK;
}

extern "C" int zc_chess_check_win(const zc_chess_state* s) {
    if (!s) return fail(ZC_EINVAL, "NULL argument");
    const chess::Board b = board_of(*s);
    const int turn = s->turn ? 1 : 0;
    uint16_t mv[chess::MAX_PSEUDO];
    return chess::generate(b, turn, mv) == 0 && chess::in_check(b, turn) ? 1 : 0;
}

// chess_backend.cpp:148-180: some prefix of the (most-recent-first) list is >= 3 repeats of a period >= 2
static bool repeated_prefix(const zc_chess_move* L, int n, int min_len, int min_rep) {
    if (n < min_len * min_rep) return false;
    auto eq = [&](int a, int b) {
        return L[a].fr == L[b].fr && L[a].fc == L[b].fc && L[a].tr == L[b].tr && L[a].tc == L[b].tc && L[a].value == L[b].value;
    };
    std::vector<int> pi((size_t)n, 0);
    for (int i = 1, j = 0; i < n; ++i) {
        while (j > 0 && !eq(i, j)) j = pi[(size_t)j - 1];
        if (eq(i, j)) ++j;
        pi[(size_t)i] = j;
    }
    for (int i = 0; i < n; ++i) {
        const int len = i + 1, period = len - pi[(size_t)i];
        if (period >= min_len && len % period == 0 && len / period >= min_rep) return true;
    }
    return false;
}

extern "C" int zc_chess_check_draw(const zc_chess_state* s, const zc_chess_move* hist_white, int n_white,
                                   const zc_chess_move* hist_black, int n_black) {
    if (!s || n_white < 0 || n_black < 0 || (n_white && !hist_white) || (n_black && !hist_black))
        return fail(ZC_EINVAL, "bad argument");
    const chess::Board b = board_of(*s);
    const int turn = s->turn ? 1 : 0;
    uint16_t mv[chess::MAX_PSEUDO];
    if (chess::generate(b, turn, mv) == 0 && !chess::in_check(b, turn)) return 1;   // stalemate branch (:419-427)
    if (s->fifty_move_rule_counter >= 50) return 1;                                       // plies, :429
    if (repeated_prefix(hist_white, n_white, 2, 3) && repeated_prefix(hist_black, n_black, 2, 3)) return 1;
    return 0;
}

extern "C" int zc_chess_to_tensor(const zc_chess_state* s, float* out) {
    if (!s || !out) return fail(ZC_EINVAL, "NULL argument");
    const chess::Board b = board_of(*s);
    const uint32_t misc = misc_of(*s);
    for (int pl = 0; pl < 17; ++pl) {
        const uint64_t bb = ChessGame::plane_bits(b, misc, pl);
        for (int i = 0; i < 64; ++i) out[pl * 64 + i] = (float)((bb >> i) & 1ull);
    }
    return ZC_OK;
}

// ------------------------------------------------------------------------------------------ device batches
struct PerftState {
    chess::Board b;
    uint32_t misc;
    uint32_t pad;
};

__global__ void k_chess_legal_batch(const zc_chess_state* __restrict__ states, int n, zc_chess_move* __restrict__ moves,
                                    int32_t* __restrict__ counts, int32_t* __restrict__ flags, uint16_t* __restrict__ scratch) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const zc_chess_state s = states[t];
    const chess::Board b = board_of(s);
    const int turn = s.turn ? 1 : 0;
    uint16_t* mv = scratch + (size_t)t * ChessGame::MOVE_SCRATCH;
    const int k = chess::generate(b, turn, mv);
    for (int i = 0; i < k; ++i) moves[(size_t)t * ZC_MAX_MOVES + i] = decode_move(b, mv[i]);
    counts[t] = k;
    const bool chk = chess::in_check(b, turn);
    flags[t] = (k == 0 && chk ? 1 : 0) | (k == 0 && !chk ? 2 : 0) | (chk ? 4 : 0);
}

// the warp-cooperative generator (one position per warp) behind the same outputs
__global__ void k_chess_legal_batch_warp(const zc_chess_state* __restrict__ states, int n, zc_chess_move* __restrict__ moves,
                                         int32_t* __restrict__ counts, int32_t* __restrict__ flags, uint16_t* __restrict__ scratch) {
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (t >= n) return;
    const zc_chess_state s = states[t];
    const chess::Board b = board_of(s);
    const int turn = s.turn ? 1 : 0;
    uint16_t* mv = scratch + (size_t)t * ChessGame::MOVE_SCRATCH;
    const int k = chess::generate_warp(b, turn, mv, 1, lane);
    for (int i = lane; i < k; i += 32) moves[(size_t)t * ZC_MAX_MOVES + i] = decode_move(b, mv[i]);
    if (lane == 0) {
        counts[t] = k;
        const bool chk = chess::in_check(b, turn);
        flags[t] = (k == 0 && chk ? 1 : 0) | (k == 0 && !chk ? 2 : 0) | (chk ? 4 : 0);
    }
}

__global__ void k_perft_count(const PerftState* __restrict__ frontier, unsigned long long n,
                              unsigned long long* __restrict__ total) {
    const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long mine = 0;
    uint16_t mv[chess::MAX_PSEUDO];
    if (t < n) mine = (unsigned long long)chess::generate(frontier[t].b, (int)(frontier[t].misc & 1u), mv);
    for (int d = 16; d >= 1; d >>= 1) mine += __shfl_xor_sync(0xFFFFFFFFu, mine, d);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(total, mine);
}

__global__ void k_perft_expand(const PerftState* __restrict__ frontier, unsigned long long n, PerftState* __restrict__ next,
                               unsigned long long* __restrict__ cursor, uint16_t* __restrict__ scratch) {
    const unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const PerftState s = frontier[t];
    uint16_t* mv = scratch + t * ChessGame::MOVE_SCRATCH;
    const int k = chess::generate(s.b, (int)(s.misc & 1u), mv);
    if (k == 0) return;
    const unsigned long long base = atomicAdd(cursor, (unsigned long long)k);
    for (int i = 0; i < k; ++i) {
        PerftState c;
        c.b = chess::play(s.b, s.misc, chess::move_from(mv[i]), chess::move_to(mv[i]), c.misc);
        c.pad = 0;
        next[base + i] = c;
    }
}

static int use_device(int device) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(ZC_ENODEVICE, "no CUDA device: libzc_b200 has no CPU path");
    }
    if (device < 0 || device >= ndev) return fail(ZC_EINVAL, "device out of range");
    CUDA_TRY(cudaSetDevice(device));
    return ZC_OK;
}

struct DevBuf {   // frees on scope exit
    void* p = nullptr;
    ~DevBuf() { cudaFree(p); }
};

static int chess_legal_moves_batch(int device, const zc_chess_state* states, int n, zc_chess_move* moves, int32_t* counts,
                                   int32_t* flags, bool warp_cooperative);
extern "C" int zc_chess_legal_moves_batch(int device, const zc_chess_state* states, int n, zc_chess_move* moves,
                                          int32_t* counts, int32_t* flags) {
    return chess_legal_moves_batch(device, states, n, moves, counts, flags, false);
}
extern "C" int zc_chess_legal_moves_batch_warp(int device, const zc_chess_state* states, int n, zc_chess_move* moves,
                                               int32_t* counts, int32_t* flags) {
    return chess_legal_moves_batch(device, states, n, moves, counts, flags, true);
}
static int chess_legal_moves_batch(int device, const zc_chess_state* states, int n, zc_chess_move* moves, int32_t* counts,
                                   int32_t* flags, bool warp_cooperative) {
    if (!states || !moves || !counts || !flags || n < 1) return fail(ZC_EINVAL, "bad argument");
    if (int rc = use_device(device)) return rc;
    DevBuf ds, dm, dc, df, dscr;
    CUDA_TRY(cudaMalloc(&ds.p, sizeof(zc_chess_state) * (size_t)n));
    CUDA_TRY(cudaMalloc(&dm.p, sizeof(zc_chess_move) * (size_t)n * ZC_MAX_MOVES));
    CUDA_TRY(cudaMalloc(&dc.p, sizeof(int32_t) * (size_t)n));
    CUDA_TRY(cudaMalloc(&df.p, sizeof(int32_t) * (size_t)n));
    CUDA_TRY(cudaMalloc(&dscr.p, sizeof(uint16_t) * (size_t)n * ChessGame::MOVE_SCRATCH));
    CUDA_TRY(cudaMemcpy(ds.p, states, sizeof(zc_chess_state) * (size_t)n, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemset(dm.p, 0, sizeof(zc_chess_move) * (size_t)n * ZC_MAX_MOVES));
    if (warp_cooperative)
        k_chess_legal_batch_warp<<<(n + 3) / 4, 128>>>((const zc_chess_state*)ds.p, n, (zc_chess_move*)dm.p, (int32_t*)dc.p,
                                                       (int32_t*)df.p, (uint16_t*)dscr.p);
    else
        k_chess_legal_batch<<<(n + 63) / 64, 64>>>((const zc_chess_state*)ds.p, n, (zc_chess_move*)dm.p, (int32_t*)dc.p,
                                                   (int32_t*)df.p, (uint16_t*)dscr.p);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpy(moves, dm.p, sizeof(zc_chess_move) * (size_t)n * ZC_MAX_MOVES, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(counts, dc.p, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(flags, df.p, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost));
    return ZC_OK;
}

extern "C" int zc_chess_perft(int device, const zc_chess_state* root, int depth, uint64_t* out) {
    if (!root || !out || depth < 1 || depth > 12) return fail(ZC_EINVAL, "bad argument");
    if (int rc = use_device(device)) return rc;
    PerftState r0;
    r0.b = board_of(*root);
    r0.misc = misc_of(*root);
    r0.pad = 0;
    DevBuf cur, counter;
    unsigned long long n = 1;
    CUDA_TRY(cudaMalloc(&cur.p, sizeof(PerftState)));
    CUDA_TRY(cudaMemcpy(cur.p, &r0, sizeof r0, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc(&counter.p, sizeof(unsigned long long)));
    for (int d = 1; d <= depth; ++d) {
        // how many positions does the next ply hold?
        unsigned long long total = 0;
        CUDA_TRY(cudaMemset(counter.p, 0, sizeof(unsigned long long)));
        if (n) {
            k_perft_count<<<(unsigned)((n + 127) / 128), 128>>>((const PerftState*)cur.p, n, (unsigned long long*)counter.p);
            CUDA_TRY(cudaGetLastError());
        }
        CUDA_TRY(cudaMemcpy(&total, counter.p, sizeof total, cudaMemcpyDeviceToHost));
        if (d == depth) {
            *out = total;
            return ZC_OK;
        }
        if (total > (1ull << 28)) return fail(ZC_ECAPACITY, "perft frontier too large for one device buffer");
        DevBuf next, scr;
        CUDA_TRY(cudaMalloc(&next.p, sizeof(PerftState) * (size_t)(total ? total : 1)));
        CUDA_TRY(cudaMalloc(&scr.p, sizeof(uint16_t) * (size_t)(n ? n : 1) * ChessGame::MOVE_SCRATCH));
        CUDA_TRY(cudaMemset(counter.p, 0, sizeof(unsigned long long)));
        if (n) {
            k_perft_expand<<<(unsigned)((n + 127) / 128), 128>>>((const PerftState*)cur.p, n, (PerftState*)next.p,
                                                                 (unsigned long long*)counter.p, (uint16_t*)scr.p);
            CUDA_TRY(cudaGetLastError());
        }
        CUDA_TRY(cudaDeviceSynchronize());
        std::swap(cur.p, next.p);
        n = total;
    }
    return ZC_OK;
}

__global__ void k_c4_rules_batch(const zc_c4_state* __restrict__ states, int n, uint8_t* __restrict__ cols,
                                 int32_t* __restrict__ flags) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const zc_c4_state s = states[t];
    c4::State r;
    r.cur = s.turn == 0 ? s.x : s.o;
    r.opp = s.turn == 0 ? s.o : s.x;
    const int mask = c4::legal_mask(r), k = zc_popc32((unsigned)mask);
    for (int i = 0; i < 8; ++i) cols[(size_t)t * 8 + i] = i < k ? (uint8_t)c4::move_col(mask, i) : 255;
    flags[t] = (c4::check_win(r) ? 1 : 0) | (c4::check_draw(r) ? 2 : 0);
}

extern "C" int zc_c4_rules_batch(int device, const zc_c4_state* states, int n, uint8_t* cols, int32_t* flags) {
    if (!states || !cols || !flags || n < 1) return fail(ZC_EINVAL, "bad argument");
    if (int rc = use_device(device)) return rc;
    if (int rc = upload_order()) return rc;
    DevBuf ds, dc, df;
    CUDA_TRY(cudaMalloc(&ds.p, sizeof(zc_c4_state) * (size_t)n));
    CUDA_TRY(cudaMalloc(&dc.p, (size_t)n * 8));
    CUDA_TRY(cudaMalloc(&df.p, sizeof(int32_t) * (size_t)n));
    CUDA_TRY(cudaMemcpy(ds.p, states, sizeof(zc_c4_state) * (size_t)n, cudaMemcpyHostToDevice));
    k_c4_rules_batch<<<(n + 127) / 128, 128>>>((const zc_c4_state*)ds.p, n, (uint8_t*)dc.p, (int32_t*)df.p);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpy(cols, dc.p, (size_t)n * 8, cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaMemcpy(flags, df.p, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost));
    return ZC_OK;
}

// ------------------------------------------------------------------------------------------ self-play step
__device__ static int root_best_edge(const uint4* arena, int k, int ss) {
    int best = -1, best_n = -1;
    for (int i = 0; i < k; ++i) {
        const uint4 e = arena[1 + ss + i];
        if (e.w && (int)e.z > best_n) { best_n = (int)e.z; best = i; }
    }
    return best;
}

// what can go wrong for a tree in the self-play step; reported through err[0] (flags) and err[1] (lowest tree index)
constexpr int ADV_ERR_OVERFLOW = 1;   // the tree outgrew its arena during the search: its move must not be played
constexpr int ADV_ERR_NO_MOVE = 2;    // an active tree has no visited root child (no simulation ran, or a terminal root)
constexpr int ADV_ERR_HISTORY = 4;    // chess: the side's move history is full (repetition detection would go stale)
__device__ static void adv_report(int32_t* err, int flag, int tree) {
    atomicOr(err, flag);
    atomicMin(err + 1, tree);
}

__global__ void k_advance_c4(const uint4* __restrict__ arena_all, uint64_t arena_slots, const TreeCtl* __restrict__ ctl, int n,
                             zc_c4_state* __restrict__ states, const uint8_t* __restrict__ active, int32_t* __restrict__ results,
                             zc_chess_move* __restrict__ moves, int32_t* __restrict__ err) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    zc_chess_move mv = {0, 0, 0, 0, 0.f};
    int result = ZC_RESULT_ONGOING;
    if (active[t]) {
        const uint4* arena = arena_all + (uint64_t)t * arena_slots;
        const int k = (int)hdr_k(arena[0]);
        const int best = root_best_edge(arena, k, C4Game::SS);
        if (ctl[t].status != 0) adv_report(err, ADV_ERR_OVERFLOW, t);
        else if (best < 0) adv_report(err, ADV_ERR_NO_MOVE, t);
        else {
            zc_c4_state s = states[t];
            c4::State r;
            r.cur = s.turn == 0 ? s.x : s.o;
            r.opp = s.turn == 0 ? s.o : s.x;
            const int col = c4::move_col(c4::legal_mask(r), best);
            const c4::State nx = c4::play(r, col);
            zc_c4_state o;
            o.turn = 1 - s.turn;
            o.x = s.turn == 0 ? nx.opp : nx.cur;
            o.o = s.turn == 0 ? nx.cur : nx.opp;
            o.reserved = 0;
            states[t] = o;
            mv.fr = (uint8_t)col;
            if (c4::check_win(nx)) result = o.turn * 2 - 1;          // engine.py:149-150
            else if (c4::check_draw(nx)) result = 0;
        }
    }
    results[t] = result;
    moves[t] = mv;
}

// chess_backend.cpp:148-180 on a history kept in playing order: L[i] = hist[len-1-i] (most recent first)
__device__ static bool dev_repeated_prefix(const zc_chess_move* hist, int n, int* pi) {
    if (n < 6) return false;
    auto eq = [&](int a, int b) {
        const zc_chess_move x = hist[n - 1 - a], y = hist[n - 1 - b];
        return x.fr == y.fr && x.fc == y.fc && x.tr == y.tr && x.tc == y.tc && x.value == y.value;
    };
    pi[0] = 0;
    for (int i = 1, j = 0; i < n; ++i) {
        while (j > 0 && !eq(i, j)) j = pi[j - 1];
        if (eq(i, j)) ++j;
        pi[i] = j;
    }
    for (int i = 0; i < n; ++i) {
        const int len = i + 1, period = len - pi[i];
        if (period >= 2 && len % period == 0 && len / period >= 3) return true;
    }
    return false;
}

__global__ void k_advance_chess(const uint4* __restrict__ arena_all, uint64_t arena_slots, const TreeCtl* __restrict__ ctl, int n,
                                zc_chess_state* __restrict__ states, const uint8_t* __restrict__ active,
                                zc_chess_move* __restrict__ hist, int32_t* __restrict__ hist_len, int hist_cap,
                                int* __restrict__ kmp_scratch, uint16_t* __restrict__ move_scratch,
                                int32_t* __restrict__ results, zc_chess_move* __restrict__ moves, int32_t* __restrict__ err) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    zc_chess_move mv = {0, 0, 0, 0, 0.f};
    int result = ZC_RESULT_ONGOING;
    if (active[t]) {
        const uint4* arena = arena_all + (uint64_t)t * arena_slots;
        const int k = (int)hdr_k(arena[0]);
        const int best = root_best_edge(arena, k, ChessGame::SS);
        const int side = states[t].turn ? 1 : 0;
        if (ctl[t].status != 0) adv_report(err, ADV_ERR_OVERFLOW, t);
        else if (best < 0) adv_report(err, ADV_ERR_NO_MOVE, t);
        else if (hist_len[t * 2 + side] >= hist_cap) adv_report(err, ADV_ERR_HISTORY, t);
        else {
            const zc_chess_state s = states[t];
            const chess::Board b = ChessGame::load_state(arena + 1);
            const uint16_t m = ChessGame::move_at(arena, k, best);
            mv = decode_move(b, m);
            const int from = chess::move_from(m), to = chess::move_to(m), turn = s.turn ? 1 : 0;
            uint32_t nm;
            const chess::Board nb = chess::play(b, misc_of(s), from, to, nm);
            zc_chess_state o = s;
            for (int i = 0; i < 64; ++i) o.board[i] = chess::char_of_code(chess::piece_at(nb, i));
            o.turn = (uint8_t)(nm & chess::MISC_TURN);
            o.w_ck = (nm & chess::MISC_WCK) != 0;
            o.w_cq = (nm & chess::MISC_WCQ) != 0;
            o.b_ck = (nm & chess::MISC_BCK) != 0;
            o.b_cq = (nm & chess::MISC_BCQ) != 0;
            o.fifty_move_rule_counter = chess::resets_fifty(b, from, to) ? 0 : (uint8_t)(s.fifty_move_rule_counter + 1);
            states[t] = o;
            // the mover's history grows by this move (hist_white.push_front / hist_black.push_front, :374)
            zc_chess_move* hw = hist + ((size_t)t * 2 + 0) * hist_cap;
            zc_chess_move* hb = hist + ((size_t)t * 2 + 1) * hist_cap;
            int nw = hist_len[t * 2 + 0], nbk = hist_len[t * 2 + 1];
            if (turn == 0) { hw[nw++] = mv; hist_len[t * 2 + 0] = nw; }      // room was checked above
            else { hb[nbk++] = mv; hist_len[t * 2 + 1] = nbk; }
            // _evaluate(new state): check_win, then check_draw (:404-441)
            const int nturn = o.turn ? 1 : 0;
            const int nmoves = chess::generate(nb, nturn, move_scratch + (size_t)t * chess::MAX_PSEUDO);
            const bool chk = chess::in_check(nb, nturn);
            if (nmoves == 0 && chk) result = nturn * 2 - 1;
            else if (nmoves == 0) result = 0;                                        // stalemate / insufficient material
            else if (o.fifty_move_rule_counter >= 50) result = 0;                    // plies, :429
            else {
                int* pi = kmp_scratch + (size_t)t * hist_cap;
                if (dev_repeated_prefix(hw, nw, pi) && dev_repeated_prefix(hb, nbk, pi)) result = 0;
            }
        }
    }
    results[t] = result;
    moves[t] = mv;
}

extern "C" int zc_search_advance(zc_search* h, void* dev_states, const uint8_t* dev_active, zc_chess_move* dev_hist,
                                 int32_t* dev_hist_len, int hist_cap, int32_t* host_results, zc_chess_move* host_moves,
                                 void* stream) {
    if (int rc = check_handle(h)) return rc;
    if (!dev_states || !dev_active || !host_results || !host_moves) return fail(ZC_EINVAL, "advance: NULL argument");
    if (h->n_trees < 1) return fail(ZC_ESTATE, "no roots set");
    CUDA_TRY(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int n = h->n_trees;
    if (!h->adv_res_dev) {
        CUDA_TRY(cudaMalloc((void**)&h->adv_res_dev, sizeof(int32_t) * (size_t)h->max_trees));
        CUDA_TRY(cudaMalloc((void**)&h->adv_mv_dev, sizeof(zc_chess_move) * (size_t)h->max_trees));
        CUDA_TRY(cudaMalloc((void**)&h->adv_err_dev, sizeof(int32_t) * 2));
    }
    const int32_t err0[2] = {0, 0x7FFFFFFF};
    CUDA_TRY(cudaMemcpyAsync(h->adv_err_dev, err0, sizeof err0, cudaMemcpyHostToDevice, st));
    if (h->game == ZC_GAME_C4) {
        k_advance_c4<<<(n + 127) / 128, 128, 0, st>>>(h->arena, h->arena_slots, h->ctl, n, (zc_c4_state*)dev_states, dev_active,
                                                    h->adv_res_dev, h->adv_mv_dev, h->adv_err_dev);
    } else {
        if (!dev_hist || !dev_hist_len || hist_cap < 8) return fail(ZC_EINVAL, "advance: chess needs history buffers");
        if (h->adv_kmp_cap < hist_cap) {
            cudaFree(h->adv_kmp_dev);
            h->adv_kmp_dev = nullptr;
            CUDA_TRY(cudaMalloc((void**)&h->adv_kmp_dev, sizeof(int) * (size_t)h->max_trees * hist_cap));
            h->adv_kmp_cap = hist_cap;
        }
        k_advance_chess<<<(n + 63) / 64, 64, 0, st>>>(h->arena, h->arena_slots, h->ctl, n, (zc_chess_state*)dev_states, dev_active,
                                                    dev_hist, dev_hist_len, hist_cap, h->adv_kmp_dev, h->scratch,
                                                    h->adv_res_dev, h->adv_mv_dev, h->adv_err_dev);
    }
    h->launches++;
    CUDA_TRY(cudaGetLastError());
    int32_t err[2] = {0, 0};
    CUDA_TRY(cudaMemcpyAsync(host_results, h->adv_res_dev, sizeof(int32_t) * (size_t)n, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(host_moves, h->adv_mv_dev, sizeof(zc_chess_move) * (size_t)n, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaMemcpyAsync(err, h->adv_err_dev, sizeof err, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    // never silent: a tree whose search was cut short, found no move, or whose history is full stays as it was
    // (state unchanged, result ONGOING) and the call fails
    if (err[0] & ADV_ERR_OVERFLOW) return fail(ZC_ECAPACITY, "advance: tree " + std::to_string(err[1]) + " outgrew its arena; its move was not played");
    if (err[0] & ADV_ERR_HISTORY) return fail(ZC_ECAPACITY, "advance: move history of tree " + std::to_string(err[1]) + " is full (hist_cap); grow the history buffers");
    if (err[0] & ADV_ERR_NO_MOVE) return fail(ZC_ESTATE, "advance: active tree " + std::to_string(err[1]) + " has no visited root move (no simulation ran or the root is terminal)");
    return ZC_OK;
}

extern "C" int zc_states_to_tensor(int game, const void* states, int n, float* out) {
    if (!states || !out || n < 0) return fail(ZC_EINVAL, "bad argument");
    if (game == ZC_GAME_C4) {
        const zc_c4_state* s = (const zc_c4_state*)states;
        for (int i = 0; i < n; ++i) zc_c4_to_tensor(s + i, out + (size_t)i * 84);
    } else if (game == ZC_GAME_CHESS) {
        const zc_chess_state* s = (const zc_chess_state*)states;
        for (int i = 0; i < n; ++i) zc_chess_to_tensor(s + i, out + (size_t)i * 17 * 64);
    } else {
        return fail(ZC_EINVAL, "unknown game");
    }
    return ZC_OK;
}
