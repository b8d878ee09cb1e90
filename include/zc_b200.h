/*
 * zc_b200.h -- C-ABI of libzc_b200.so, the B200-native batched MCTS engine.
 *
 * This is the drop-in boundary for ONE hot path of rishabhgoel0213/ZeroClone:
 *     Engine.play_mcts / play_mcts_parallel   (engine/engine.py:119-138)
 *       -> engine.mcts.get_move               (engine/mcts/src/mcts.cpp:102-160,
 *                                              bound at engine/mcts/src/bindings_mcts.cpp:9-11)
 *       -> backend.get_legal_moves/play_move  (engine/games/chess/src/bindings_chess.cpp:44-57,
 *                                              engine/games/connect4/c4_backend.py:11-61)
 *       -> Value.batch                        (engine/value_functions.py:16-32)
 * In the reference these are pybind11 modules called once per tree and once per simulation;
 * here one handle owns thousands of trees in HBM and one call advances all of them.
 * INTEGRATION.md shows the ctypes binding a reference maintainer would add.
 *
 * Conventions
 *   - every function returns 0 on success or a negative ZC_E* code; zc_last_error() gives text;
 *     nothing throws across the ABI; no torch types anywhere;
 *   - "host" pointers are ordinary (preferably pinned) memory, "dev" pointers are device memory
 *     on the handle's device (e.g. torch.Tensor.data_ptr());
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream); all device
 *     work of a call is enqueued on it, host-buffer variants synchronise that stream before
 *     returning;
 *   - no CPU fallback exists: without a CUDA device every entry point that computes returns
 *     ZC_ENODEVICE.
 */
#ifndef ZC_B200_H
#define ZC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZC_ABI_VERSION 2

/* error codes */
#define ZC_OK 0
#define ZC_EINVAL (-1)    /* bad argument                                        */
#define ZC_ENODEVICE (-2) /* no usable CUDA device                               */
#define ZC_ECUDA (-3)     /* a CUDA call failed, see zc_last_error()             */
#define ZC_ECAPACITY (-4) /* a tree outgrew its node arena (raise arena_slots)   */
#define ZC_ESTATE (-5)    /* call sequence violated (e.g. backprop before select) */

/* games: engine/games/<game>/ in the reference */
#define ZC_GAME_C4 0
#define ZC_GAME_CHESS 1

/* leaf evaluators (engine/value_functions.py) */
#define ZC_EVAL_C4_TERMINAL 0   /* -1 if check_win(state) else 0 : terminal branch of random_rollout, :41-43   */
#define ZC_EVAL_C4_POSITIONAL 1 /* -1 on a win, else sum_discs w[col]*(+1 side to move / -1 opponent) / 64       */
#define ZC_EVAL_CHESS_CRUDE 2   /* crude_chess_score, :49-55                                                   */
#define ZC_EVAL_EXTERNAL 3      /* values supplied by the caller per batch (the neural evaluator, :61-129)     */
#define ZC_EVAL_C4_ROLLOUT 4    /* random_rollout, :35-45, device RNG (throughput only, not bit-reproducible)   */

/* expansion policies (engine/policy_functions.py; mcts.cpp:65-78 takes any callable) */
#define ZC_POLICY_FIRST 0  /* policy(moves) = moves[0]  : deterministic, used for parity                        */
#define ZC_POLICY_LAST 1   /* policy(moves) = moves[-1] : deterministic, used for parity                        */
#define ZC_POLICY_RANDOM 2 /* Policy.random, :10-12 : uniform among untried moves, device RNG                   */
#define ZC_POLICY_IMMEDIATE_VALUE 3 /* Policy.immediate_value, :14-17 : uniform among the untried moves whose capture value
                                     * move[1] is >= (best untried value - policy_freedom), device RNG.  Connect Four
                                     * moves all carry value 0 (c4_backend.py:50), so there it equals ZC_POLICY_RANDOM. */

/* ---- packed root states ---------------------------------------------------------------- */

/* Connect Four (c4_backend.py:4 State(board, turn)).  Cell (row r, column c), row 0 = top,
 * is bit  c*7 + (5-r)  ; bit c*7+6 is always 0.  x = discs of 'X' (turn 0), o = discs of 'O'. */
typedef struct zc_c4_state {
    uint64_t x;
    uint64_t o;
    int32_t turn;
    int32_t reserved;
} zc_c4_state;

/* Chess (include/state.h:9-23 without the two history deques, which only check_draw reads).
 * board[r*8+c], row 0 = rank 8, ASCII piece letters, ' ' or 0 = empty. */
typedef struct zc_chess_state {
    uint8_t board[64];
    uint8_t turn; /* 0 = white */
    uint8_t fifty_move_rule_counter;
    uint8_t w_ck, w_cq, b_ck, b_cq;
    uint8_t reserved[2];
} zc_chess_state;

/* a chess move as the reference prints it: ((fr,fc,tr,tc), value) -- chess_backend.cpp:52-64 */
typedef struct zc_chess_move {
    uint8_t fr, fc, tr, tc;
    float value; /* captured piece value 0/1/3/5/9 */
} zc_chess_move;

#define ZC_MAX_MOVES 256 /* >= 218, the chess maximum; C4 uses 7 */

/* per-tree result of a search: what get_move returns (mcts.cpp:150-159) plus the statistics the
 * reference computes but does not expose */
typedef struct zc_root_result {
    int32_t n_moves;            /* legal moves at the root, in backend order                */
    int32_t best;               /* index of the chosen move, -1 if no child was visited     */
    int32_t root_visits;        /* root N                                                   */
    int32_t status;             /* 0, or ZC_ECAPACITY for this tree                         */
    uint8_t best_move[4];       /* chess fr,fc,tr,tc ; C4 col,0,0,0                         */
    float best_move_value;      /* chess capture value of the chosen move                   */
    int32_t nodes;              /* nodes in the tree                                        */
    int64_t sum_leaf_depth;     /* sum over simulations of the evaluated leaf's depth       */
    int32_t max_leaf_depth;
    int32_t reevaluated_leaves; /* simulations that re-evaluated a move-less node           */
} zc_root_result;

typedef struct zc_search zc_search; /* opaque: node arenas + control blocks of up to max_trees trees */

/* ---- library ----------------------------------------------------------------------------- */
int zc_abi_version(void);
const char *zc_last_error(void);
int zc_device_count(void);

/* CPython set-iteration order of C4 moves (c4_backend.py:49-50 returns a set): table[mask][i] is
 * the i-th column for the playable-column bit mask, 255-terminated.  A CPython 3.12 table is
 * built in; the Python host re-derives it from the running interpreter at import. */
int zc_c4_set_move_order(const uint8_t *table /* [128][8] */);
int zc_c4_get_move_order(uint8_t *table /* [128][8] */);

/* ---- search handle: replaces get_move() for a whole batch of trees ----------------------- */

/* arena_slots_per_tree: 16-byte slots per tree, 0 = default (C4: exact worst case, 11 slots per node;
 * chess: (max_sims+1) * 24 -- leaves are 3-slot stubs, about 5 slots per node in practice; an arena that
 * overflows makes zc_search_results return ZC_ECAPACITY).  Memory = max_trees * arena_slots_per_tree * 16 B. */
int zc_search_create(int game, int device, int max_trees, int max_sims, int64_t arena_slots_per_tree,
                     zc_search **out);
int zc_search_destroy(zc_search *h);
int64_t zc_search_device_bytes(const zc_search *h);

/* Upload n root states (host memory) and build the root nodes (backend.get_legal_moves(root),
 * mcts.cpp:104-109).  states: zc_c4_state[n] or zc_chess_state[n] according to the game. */
int zc_search_set_roots(zc_search *h, const void *host_states, int n, void *stream);
/* same, states already in device memory */
int zc_search_set_roots_dev(zc_search *h, const void *dev_states, int n, void *stream);

/* The whole of get_move's simulation loop (mcts.cpp:129-149) for all trees with a built-in
 * evaluator, in one persistent kernel: `simulations` sims per tree, frozen-statistics batches of
 * `batch_size` (1..32, the reference default is 32), UCB1 constant c.  Asynchronous on stream. */
int zc_search_run(zc_search *h, int simulations, double c, int batch_size, int evaluator, int policy,
                  uint64_t seed, void *stream);

/* ---- optional selection rule: PUCT with stored priors and virtual loss -----------------------------------------
 * The reference selects with UCB1 over frozen batches (mcts.cpp:41-63) and that is the default (ZC_SELECT_UCB1, the
 * bit-exact path).  ZC_SELECT_PUCT is the AlphaZero-style rule BASELINE.json's north star names; the reference has
 * no counterpart, so its definition is this library's (zeroclone_b200/csrc/puct.cuh) and its checker is
 * oracle/zc_oracle.c:zo_search_puct (whole-tree hashes must agree):
 *   a* = argmax_a  Wa/Na + c * P(a) * sqrt(N + 1) / (1 + Na),   lowest index on ties;   on the way down N += 1,
 *   Na += 1, Wa -= virtual_loss;  one new node per simulation;  a batch of `batch_size` simulations is selected one
 *   after the other, evaluated together and backed up in order (Wa += virtual_loss; Wa -= result; result = -result).
 *   Priors P(a) = (1 + prior_weight * move_value(a)) / sum are stored with each node (one float per edge);
 *   move_value is the capture value the reference attaches to a chess move (chess_backend.cpp:50-64), 0 for Connect
 *   Four (uniform priors).  zc_search_set_root_priors replaces the roots' priors (policy head, exploration noise).
 * Call zc_search_set_mode BEFORE zc_search_set_roots*.  In PUCT mode `c` of zc_search_run / zc_search_begin is c_puct,
 * `policy` is ignored (the rule itself decides which move is expanded next), the rollout evaluator is not available,
 * and a leaf deeper than 63 plies below the root is ZC_ECAPACITY. */
#define ZC_SELECT_UCB1 0
#define ZC_SELECT_PUCT 1
int zc_search_set_mode(zc_search *h, int select_mode, double virtual_loss, int prior_weight);
/* host_priors: float[n_trees][stride], row t = the priors of tree t's root moves in backend order */
int zc_search_set_root_priors(zc_search *h, const float *host_priors, int stride, void *stream);

/* The same loop cut at the evaluator for ZC_EVAL_EXTERNAL (value.batch, mcts.cpp:112-127):
 *   zc_search_begin(...)
 *   while (zc_search_pending(h) > 0) {
 *       zc_search_select(h, planes, ...)      // select + expand one batch per tree, pack leaves
 *       values = network(planes)              // caller (PyTorch)
 *       zc_search_backprop(h, values, ...)    // backprop in pending order
 *   }
 * Leaf i of tree t is row t*batch_size+i of `dev_planes` / `dev_values`. */
int zc_search_begin(zc_search *h, int simulations, double c, int batch_size, int policy, uint64_t seed);
int zc_search_pending(const zc_search *h); /* simulations per tree still to run */
/* plane element types of zc_search_select / the tower's operand formats */
#define ZC_PLANE_BF16 0
#define ZC_PLANE_F32 1
#define ZC_PLANE_F16 2
/* dev_planes: [n_trees*batch_size][C][H][W] (C4: 2x6x7, c4_backend.py:52-61; chess: 17x8x8,
 * chess_backend.cpp:461-521), plane_dtype 0 = bf16, 1 = f32, 2 = f16.  Rows of batches shorter
 * than batch_size (the last batch) are zero-filled. */
int zc_search_select(zc_search *h, void *dev_planes, int plane_dtype, void *stream);
int zc_search_backprop(zc_search *h, const float *dev_values, void *stream);

/* Root readout: chosen move (mcts.cpp:150-155: most-visited child, lowest index on ties) and
 * per-child Na / Wa.  results: zc_root_result[n]; visits: int32[n][stride]; value_sums:
 * double[n][stride]; moves: zc_chess_move[n][stride] (C4: fr = column); the last three may be
 * NULL.  Host buffers; synchronises stream. */
int zc_search_results(zc_search *h, zc_root_result *results, int32_t *visits, double *value_sums,
                      zc_chess_move *moves, int stride, void *stream);

/* The same readout (chosen move and counters per tree, no per-child arrays) without stalling the host:
 * _begin enqueues the readout kernel on `stream` and the device->pinned-host copy on the handle's own copy
 * stream, so the next zc_search_set_roots_dev / zc_search_run can be enqueued immediately (device results are
 * double-buffered); _end waits for the most recent _begin and copies results[n] out (ZC_ECAPACITY as above).
 * This is how a self-play loop pipelines searches of a few milliseconds (scripts/train.py:151-170). */
int zc_search_results_begin(zc_search *h, void *stream);
int zc_search_results_end(zc_search *h, zc_root_result *results);

/* Order-dependent 64-bit hash over every node and edge of each tree (same function as
 * oracle/zc_oracle.c:hash_tree): equality means the whole tree matches the reference's bit for bit. */
int zc_search_tree_hash(zc_search *h, uint64_t *host_hashes, void *stream);

/* Copy the node arena of one tree to the host (16-byte slots; layout in zeroclone_b200/csrc/tree.cuh)
 * for inspection and invariant checks.  Writes min(used, max_slots) slots, *used = slots in use.
 * state_slots_out (may be NULL) receives the number of state slots per node (C4 1, chess 2). */
int zc_search_read_tree(zc_search *h, int tree, void *host_slots, int64_t max_slots, int64_t *used,
                        int32_t *state_slots_out, void *stream);

/* counters for roofline accounting, summed over trees since the last set_roots */
typedef struct zc_search_counters {
    int64_t simulations;
    int64_t nodes;
    int64_t sum_leaf_depth;
    int64_t sum_path_children; /* sum over descent steps of the children scanned */
    int64_t arena_slots_used;
    int64_t kernel_launches;   /* kernels launched by this handle since creation */
} zc_search_counters;
int zc_search_get_counters(zc_search *h, zc_search_counters *out, void *stream);

/* ---- game rules: the six-function backend interface (engine/README.md:13-28) --------------- */

/* Single-state helpers for the Python backend shims (root bookkeeping of Engine.play_move /
 * legal_moves / _evaluate, engine.py:94-108,148-157).  They run the same host+device rule code as
 * the kernels, on the host, one state at a time; they are not a fallback for the search. */
int zc_c4_init_state(zc_c4_state *out);                                    /* c4_backend.py:11-12 */
int zc_c4_legal_moves(const zc_c4_state *s, int32_t *cols /* [7] */);      /* :49-50, returns count, set order */
int zc_c4_play_move(const zc_c4_state *s, int col, zc_c4_state *out);      /* :14-23 */
int zc_c4_check_win(const zc_c4_state *s);                                 /* :25-44, 0/1 */
int zc_c4_check_draw(const zc_c4_state *s);                                /* :46-47, 0/1 */
int zc_c4_to_tensor(const zc_c4_state *s, float *out /* [2][6][7] */);     /* :52-61 */

int zc_chess_init_state(zc_chess_state *out);                              /* chess_backend.cpp:445-457 */
int zc_chess_from_fen(const char *fen, zc_chess_state *out);               /* :525-556 */
int zc_chess_legal_moves(const zc_chess_state *s, zc_chess_move *out /* [ZC_MAX_MOVES] */); /* :184-360, returns count */
int zc_chess_play_move(const zc_chess_state *s, const zc_chess_move *m, zc_chess_state *out); /* :364-400 */
int zc_chess_check_win(const zc_chess_state *s);                           /* :404-412 */
/* hist_*: the side's moves, most recent first (state.h:14, chess_backend.cpp:374) */
int zc_chess_check_draw(const zc_chess_state *s, const zc_chess_move *hist_white, int n_white,
                        const zc_chess_move *hist_black, int n_black);     /* :416-441 */
int zc_chess_to_tensor(const zc_chess_state *s, float *out /* [17][8][8] */); /* :461-521 */

/* Batched rules ON THE DEVICE (the kernels the search uses), host buffers in and out:
 * moves[n][ZC_MAX_MOVES], counts[n], flags[n]: bit 0 = check_win, bit 1 = no moves and not in check
 * (the stalemate branch of check_draw), bit 2 = side to move in check. */
int zc_chess_legal_moves_batch(int device, const zc_chess_state *states, int n, zc_chess_move *moves,
                               int32_t *counts, int32_t *flags);
/* the same through the warp-cooperative generator (one position per warp, one piece per lane) that the
 * search uses when a single node needs its move list */
int zc_chess_legal_moves_batch_warp(int device, const zc_chess_state *states, int n, zc_chess_move *moves,
                                    int32_t *counts, int32_t *flags);
/* perft as the reference's rules count it (number of legal moves at the last ply), breadth-first
 * on the device */
int zc_chess_perft(int device, const zc_chess_state *root, int depth, uint64_t *out);
/* C4 on the device: per state the ordered legal columns (255-padded to 8), flags bit 0 = check_win,
 * bit 1 = check_draw */
int zc_c4_rules_batch(int device, const zc_c4_state *states, int n, uint8_t *cols /* [n][8] */, int32_t *flags);

/* ---- device-resident self-play step (Engine.play_move + _evaluate for every game at once,
 * engine.py:98-108,148-153; the loop of scripts/train.py:151-170 keeps its roots on the device) ---- */

/* game results as Engine._evaluate reports them, plus "still running" */
#define ZC_RESULT_ONGOING 2

/* After a search: for every tree whose `dev_active[i]` is non-zero, take the chosen root move
 * (most-visited child, mcts.cpp:150-155), apply it to dev_states[i] (zc_c4_state / zc_chess_state
 * records in device memory, the array that was given to zc_search_set_roots_dev), and evaluate the new
 * position: check_win -> new.turn*2-1, check_draw -> 0, else ZC_RESULT_ONGOING.
 * Never silent: if a tree outgrew its arena during the search (ZC_ECAPACITY), an active tree has no visited
 * root move (ZC_ESTATE: no simulation ran or the root is terminal), or a chess side's history is full
 * (ZC_ECAPACITY: grow hist_cap), that tree's state is left unchanged with ZC_RESULT_ONGOING and the call
 * fails; zc_last_error() names the lowest offending tree.
 * Chess: dev_hist holds each side's moves [n][2][hist_cap] in playing order (zc_chess_move), dev_hist_len
 * [n][2]; both are updated (chess_backend.cpp:374) and used for the repetition rule (:148-180,434-438);
 * the fifty-ply counter lives in the state record.  C4: pass NULL for both.
 * host_results[n] (int32) and host_moves[n] receive the result and the move played (inactive trees:
 * ZC_RESULT_ONGOING and a zero move).  Synchronises stream. */
int zc_search_advance(zc_search *h, void *dev_states, const uint8_t *dev_active, zc_chess_move *dev_hist,
                      int32_t *dev_hist_len, int hist_cap, int32_t *host_results, zc_chess_move *host_moves,
                      void *stream);

/* state_to_tensor for many packed states at once (host buffers): out[n][C][H][W] float32 */
int zc_states_to_tensor(int game, const void *states, int n, float *out);

/* `policy_freedom` of Policy.immediate_value (policy_functions.py:16, the YAML key policy.policy_freedom); default 0 */
int zc_search_set_policy_freedom(zc_search *h, double policy_freedom);

/* ---- value-network evaluator: the whole residual tower as one fused kernel --------------------
 * Replaces the forward of engine/value_functions.py:78-99 (_batch_worker: model(batch)) over
 * models/chess_value/network.py:24-45 (stem conv3x3+BN+ReLU, n_blocks residual blocks, average
 * pool, Linear(128,1), tanh).  BatchNorm (eval mode) must already be folded into conv_w / conv_b.
 *   conv_w : float32, PyTorch layout [128][Cin][3][3] per convolution, stem first (Cin = 2 for
 *            Connect Four, 17 for chess), then 2*n_blocks convolutions with Cin = 128
 *   conv_b : float32 [1 + 2*n_blocks][128]
 *   head_w : float32 [128], head_b : Linear bias
 *   plane_dtype : operand format of the tensor-core path, ZC_PLANE_F16 or ZC_PLANE_BF16 (same rate).
 *            fp16 is what the reference itself evaluates in on a GPU (value_functions.py:6) and is the default of
 *            the Python host: with 11 significand bits the per-leaf error vs fp32 is ~5e-5 and root values of an
 *            800-simulation search agree with the fp32 reference within 1e-3 on chess as well as Connect Four;
 *            bf16 (8 bits, ~5e-4 per leaf) holds 1e-3 at the root on Connect Four only (tests/test_gpu_parity_bench_sets.py).
 *            Activations saturate at +-65504 instead of overflowing.
 * zc_tower_forward evaluates n_leaves positions packed as 16-bit planes [n][Cin][H][W] in the tower's format (the
 * layout zc_search_select writes with the same plane_dtype) into dev_values[n] float32, on `stream`, without
 * synchronising.  Arithmetic: 16-bit operands, fp32 accumulation, 16-bit activations between layers. */
typedef struct zc_tower zc_tower;
int zc_tower_create(int game, int device, int n_blocks, int plane_dtype, const float *conv_w, const float *conv_b,
                    const float *head_w, float head_b, zc_tower **out);
void zc_tower_destroy(zc_tower *t);
int zc_tower_forward(zc_tower *t, const void *dev_planes, int n_leaves, float *dev_values, void *stream);
int zc_tower_plane_dtype(const zc_tower *t);
/* New weights into an existing tower (same arguments as zc_tower_create): what the training loop does after
 * every cycle (scripts/train.py:143-146 saves latest.pth and the next self-play evaluates with it).  The weight
 * image is re-packed on the host and uploaded on `stream`, ordered after the forwards already enqueued there. */
int zc_tower_update_weights(zc_tower *t, const float *conv_w, const float *conv_b, const float *head_w, float head_b,
                            void *stream);
/* After a failed launch or synchronisation: the kernel's own fault word (0 = none; else 0x80000000 | wait tag,
 * written when a bounded barrier wait timed out -- a protocol bug, never expected).  Clears it. */
unsigned int zc_tower_fault(zc_tower *t);
int64_t zc_tower_launches(const zc_tower *t);

#ifdef __cplusplus
}
#endif
#endif /* ZC_B200_H */
