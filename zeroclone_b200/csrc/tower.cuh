// Value-network forward (models/chess_value/network.py:24-45 of the reference: stem conv3x3 + BN + ReLU,
// 8 residual blocks of two conv3x3 + BN, global average pool, Linear(128,1), tanh) as ONE persistent
// sm_100a kernel.  The reference evaluates leaves through PyTorch (engine/value_functions.py:78-99); here
// every leaf's activations stay in shared memory for the whole tower and only weights stream in.
//
// Mapping
//   - a tile = NB boards = 128 GEMM rows (Connect Four: 3 boards x 42 cells = 126 rows; chess: 2 x 64);
//     row p = y*(NB*W) + b*W + x, so a vertical tap is a shift of +-NB*W rows that falls off the tile
//     into a zero halo, and a horizontal tap is a shift of +-1 row;
//   - activations live in shared memory K-major WITHOUT swizzle, chunk-major: the 16-byte chunk c
//     (8 channels) of row r sits at c*A_LBO + r*16.  With SBO = 128 B every row is 16 B after the
//     previous one, so "the tile shifted by s rows" is the same buffer with the UMMA descriptor's start
//     address moved by 16*s: the nine taps of a 3x3 convolution are nine descriptors over ONE resident
//     copy -- no im2col gather, no re-read of activations from L2 (which bounds a generic implicit-GEMM);
//   - a horizontal shift wraps into the neighbouring board at x = 0 / W-1.  The three tap columns
//     (dx = -1, 0, +1) therefore accumulate into three TMEM accumulators and the epilogue adds the
//     dx = -1 (dx = +1) accumulator only to rows with x != 0 (x != W-1);
//   - tcgen05.mma kind::f16 (fp16 x fp16 or bf16 x bf16 -> fp32: template parameter F16), K = 16, issued by one elected thread from warp-uniform
//     control flow; per layer and tile 9 taps x 8 k-steps.  CTAs run as pairs (cta_group::2, M = 256 over two
//     SMs): each CTA keeps its own 128-row tile and HALF of every tap's weights (64 of the 128 output
//     channels), rank 0 issues for both -- the single-CTA form is bound by shared-memory bandwidth;
//   - weights: one 16 KB image per (layer, tap, CTA rank) already in the shared-memory layout of the B operand,
//     fetched with cp.async.bulk into an 8-stage ring (mbarrier complete_tx); rank 1 relays "my half has
//     landed" to rank 0 with a remote mbarrier arrive, tcgen05.commit multicast frees a stage in both CTAs;
//   - epilogue (8 warps): phase 1 reads the two side accumulators, applies the edge masks and hands their
//     TMEM columns back at once; phase 2 reads the centre accumulator, adds bias (shared memory) and the
//     residual (kept in registers as packed bf16 for the whole block), ReLU, and overwrites the tile in place.
//     Two tiles per CTA alternate so the epilogue of one runs under the MMAs of the other; accumulators
//     rotate through the 4 x 128 TMEM columns;
//   - after the last block: per-row dot with the head weights, per-board sum in a fixed order, tanh;
//   - every mbarrier wait is bounded (~2 s): a protocol bug ends the kernel with a fault word the host reports
//     (zc_last_error names the wait) instead of hanging the GPU.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace zc {
namespace tower {

constexpr int CH = 128;                       // tower width (network.py:26)
constexpr int KCHUNKS = CH / 8;               // 16-byte chunks per activation row
constexpr int HALO = 24;                      // zero rows above and below the tile (>= NB*W + 1)
constexpr int MROWS = 128;                    // UMMA M
constexpr int BUF_ROWS = HALO + MROWS + HALO;
constexpr int A_LBO = BUF_ROWS * 16;          // byte distance between k-chunks of the activation buffer
constexpr int A_BUF_BYTES = KCHUNKS * A_LBO;  // 45056
constexpr int B_LBO = CH * 16;                // weights: N = 128 rows of 16 B per k-chunk
constexpr int W_TAP_BYTES = KCHUNKS * B_LBO;  // 32768
constexpr int RING_BYTES = 4 * W_TAP_BYTES;    // weight ring: 8 stages of the half of a tap image a CTA of a pair holds
constexpr int NT = 2;                         // tiles in flight per CTA
constexpr int N_EPI_WARPS = 8;
constexpr int N_THREADS = (2 + N_EPI_WARPS) * 32;
constexpr int SMEM_BARS = 0;                  // mbarriers + tmem pointer
constexpr int SMEM_PART = 256;                // head partial sums [NT][2][128] float
constexpr int SMEM_ABUF = SMEM_PART + NT * 2 * MROWS * 4;   // 2304
constexpr int SMEM_WRING = SMEM_ABUF + NT * A_BUF_BYTES;
constexpr int SMEM_BIAS = SMEM_WRING + RING_BYTES;    // biases of every layer, fp32
constexpr int MAX_LAYERS = 17;                                  // network.py:33: stem + 8 blocks
constexpr int SMEM_TOTAL = SMEM_BIAS + MAX_LAYERS * CH * 4;     // 232192
static_assert(SMEM_ABUF % 16 == 0 && SMEM_WRING % 16 == 0, "descriptor start addresses are in 16-byte units");
static_assert(SMEM_TOTAL <= 227 * 1024, "shared memory budget");

struct Params {
    const uint16_t* planes;        // [n_leaves][CIN][H][W], 16-bit values in the tower's operand format (bf16 or fp16)
    const uint8_t* wimg2;          // [n_layers][9 taps][2 halves of N][KCHUNKS][64][8]: what each CTA of a pair holds
    const float* bias;             // [n_layers][128] (BatchNorm folded)
    const float* head_w;           // [128]
    float head_b;
    float* out;                    // [n_leaves]
    int n_leaves;
    int n_layers;                  // 1 + 2*blocks
    unsigned int* fault;           // host-mapped word, set when a barrier wait times out (readable after the trap)
};

// ------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded wait: a protocol bug must end the kernel with an error, never hang the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, unsigned int* fault, int tag) {
    uint32_t ok = 0;
    long long t0 = 0;
    for (uint32_t it = 0;; ++it) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (ok) return;
        if ((it & 1023u) == 1023u) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            else if (now - t0 > 4000000000ll) {    // ~2 s at 1.9 GHz; a whole launch takes milliseconds
                atomicExch(fault, 0x80000000u | (unsigned)tag);
                __threadfence_system();
                __trap();
            }
        }
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
// cta_group::2: one instruction drives the tensor cores of both CTAs of a pair (M = 256: each CTA its own
// 128 rows of A from its own shared memory at the descriptor's offset, and half of B's N rows)
__device__ __forceinline__ void tc_mma2(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives on the barrier at this offset in both CTAs of the pair once the MMAs issued so far are complete
__device__ __forceinline__ void tc_commit2(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((uint16_t)3)
                 : "memory");
}
// address of the same shared-memory offset in CTA `rank` of the cluster, and an arrive on a barrier there
__device__ __forceinline__ uint32_t map_to_cta(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, K-major, SWIZZLE_NONE (cute/arch/mma_sm100_desc.hpp SmemDescriptor):
// [0,14) start>>4, [16,30) leading byte offset>>4 (between the two k-chunks of one MMA),
// [32,46) stride byte offset>>4 (between 8-row groups), [46,48) version = 1, [61,64) layout = 0.
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}
// instruction descriptor (InstrDescriptor): c = f32, a = b = bf16, both K-major, N = 128
// (a/b format field: 0 = fp16, 1 = bf16 -- same tensor-core rate; fp16 carries 3 more mantissa bits)
template <bool F16>
struct Idesc2 {
    static constexpr uint32_t value = (1u << 4) | ((F16 ? 0u : 1u) << 7) | ((F16 ? 0u : 1u) << 10) | ((128u >> 3) << 17) | ((256u >> 4) << 24);   // M = 256 over the pair
};

// two fp32 lanes per instruction (FADD2 / FMUL2 / FFMA2): the epilogue's arithmetic in half the issue slots
__device__ __forceinline__ uint64_t f2_pack(uint32_t lo, uint32_t hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
    return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float& lo, float& hi) {
    uint32_t a, b;
    asm("mov.b64 {%0, %1}, %2;" : "=r"(a), "=r"(b) : "l"(v));
    lo = __uint_as_float(a);
    hi = __uint_as_float(b);
}
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xFFFF0000u); }
// the activation format of the tower: two 16-bit values per register, bf16 or fp16 (F16)
template <bool F16>
__device__ __forceinline__ uint32_t act_pack(float lo, float hi) {
    if constexpr (F16) {
        uint32_t r;                       // saturating: an activation beyond 65504 stays finite
        asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
        return r;
    } else {
        return pack_bf16(lo, hi);
    }
}
template <bool F16>
__device__ __forceinline__ float2 act_unpack(uint32_t u) {
    if constexpr (F16) return __half22float2(*reinterpret_cast<const __half2*>(&u));
    else return make_float2(bf16_lo(u), bf16_hi(u));
}

template <int H_, int W_, int NB_, int CIN_>
struct Geom {
    static constexpr int H = H_, W = W_, NB = NB_, CIN = CIN_;
    static constexpr int HW = H * W;
    static constexpr int RS = NB * W;              // rows per board line = vertical shift
    static constexpr int ROWS = H * RS;            // real rows of a tile
    static constexpr int CIN16 = (CIN + 15) / 16 * 16;
    static_assert(ROWS <= MROWS && RS + 1 <= HALO, "tile does not fit");
};
using GeomC4 = Geom<6, 7, 3, 2>;      // c4_backend.py:52-61
using GeomChess = Geom<8, 8, 2, 17>;  // chess_backend.cpp:461-521

// ------------------------------------------------------------------------------------ the kernel
template <class G, bool F16>
__global__ void __launch_bounds__(N_THREADS, 1) k_value_tower(const Params p) {
    constexpr uint32_t IDESC2 = Idesc2<F16>::value;
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // barriers
    constexpr int NSTAGE = 8;
    constexpr uint32_t STAGE_BYTES = RING_BYTES / NSTAGE;        // this CTA's half of a tap image
    constexpr uint32_t BLBO = B_LBO / 2;                         // k-chunk stride of the B operand held here
    const uint32_t bar_wfull = sbase + SMEM_BARS, bar_wempty = bar_wfull + 8 * NSTAGE, bar_aready = bar_wempty + 8 * NSTAGE,
                   bar_accfull = bar_aready + 8 * NT, bar_pfree = bar_accfull + 8 * NT, bar_wpeer = bar_pfree + 8 * NT;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + SMEM_BARS + 8 * (3 * NSTAGE + 3 * NT));
    static_assert(8 * (3 * 8 + 3 * NT) + 4 <= SMEM_PART, "barrier area");
    float* part = reinterpret_cast<float*>(smem + SMEM_PART);

    // this CTA's share of board groups and the (identical) step sequence every role walks
    const int n_groups = (p.n_leaves + G::NB - 1) / G::NB;
    const int cta = blockIdx.x, grid = gridDim.x;
    const int nj = cta < n_groups ? (n_groups - cta + grid - 1) / grid : 0;
    // groups per tile slot, padded to the longest CTA: the CTAs of a cluster share one weight stream
    const int ns = ((n_groups + grid - 1) / grid + NT - 1) / NT;
    const uint32_t crank = cluster_ctarank();               // 0 issues the pair's MMAs
    const int NL = p.n_layers;
    const int steps_per_slot = ns * NL;

    // zero both activation buffers (halo rows stay zero for the whole kernel)
    for (int i = threadIdx.x; i < NT * A_BUF_BYTES / 16; i += N_THREADS)
        reinterpret_cast<uint4*>(smem + SMEM_ABUF)[i] = make_uint4(0, 0, 0, 0);
    for (int i = threadIdx.x; i < NL * CH; i += N_THREADS) reinterpret_cast<float*>(smem + SMEM_BIAS)[i] = p.bias[i];
    fence_proxy_async();
    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTAGE; ++s) {
            mbar_init(bar_wfull + 8 * s, 1);
            mbar_init(bar_wempty + 8 * s, 1);
            mbar_init(bar_wpeer + 8 * s, 1);      // pair: the other CTA's half of a stage has landed
        }
        for (int t = 0; t < NT; ++t) {
            mbar_init(bar_aready + 8 * t, 2 * N_EPI_WARPS);   // one arrival per epilogue warp; both CTAs report to rank 0
            mbar_init(bar_accfull + 8 * t, 1);
            mbar_init(bar_pfree + 8 * t, 2 * N_EPI_WARPS);    // side accumulators of a step have been read
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync();   // the peer's barriers and zeroed buffers exist before anything touches them
    tc_fence_after();
    const uint32_t tmem = *tmem_ptr_smem;

    if (warp == 0) {
        // ============================== weight producer ==============================
        if (lane == 0) {
            uint32_t cnt = 0;
            for (int m = 0; m < steps_per_slot; ++m) {
                const int layer = m % NL;
                const uint32_t bytes = layer == 0 ? (G::CIN16 / 8) * BLBO : STAGE_BYTES;
                for (int slot = 0; slot < NT; ++slot)
                    for (int tap = 0; tap < 9; ++tap, ++cnt) {
                        const uint32_t st = cnt % NSTAGE, ph = (cnt / NSTAGE) & 1u;
                        mbar_wait(bar_wempty + 8 * st, ph ^ 1u, p.fault, 1);
                        mbar_expect_tx(bar_wfull + 8 * st, bytes);
                        const uint8_t* src = p.wimg2 + ((size_t)(layer * 9 + tap) * 2 + crank) * STAGE_BYTES;
                        bulk_g2s(sbase + SMEM_WRING + st * STAGE_BYTES, src, bytes, bar_wfull + 8 * st);
                    }
            }
        }
    } else if (warp == 1) {
        // ============================== MMA issuer ==============================
        // The whole warp walks the loop (warp-uniform control flow keeps descriptors in uniform
        // registers); one elected lane issues every tcgen05.mma / tcgen05.commit.
        uint32_t leader;
        asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
        const uint64_t desc_a_hi = smem_desc(0, A_LBO, 128), desc_b_hi = smem_desc(0, BLBO, 128);
        uint32_t wcnt = 0, gcnt = 0;
        if (crank != 0) {
            // the pair's second CTA issues nothing; it tells rank 0 when its half of each weight stage has landed
            const uint32_t total = (uint32_t)steps_per_slot * NT * 9u;
            for (; wcnt < total; ++wcnt) {
                const uint32_t st = wcnt % NSTAGE, ph = (wcnt / NSTAGE) & 1u;
                mbar_wait(bar_wfull + 8 * st, ph, p.fault, 6);
                if (leader) mbar_arrive_cluster(map_to_cta(bar_wpeer + 8 * st, 0));
                __syncwarp();
            }
        } else
        for (int m = 0; m < steps_per_slot; ++m) {
            const int layer = m % NL;
            const int ksteps = layer == 0 ? G::CIN16 / 16 : CH / 16;
#pragma unroll 1
            for (int slot = 0; slot < NT; ++slot) {
                const int n = m * NT + slot;
                mbar_wait(bar_aready + 8 * slot, (uint32_t)m & 1u, p.fault, 2);   // input of this step is in place
                tc_fence_after();
                const uint32_t abuf = sbase + SMEM_ABUF + slot * A_BUF_BYTES + HALO * 16;
#pragma unroll 1
                for (int g = 0; g < 3; ++g, ++gcnt) {
                    // Accumulator gcnt % 4 was last used three groups ago, by the previous step: group 1 takes over
                    // that step's dx = -1 accumulator, which its epilogue reads first and releases early (pfree);
                    // group 2 takes over its dx = 0 accumulator, free when that epilogue is complete, which it
                    // reports through the a_ready of the step after this one.  (Group 0 reuses an accumulator of
                    // the step before the previous one, released long ago.)
                    if (NT == 2 && n >= 1) {
                        if (g == 1) {
                            const int n0 = n - 1;
                            mbar_wait(bar_pfree + 8 * (n0 % NT), (uint32_t)(n0 / NT) & 1u, p.fault, 3);
                            tc_fence_after();
                        } else if (g == 2) {
                            const int n1 = n + 1;
                            mbar_wait(bar_aready + 8 * (n1 % NT), (uint32_t)(n1 / NT) & 1u, p.fault, 8);
                            tc_fence_after();
                        }
                    }
                    const uint32_t acc = tmem + (gcnt & 3u) * 128u;
#pragma unroll 1
                    for (int dyi = 0; dyi < 3; ++dyi, ++wcnt) {
                        const uint32_t st = wcnt % NSTAGE, ph = (wcnt / NSTAGE) & 1u;
                        mbar_wait(bar_wfull + 8 * st, ph, p.fault, 4);
                        mbar_wait(bar_wpeer + 8 * st, ph, p.fault, 7);
                        tc_fence_after();
                        const int shift = (dyi - 1) * G::RS + (g - 1);
                        const uint64_t ad = desc_a_hi | (uint64_t)(((abuf + shift * 16) >> 4) & 0x3FFFu);
                        const uint64_t bd = desc_b_hi | (uint64_t)(((sbase + SMEM_WRING + st * STAGE_BYTES) >> 4) & 0x3FFFu);
                        if (leader) {
#pragma unroll
                            for (int k = 0; k < CH / 16; ++k)
                                if (k < ksteps)
                                    tc_mma2(acc, ad + (uint64_t)(k * (2 * A_LBO >> 4)), bd + (uint64_t)(k * (2 * BLBO >> 4)), IDESC2,
                                            (uint32_t)((dyi | k) != 0));
                            // stage reusable (in both CTAs of the pair) once these MMAs have read it
                            tc_commit2(bar_wempty + 8 * st);
                        }
                        __syncwarp();
                    }
                }
                if (leader) tc_commit2(bar_accfull + 8 * slot);
                __syncwarp();
            }
        }
    } else {
        // ============================== epilogue warps ==============================
        const int ew = warp - 2;                 // 0..7
        const int quarter = warp & 3;            // TMEM lane quarter this warp may access
        const int half = ew >> 2;                // which 64 output channels
        const int r = quarter * 32 + lane;       // tile row = TMEM lane
        const bool valid = r < G::ROWS;
        const int y = r / G::RS, b = (r / G::W) % G::NB, x = r % G::W;
        const uint32_t tlane = tmem + ((uint32_t)(quarter * 32) << 16);
        uint32_t xreg[NT][32];                   // residual stream of this row, packed bf16 (64 channels)
#pragma unroll
        for (int i = 0; i < 32; ++i) xreg[0][i] = xreg[1][i] = 0u;

        auto load_planes = [&](int slot, int it) {
            const int j = it * NT + slot;
            const long long leaf = (long long)(cta + (long long)j * grid) * G::NB + b;
            const bool live = valid && j < nj && leaf < p.n_leaves;
            uint8_t* abuf = smem + SMEM_ABUF + slot * A_BUF_BYTES + (HALO + r) * 16;
            if (valid) {
                for (int chunk = half; chunk < G::CIN16 / 8; chunk += 2) {
                    uint32_t w[4] = {0, 0, 0, 0};
                    if (live) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const int ch = chunk * 8 + e;
                            if (ch < G::CIN) {
                                const unsigned short v = p.planes[((size_t)leaf * G::CIN + ch) * G::HW + y * G::W + x];
                                w[e >> 1] |= (uint32_t)v << ((e & 1) * 16);
                            }
                        }
                    }
                    *reinterpret_cast<uint4*>(abuf + chunk * A_LBO) = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
        };

        // the MMA issuer (rank 0 of a pair) learns from both CTAs that a tile's input is in place
        const uint32_t aready_dst = crank != 0 ? map_to_cta(bar_aready, 0) : bar_aready;
        const uint32_t pfree_dst = crank != 0 ? map_to_cta(bar_pfree, 0) : bar_pfree;
        auto signal = [&](uint32_t bar) {   // every lane has fenced its own accesses; one lane reports for the warp
            __syncwarp();
            if (lane == 0) {
                if (crank != 0) mbar_arrive_cluster(bar);
                else mbar_arrive(bar);
            }
        };
        auto signal_ready = [&](int slot) { signal(aready_dst + 8 * slot); };
        for (int slot = 0; slot < NT; ++slot) {
            load_planes(slot, 0);
            fence_proxy_async();
            signal_ready(slot);
        }

        const float fm = x != 0 ? 1.f : 0.f, fp = x != G::W - 1 ? 1.f : 0.f;   // horizontal taps that stay on the board
        const uint64_t fm2 = f2_pack(__float_as_uint(fm), __float_as_uint(fm)), fp2 = f2_pack(__float_as_uint(fp), __float_as_uint(fp));
        uint32_t gbase = 0;   // first accumulator group of the current step
        for (int m = 0; m < steps_per_slot; ++m) {
            const int layer = m % NL, it = m / NL;
            const bool last = layer == NL - 1;
            const float fres = (layer != 0 && (layer & 1) == 0) ? 1.f : 0.f;   // second conv of a block adds its input (network.py:20-21)
            const bool keep = (layer & 1) == 0;                                // output is the next block's input
            const float4* sb = reinterpret_cast<const float4*>(smem + SMEM_BIAS) + (layer * CH + half * 64) / 4;
            const uint64_t fres2 = f2_pack(__float_as_uint(fres), __float_as_uint(fres));
#pragma unroll
            for (int slot = 0; slot < NT; ++slot, gbase += 3) {
                mbar_wait(bar_accfull + 8 * slot, (uint32_t)m & 1u, p.fault, 5);
                tc_fence_after();
                uint8_t* arow = smem + SMEM_ABUF + slot * A_BUF_BYTES + (HALO + r) * 16;
                // phase 1: the two side accumulators (dx = -1, +1), masked at the board edges, into registers;
                // their TMEM columns go back to the MMA issuer before the rest of the epilogue runs
                uint64_t side[32];                 // pairs of channels
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    const uint32_t col = half * 64 + cc * 16;
                    uint32_t am[16], ap[16];
                    tc_ld16(tlane + ((gbase + 0) & 3u) * 128u + col, am);
                    tc_ld16(tlane + ((gbase + 2) & 3u) * 128u + col, ap);
                    tc_wait_ld();
#pragma unroll
                    for (int i = 0; i < 16; i += 2)
                        side[cc * 8 + (i >> 1)] = f2_fma(fp2, f2_pack(ap[i], ap[i + 1]), f2_mul(fm2, f2_pack(am[i], am[i + 1])));
                }
                tc_fence_before();
                signal(pfree_dst + 8 * slot);
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    const uint32_t col = half * 64 + cc * 16;
                    uint32_t a0[16];
                    tc_ld16(tlane + ((gbase + 1) & 3u) * 128u + col, a0);
                    tc_wait_ld();
                    uint32_t o[8];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 bb = sb[cc * 4 + q];
                        const float bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
                        for (int e = 0; e < 4; e += 2) {
                            const int i = q * 4 + e;
                            const uint32_t xr = xreg[slot][cc * 8 + (i >> 1)];
                            uint64_t v = f2_add(f2_pack(a0[i], a0[i + 1]), side[cc * 8 + (i >> 1)]);
                            v = f2_add(v, f2_pack(__float_as_uint(bv[e]), __float_as_uint(bv[e + 1])));
                            const float2 xf = act_unpack<F16>(xr);
                            v = f2_fma(fres2, f2_pack(__float_as_uint(xf.x), __float_as_uint(xf.y)), v);
                            float v0, v1;
                            f2_unpack(v, v0, v1);
                            const uint32_t pk = act_pack<F16>(fmaxf(v0, 0.f), fmaxf(v1, 0.f));
                            o[i >> 1] = pk;
                            xreg[slot][cc * 8 + (i >> 1)] = keep ? pk : xr;
                        }
                    }
                    if (valid && !last) {
                        const int chunk = half * 8 + cc * 2;
                        *reinterpret_cast<uint4*>(arow + chunk * A_LBO) = make_uint4(o[0], o[1], o[2], o[3]);
                        *reinterpret_cast<uint4*>(arow + (chunk + 1) * A_LBO) = make_uint4(o[4], o[5], o[6], o[7]);
                    }
                }
                if (last) {
                    // head: mean over the board's cells, Linear(128,1), tanh (network.py:36-39), fixed summation order.
                    // The last layer closes a block, so xreg holds this row's final activations.
                    float dot = 0.f;
#pragma unroll
                    for (int i = 0; i < 32; ++i) {
                        const float2 hw = __ldg(reinterpret_cast<const float2*>(p.head_w + half * 64) + i);
                        const float2 xf = act_unpack<F16>(xreg[slot][i]);
                        dot = fmaf(xf.x, hw.x, dot);
                        dot = fmaf(xf.y, hw.y, dot);
                    }
                    float* pt = part + slot * 2 * MROWS;
                    pt[half * MROWS + r] = valid ? dot : 0.f;
                    asm volatile("bar.sync 1, %0;" ::"n"(N_EPI_WARPS * 32) : "memory");
                    const int et = threadIdx.x - 64;
                    if (et < G::NB) {
                        const int j = it * NT + slot;
                        const long long leaf = (long long)(cta + (long long)j * grid) * G::NB + et;
                        if (j < nj && leaf < p.n_leaves) {
                            float s = 0.f;
                            for (int yy = 0; yy < G::H; ++yy)
                                for (int xx = 0; xx < G::W; ++xx) {
                                    const int rr = yy * G::RS + et * G::W + xx;
                                    s += pt[rr] + pt[MROWS + rr];
                                }
                            p.out[leaf] = tanhf(s * (1.0f / G::HW) + p.head_b);
                        }
                    }
                    if (it + 1 < ns) load_planes(slot, it + 1);
                }
                fence_proxy_async();
                tc_fence_before();
                signal_ready(slot);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync();   // no CTA leaves while its peer may still use its shared memory, barriers or TMEM
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512) : "memory");
    }
}

}  // namespace tower
}  // namespace zc
