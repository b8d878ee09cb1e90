// zc_tower_*: the value-network evaluator of the hot path (engine/value_functions.py:61-129 calling
// models/chess_value/network.py:24-45) as one fused sm_100a kernel.  Included by zc_api.cu.
#include "tower.cuh"

struct zc_tower {
    int game = 0, device = 0, n_layers = 0, cin = 0;
    bool f16 = true;                     // operand format: fp16 (default) or bf16
    uint8_t* wimg2 = nullptr;
    float* bias = nullptr;
    float* head_w = nullptr;
    float head_b = 0.f;
    unsigned int* fault = nullptr;       // device alias of fault_host
    unsigned int* fault_host = nullptr;  // pinned + mapped: still readable after the kernel trapped
    uint16_t* wstage = nullptr;          // pinned staging of the weight image (zc_tower_update_weights)
    float* bstage = nullptr;             // pinned staging: biases, then head weights
    size_t wimg_elems = 0;
    int n_sms = 0;
    int64_t launches = 0;
};

static inline uint16_t f32_to_bf16_rne(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7F800000u) == 0x7F800000u) return (uint16_t)((u >> 16) | ((u & 0xFFFFu) ? 0x40u : 0u));
    u += 0x7FFFu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}

// fp32 -> fp16, round to nearest even, saturating to +-65504 (no infinities enter the tower)
static inline uint16_t f32_to_f16_rne_sat(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    const uint32_t sign = (u >> 16) & 0x8000u;
    u &= 0x7FFFFFFFu;
    if (u > 0x7F800000u) return (uint16_t)(sign | 0x7E00u);            // NaN
    if (u >= 0x477FF000u) return (uint16_t)(sign | 0x7BFFu);           // >= 65520 rounds past the largest finite half
    if (u < 0x33000001u) return (uint16_t)sign;                        // < 2^-25: rounds to zero
    int e = (int)(u >> 23) - 127;
    uint32_t m = (u & 0x7FFFFFu) | 0x800000u;
    int shift = e < -14 ? 13 + (-14 - e) : 13;                         // subnormal halves lose further bits
    uint32_t h = m >> shift, rem = m & ((1u << shift) - 1u), halfway = 1u << (shift - 1);
    if (rem > halfway || (rem == halfway && (h & 1u))) ++h;
    // h holds the 11-bit significand (hidden bit at 0x400) or a subnormal; the exponent is added arithmetically so a
    // rounding carry moves into it
    const uint32_t out = e < -14 ? h : (uint32_t)((e + 15 - 1) << 10) + h;
    return (uint16_t)(sign | out);
}

// Weight images in the shared-memory layout of the B operand, split by halves of N:
// [layer][tap][n / 64][k-chunk][n % 64][8] bf16 -- the part each CTA of a cta_group::2 pair keeps in its own
// shared memory; taps ordered by column offset first (dx = -1, 0, +1), then row offset (dy = -1, 0, +1).
// conv_w is PyTorch's [Cout][Cin][kH][kW] per layer, stem first (cross-correlation: dy = kh-1, dx = kw-1).
static void pack_tower_weights(const float* conv_w, int nl, int cin, bool f16, uint16_t* img2) {
    using namespace zc::tower;
    memset(img2, 0, (size_t)nl * 9 * KCHUNKS * CH * 8 * sizeof(uint16_t));
    size_t woff = 0;
    for (int l = 0; l < nl; ++l) {
        const int ci = l == 0 ? cin : CH;
        for (int g = 0; g < 3; ++g)
            for (int dyi = 0; dyi < 3; ++dyi) {
                uint16_t* dst2 = img2 + ((size_t)l * 9 + g * 3 + dyi) * KCHUNKS * CH * 8;
                for (int n = 0; n < CH; ++n)
                    for (int k = 0; k < ci; ++k) {
                        const float wv = conv_w[woff + (((size_t)n * ci + k) * 3 + dyi) * 3 + g];
                        const uint16_t v = f16 ? f32_to_f16_rne_sat(wv) : f32_to_bf16_rne(wv);
                        dst2[(((size_t)(n / 64) * KCHUNKS + k / 8) * 64 + n % 64) * 8 + (k % 8)] = v;
                    }
            }
        woff += (size_t)CH * ci * 9;
    }
}

// every live tower's fault word: a failed CUDA call anywhere in the library reports a tower protocol fault by name
static std::vector<zc_tower*> g_towers;
static const char* tower_wait_name(unsigned tag) {
    switch (tag) {
    case 1: return "weight producer waiting for a free ring stage";
    case 2: return "MMA issuer waiting for a tile's input";
    case 3: return "MMA issuer waiting for the side accumulators to be read";
    case 4: return "MMA issuer waiting for a weight stage";
    case 5: return "epilogue waiting for a full accumulator";
    case 6: return "peer CTA waiting for its half of a weight stage";
    case 7: return "MMA issuer waiting for the peer's half of a weight stage";
    case 8: return "MMA issuer waiting for the centre accumulator to be read";
    default: return "unknown wait";
    }
}
static std::string tower_fault_note() {
    std::string note;
    for (zc_tower* t : g_towers) {
        const unsigned f = t->fault_host ? *(volatile unsigned int*)t->fault_host : 0u;
        if (f) note += " [k_value_tower fault on device " + std::to_string(t->device) + ": bounded wait timed out, tag " +
                       std::to_string(f & 0xFFFFu) + " (" + tower_wait_name(f & 0xFFFFu) + ")]";
    }
    return note;
}

extern "C" int zc_tower_create(int game, int device, int n_blocks, int plane_dtype, const float* conv_w, const float* conv_b,
                               const float* head_w, float head_b, zc_tower** out) {
    if (!out) return fail(ZC_EINVAL, "out is NULL");
    *out = nullptr;
    if (game != ZC_GAME_C4 && game != ZC_GAME_CHESS) return fail(ZC_EINVAL, "unknown game");
    if (plane_dtype != ZC_PLANE_BF16 && plane_dtype != ZC_PLANE_F16) return fail(ZC_EINVAL, "tower operands are ZC_PLANE_F16 or ZC_PLANE_BF16");
    if (n_blocks < 1 || 1 + 2 * n_blocks > zc::tower::MAX_LAYERS || !conv_w || !conv_b || !head_w) return fail(ZC_EINVAL, "bad tower arguments");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) return fail(ZC_ENODEVICE, "no CUDA device: libzc_b200 has no CPU path");
    if (device < 0 || device >= ndev) return fail(ZC_EINVAL, "device out of range");
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(ZC_ENODEVICE, "zc_tower needs an sm_100 device (tcgen05/TMEM)");

    using namespace zc::tower;
    const int cin = game == ZC_GAME_C4 ? GeomC4::CIN : GeomChess::CIN;
    const int nl = 1 + 2 * n_blocks;
    zc_tower* t = new zc_tower;
    t->wimg_elems = (size_t)nl * 9 * KCHUNKS * CH * 8;
    t->game = game;
    t->f16 = plane_dtype == ZC_PLANE_F16;
    t->device = device;
    t->n_layers = nl;
    t->cin = cin;
    t->head_b = head_b;
    t->n_sms = prop.multiProcessorCount;
    auto cleanup = [&](int rc) {
        cudaFree(t->wimg2); cudaFree(t->bias); cudaFree(t->head_w);
        cudaFreeHost(t->fault_host); cudaFreeHost(t->wstage); cudaFreeHost(t->bstage);
        delete t;
        return rc;
    };
#define TOWER_TRY(expr)                                                                                    \
    do {                                                                                                   \
        cudaError_t _e = (expr);                                                                           \
        if (_e != cudaSuccess) return cleanup(fail(ZC_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(_e))); \
    } while (0)
    TOWER_TRY(cudaMalloc(&t->wimg2, t->wimg_elems * 2));
    TOWER_TRY(cudaMalloc(&t->bias, sizeof(float) * nl * CH));
    TOWER_TRY(cudaMalloc(&t->head_w, sizeof(float) * CH));
    TOWER_TRY(cudaHostAlloc(&t->fault_host, sizeof(unsigned int), cudaHostAllocMapped));
    *t->fault_host = 0u;
    TOWER_TRY(cudaHostGetDevicePointer(&t->fault, t->fault_host, 0));
    TOWER_TRY(cudaMallocHost(&t->wstage, t->wimg_elems * 2));
    TOWER_TRY(cudaMallocHost(&t->bstage, sizeof(float) * (nl + 1) * CH));
    TOWER_TRY(cudaFuncSetAttribute(k_value_tower<GeomC4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    TOWER_TRY(cudaFuncSetAttribute(k_value_tower<GeomChess, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    TOWER_TRY(cudaFuncSetAttribute(k_value_tower<GeomC4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
    TOWER_TRY(cudaFuncSetAttribute(k_value_tower<GeomChess, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL));
#undef TOWER_TRY
    if (int rc = zc_tower_update_weights(t, conv_w, conv_b, head_w, head_b, nullptr)) return cleanup(rc);
    if (cudaStreamSynchronize(nullptr) != cudaSuccess) return cleanup(fail(ZC_ECUDA, "zc_tower_create: weight upload failed"));
    g_towers.push_back(t);
    *out = t;
    return ZC_OK;
}

extern "C" int zc_tower_update_weights(zc_tower* t, const float* conv_w, const float* conv_b, const float* head_w, float head_b,
                                       void* stream) {
    if (!t) return fail(ZC_EINVAL, "tower handle is NULL");
    if (!conv_w || !conv_b || !head_w) return fail(ZC_EINVAL, "bad tower arguments");
    using namespace zc::tower;
    CUDA_TRY(cudaSetDevice(t->device));
    cudaStream_t st = (cudaStream_t)stream;
    // the staging buffers may still feed the previous update's copies, and no forward may read half-new weights
    CUDA_TRY(cudaDeviceSynchronize());
    pack_tower_weights(conv_w, t->n_layers, t->cin, t->f16, t->wstage);
    memcpy(t->bstage, conv_b, sizeof(float) * t->n_layers * CH);
    memcpy(t->bstage + (size_t)t->n_layers * CH, head_w, sizeof(float) * CH);
    t->head_b = head_b;
    CUDA_TRY(cudaMemcpyAsync(t->wimg2, t->wstage, t->wimg_elems * 2, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(t->bias, t->bstage, sizeof(float) * t->n_layers * CH, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(t->head_w, t->bstage + (size_t)t->n_layers * CH, sizeof(float) * CH, cudaMemcpyHostToDevice, st));
    return ZC_OK;
}

extern "C" unsigned int zc_tower_fault(zc_tower* t) {
    if (!t || !t->fault_host) return 0u;
    const unsigned f = *(volatile unsigned int*)t->fault_host;
    *t->fault_host = 0u;
    return f;
}

extern "C" void zc_tower_destroy(zc_tower* t) {
    if (!t) return;
    cudaSetDevice(t->device);
    cudaFree(t->wimg2);
    cudaFree(t->bias);
    cudaFree(t->head_w);
    cudaFreeHost(t->fault_host);
    cudaFreeHost(t->wstage);
    cudaFreeHost(t->bstage);
    g_towers.erase(std::remove(g_towers.begin(), g_towers.end(), t), g_towers.end());
    delete t;
}

extern "C" int zc_tower_forward(zc_tower* t, const void* dev_planes, int n_leaves, float* dev_values, void* stream) {
    if (!t) return fail(ZC_EINVAL, "tower handle is NULL");
    if (n_leaves < 0 || (n_leaves > 0 && (!dev_planes || !dev_values))) return fail(ZC_EINVAL, "bad forward arguments");
    if (n_leaves == 0) return ZC_OK;
    using namespace zc::tower;
    CUDA_TRY(cudaSetDevice(t->device));
    Params p;
    p.planes = reinterpret_cast<const uint16_t*>(dev_planes);
    p.wimg2 = t->wimg2;
    p.bias = t->bias;
    p.head_w = t->head_w;
    p.head_b = t->head_b;
    p.out = dev_values;
    p.n_leaves = n_leaves;
    p.n_layers = t->n_layers;
    p.fault = t->fault;
    const int nb = t->game == ZC_GAME_C4 ? GeomC4::NB : GeomChess::NB;
    const int n_groups = (n_leaves + nb - 1) / nb;
    constexpr int csz = 2;                             // CTA pairs (tcgen05 cta_group::2)
    int grid = std::max(1, std::min(t->n_sms, (n_groups + NT - 1) / NT));
    grid = std::max(csz, grid / csz * csz);            // whole pairs
    cudaStream_t st = (cudaStream_t)stream;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(N_THREADS);
    cfg.dynamicSmemBytes = SMEM_TOTAL;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)csz;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const bool c4 = t->game == ZC_GAME_C4;
    void (*kern)(const Params) = t->f16 ? (c4 ? k_value_tower<GeomC4, true> : k_value_tower<GeomChess, true>)
                                        : (c4 ? k_value_tower<GeomC4, false> : k_value_tower<GeomChess, false>);
    CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, p));
    ++t->launches;
    return ZC_OK;
}

extern "C" int64_t zc_tower_launches(const zc_tower* t) { return t ? t->launches : 0; }
extern "C" int zc_tower_plane_dtype(const zc_tower* t) { return t && !t->f16 ? ZC_PLANE_BF16 : ZC_PLANE_F16; }
