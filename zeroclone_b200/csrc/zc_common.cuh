// Shared helpers for host+device code of libzc_b200.so
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define ZC_HD __host__ __device__ __forceinline__
#define ZC_D __device__ __forceinline__
// one out-of-line copy per kernel: bodies that are large and have several call sites (instruction-cache footprint)
#ifndef ZC_AB_CALL_ATTR
#define ZC_AB_CALL_ATTR __noinline__
#endif
#define ZC_HD_CALL __host__ __device__ ZC_AB_CALL_ATTR
#else
#define ZC_HD inline
#define ZC_D inline
#define ZC_HD_CALL inline
#endif

ZC_HD int zc_popc64(uint64_t v) {
#ifdef __CUDA_ARCH__
    return __popcll(v);
#else
    return __builtin_popcountll(v);
#endif
}
ZC_HD int zc_popc32(uint32_t v) {
#ifdef __CUDA_ARCH__
    return __popc(v);
#else
    return __builtin_popcount(v);
#endif
}
// index of the lowest set bit (v != 0)
ZC_HD int zc_ctz64(uint64_t v) {
#ifdef __CUDA_ARCH__
    return __ffsll((long long)v) - 1;
#else
    return __builtin_ctzll(v);
#endif
}
// number of leading zero bits (v != 0)
ZC_HD int zc_clz64(uint64_t v) {
#ifdef __CUDA_ARCH__
    return __clzll((long long)v);
#else
    return __builtin_clzll(v);
#endif
}
