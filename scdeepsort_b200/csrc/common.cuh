// Shared host/device helpers for libwsage (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <atomic>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/wsage.h"

namespace wsage {

constexpr int kNumSMs = 148;   // B200: 2 dies x 74 SMs; grids are sized in multiples of this

inline thread_local char g_err[512] = "";
inline std::atomic<int64_t> g_launches{0};   // process-wide: autograd runs backward on its own thread

inline int fail(int code, const char* fmt, const char* a = "", const char* b = "") {
    snprintf(g_err, sizeof(g_err), fmt, a, b);
    return code;
}

inline int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(WSAGE_ECUDA, "%s: %s", what, cudaGetErrorString(e));
    ++g_launches;
    return WSAGE_OK;
}

#define WSAGE_REQUIRE(cond, msg) \
    do { if (!(cond)) return ::wsage::fail(WSAGE_EINVAL, "%s: %s", __func__, msg); } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// The alpha-index cascade of /root/reference/models/gnn.py:49-53, evaluated per edge on device.
__device__ __forceinline__ int alpha_index(int src_id, int dst_id, int gene_num) {
    int k = gene_num + 1;                                   // cell-cell (self loop)
    if (src_id >= 0 && dst_id < 0) k = src_id;              // gene -> cell
    if (dst_id >= 0 && src_id < 0) k = dst_id;              // cell -> gene
    if (dst_id >= 0 && src_id >= 0) k = gene_num;           // gene - gene
    return k;
}

template <int VEC> struct Vec;
template <> struct Vec<4> {
    using T = float4;
    static __device__ __forceinline__ T ldg(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
    static __device__ __forceinline__ T ld(const float* p) { return *reinterpret_cast<const float4*>(p); }
    static __device__ __forceinline__ void st(float* p, T v) { *reinterpret_cast<float4*>(p) = v; }
    static __device__ __forceinline__ T zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
    static __device__ __forceinline__ void fma(T& a, float s, T v) {
        a.x = fmaf(s, v.x, a.x); a.y = fmaf(s, v.y, a.y); a.z = fmaf(s, v.z, a.z); a.w = fmaf(s, v.w, a.w);
    }
    static __device__ __forceinline__ T scale(float s, T v) { return make_float4(s * v.x, s * v.y, s * v.z, s * v.w); }
    static __device__ __forceinline__ float dot(T a, T b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
    static __device__ __forceinline__ void atomic_add(float* p, T v) { atomicAdd(reinterpret_cast<float4*>(p), v); }
};
template <> struct Vec<1> {
    using T = float;
    static __device__ __forceinline__ T ldg(const float* p) { return __ldg(p); }
    static __device__ __forceinline__ T ld(const float* p) { return *p; }
    static __device__ __forceinline__ void st(float* p, T v) { *p = v; }
    static __device__ __forceinline__ T zero() { return 0.f; }
    static __device__ __forceinline__ void fma(T& a, float s, T v) { a = fmaf(s, v, a); }
    static __device__ __forceinline__ T scale(float s, T v) { return s * v; }
    static __device__ __forceinline__ float dot(T a, T b) { return a * b; }
    static __device__ __forceinline__ void atomic_add(float* p, T v) { atomicAdd(p, v); }
};

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace wsage
