// GPU neighbour sampler (sm_100a): uniform sampling WITHOUT replacement of at most `fanout` in-edges
// per destination node, the operation behind dgl.contrib.sampling.NeighborSampler(expand_factor=k,
// neighbor_type='in') (call sites /root/reference/train.py:71-78, predict.py:64-71; DGL 0.4.3's
// _CAPI_UniformSampling, C++/OpenMP on the host).  The self-loop is an ordinary in-edge and weights are
// not renormalised, as in the reference.
//
// One warp per destination node.  deg <= fanout: every in-edge is kept.  Otherwise Floyd's algorithm
// draws an exactly uniform fanout-subset of [0, deg) in fanout sequential steps (lane i keeps the i-th
// pick; membership test is one ballot), a bitonic sort across the lanes restores ascending edge order,
// and the picks are written as absolute positions into the parent CSR.  Randomness is a counter-based
// hash of (seed, node, step): reproducible, no generator state.
#pragma once
#include "common.cuh"

namespace wsage {

constexpr int kSampleMaxFanout = 32;

__device__ __forceinline__ uint32_t mix32(uint32_t h) {
    h = (h ^ (h >> 16)) * 0x45D9F3Bu;
    h = (h ^ (h >> 16)) * 0x45D9F3Bu;
    return h ^ (h >> 16);
}

__global__ void __launch_bounds__(256)
sample_neighbors_kernel(const int64_t* __restrict__ rowptr, const int64_t* __restrict__ nodes, int64_t n_nodes,
                        int fanout, uint64_t seed, int64_t* __restrict__ out_eid, int32_t* __restrict__ out_deg) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * 8;
    for (int64_t i = warp0; i < n_nodes; i += nwarps) {
        const int64_t v = nodes[i];
        const int64_t beg = rowptr[v];
        const int64_t deg = rowptr[v + 1] - beg;
        int64_t* dst = out_eid + i * fanout;
        if (deg <= fanout) {
            if (lane < deg) dst[lane] = beg + lane;
            if (lane == 0) out_deg[i] = (int32_t)deg;
            continue;
        }
        // Floyd: for j = deg-k .. deg-1: t = U[0, j]; pick t unless already picked, else j
        uint32_t mine = 0xFFFFFFFFu;                 // lane s holds the pick of step s (as offset in [0, deg))
        const uint32_t base = mix32((uint32_t)seed ^ mix32((uint32_t)(seed >> 32) + 0x9E3779B9u)) ^
                              mix32((uint32_t)v * 0x9E3779B1u + (uint32_t)((uint64_t)v >> 32));
        for (int s = 0; s < fanout; ++s) {
            const uint64_t j = (uint64_t)(deg - fanout + s);
            const uint32_t r1 = mix32(base + (uint32_t)s * 0x85EBCA77u);
            const uint32_t r2 = mix32(r1 ^ 0x68E31DA4u);
            const uint64_t r = ((uint64_t)r1 << 32) | r2;
            // deg < 2^31 here (edge offsets inside one row), so 64-bit multiply-high keeps the bias < 2^-33
            uint32_t t = (uint32_t)__umul64hi(r, j + 1);
            const bool dup = __ballot_sync(0xffffffffu, lane < s && mine == t) != 0;
            if (dup) t = (uint32_t)j;
            if (lane == s) mine = t;
        }
        // ascending order: bitonic sort of 32 lanes (unused lanes carry 0xFFFFFFFF and sink to the end)
#pragma unroll
        for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
            for (int j = k >> 1; j > 0; j >>= 1) {
                const uint32_t other = __shfl_xor_sync(0xffffffffu, mine, j);
                const bool up = ((lane & k) == 0);
                const bool lower = ((lane & j) == 0);
                const uint32_t lo = min(mine, other), hi = max(mine, other);
                mine = (up == lower) ? lo : hi;
            }
        }
        if (lane < fanout) dst[lane] = beg + (int64_t)mine;
        if (lane == 0) out_deg[i] = fanout;
    }
}

}  // namespace wsage
