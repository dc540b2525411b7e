// The exchange step of the cell-sharded full-graph pass, fused with the reduction that precedes it
// and the epilogue that follows it, over NVLink peer memory:
//
//   raw_g  = SUM over ranks of  SUM over split-K slabs of this rank's  X_shard^T · H_cells        (gene rows)
//   out_g  = dscale_g · raw_g + selfcoef_g · hself_g
//
// i.e. the `fn.mean` reduce of /root/reference/models/gnn.py:65 for the gene destinations, whose incoming messages are
// spread over the ranks' cell shards.  One persistent kernel per rank:
//
//   phase 0   slabs of wsage_dense16 (side 1), local           -> partial[rank]           (every gene row)
//   barrier   flags in every peer's buffer (st.release.sys / ld.acquire.sys), grid-wide through a local counter
//   phase A   this rank's slice of rows: SUM_p partial[p] (peer loads, rank order: every rank gets the same bits)
//                                                               -> result[rank], raw / out of the slice
//   barrier
//   phase B   the other slices: result[owner] (peer loads)     -> raw / out
//
// = reduce-scatter + all-gather by loads only; no trailing barrier: partial / result are double-buffered by call parity
// and a rank cannot be more than one call ahead of the slowest peer (it would be waiting in that call's first barrier).
// A barrier that does not complete within `timeout_ns` sets *status and the kernel returns (no hang, no trap); the
// buffers are then unusable and the host side raises.
#pragma once
#include "common.cuh"

namespace wsage {

constexpr int kPeerMax = 8;
constexpr int kPeerThreads = 512;     // at most 64 registers: two CTAs per SM, the bandwidth phases need the occupancy
// layout of the header at the start of every rank's peer allocation (bytes)
constexpr int kPeerFlagsOff = 0;        // unsigned flags[kPeerMax]: flags[r] written by rank r
constexpr int kPeerArriveOff = 256;     // unsigned: CTAs of the local grid that reached the barrier (monotonic)
constexpr int kPeerReleaseOff = 260;    // unsigned: last barrier the local grid may pass
constexpr int kPeerStatusOff = 264;     // int: 0 ok, 1 barrier timed out
constexpr int kPeerHeaderBytes = 4096;

struct PeerReduceParams {
    int rank, world;
    const float* slabs;          // [n_slabs][slab_rows][dim]
    int n_slabs;
    int64_t slab_rows;
    const int32_t* slot_of_row;  // slab row of gene row r (NULL: r)
    int64_t rows;
    int dim;
    float* partial[kPeerMax];    // [rows][dim] of this call's parity, one per rank (peer-mapped)
    float* result[kPeerMax];
    unsigned char* header[kPeerMax];
    unsigned epoch;              // number of the call's first barrier (the second is epoch + 1); starts at 1
    const float* dscale;
    const float* selfcoef;
    const float* hself;
    int64_t ld_hself;
    float* out;                  // [rows][ld_out] or NULL
    int64_t ld_out;
    float* raw;                  // [rows][ld_raw] or NULL
    int64_t ld_raw;
    unsigned long long timeout_ns;
};

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// data another GPU wrote before its release: never from a stale L1 line
__device__ __forceinline__ float4 ld_peer_f4(const float* p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// All CTAs of every rank's grid have finished what precedes; returns false after a timeout (uniform over the CTA).
__device__ bool peer_barrier(const PeerReduceParams& p, unsigned epoch, unsigned long long t_start) {
    __shared__ int s_fail;
    unsigned char* hdr = p.header[p.rank];
    unsigned* arrive = reinterpret_cast<unsigned*>(hdr + kPeerArriveOff);
    unsigned* release = reinterpret_cast<unsigned*>(hdr + kPeerReleaseOff);
    int* status = reinterpret_cast<int*>(hdr + kPeerStatusOff);
    if (threadIdx.x == 0) s_fail = 0;
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();          // this CTA's stores (ordered before by the barrier above) are visible system-wide
        atomicAdd(arrive, 1u);
    }
    if (blockIdx.x == 0) {
        if (threadIdx.x == 0) {
            const unsigned target = epoch * gridDim.x;
            while (ld_acquire_gpu(arrive) < target)
                if (global_ns() - t_start > p.timeout_ns) { s_fail = 1; break; }
        }
        __syncthreads();
        if (threadIdx.x < p.world && !s_fail) {
            const int peer = threadIdx.x;
            st_release_sys(reinterpret_cast<unsigned*>(p.header[peer] + kPeerFlagsOff) + p.rank, epoch);
            const unsigned* mine = reinterpret_cast<const unsigned*>(hdr + kPeerFlagsOff) + peer;
            while (ld_acquire_sys(mine) < epoch)
                if (global_ns() - t_start > p.timeout_ns) { s_fail = 1; break; }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            if (s_fail) *status = 1;
            __threadfence();
            st_release_gpu(release, s_fail ? 0xffffffffu : epoch);      // all-ones: every later wait falls through and reports
        }
    }
    if (threadIdx.x == 0) {
        unsigned v;
        while ((v = ld_acquire_gpu(release)) < epoch)
            if (global_ns() - t_start > 2 * p.timeout_ns) { v = 0xffffffffu; *status = 1; break; }
        if (v == 0xffffffffu) s_fail = 1;
    }
    __syncthreads();
    return !s_fail;
}

__device__ __forceinline__ void peer_emit(const PeerReduceParams& p, int64_t r, int c4, float4 v) {
    if (p.raw) *reinterpret_cast<float4*>(p.raw + r * p.ld_raw + c4 * 4) = v;
    if (p.out) {
        const float d = p.dscale ? __ldg(p.dscale + r) : 1.f;
        float4 o = make_float4(v.x * d, v.y * d, v.z * d, v.w * d);
        if (p.selfcoef) {
            const float s = __ldg(p.selfcoef + r);
            const float4 h = __ldg(reinterpret_cast<const float4*>(p.hself + r * p.ld_hself + c4 * 4));
            o.x = fmaf(s, h.x, o.x); o.y = fmaf(s, h.y, o.y); o.z = fmaf(s, h.z, o.z); o.w = fmaf(s, h.w, o.w);
        }
        *reinterpret_cast<float4*>(p.out + r * p.ld_out + c4 * 4) = o;
    }
}

__global__ void __launch_bounds__(kPeerThreads, 2) peer_reduce_kernel(const __grid_constant__ PeerReduceParams p) {
    const unsigned long long t_start = global_ns();
    const int q4 = p.dim / 4;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;

    // ---- phase 0: this rank's split-K slabs -> partial[rank] ----
    {
        float4* dst = reinterpret_cast<float4*>(p.partial[p.rank]);
        const int64_t slab_stride = p.slab_rows * p.dim;
        const int64_t total = p.rows * q4;
        for (int64_t i0 = tid; i0 < total; i0 += 2 * nthreads) {          // two independent chains per thread
            const int64_t i1 = i0 + nthreads;
            const bool two = i1 < total;
            const int64_t r0 = i0 / q4, r1 = two ? i1 / q4 : r0;
            const int64_t s0 = p.slot_of_row ? __ldg(p.slot_of_row + r0) : r0;
            const int64_t s1 = p.slot_of_row ? __ldg(p.slot_of_row + r1) : r1;
            const float* a0 = p.slabs + s0 * p.dim + (i0 - r0 * q4) * 4;
            const float* a1 = p.slabs + s1 * p.dim + ((two ? i1 : i0) - r1 * q4) * 4;
            float4 x = __ldg(reinterpret_cast<const float4*>(a0));
            float4 y = __ldg(reinterpret_cast<const float4*>(a1));
            for (int s = 1; s < p.n_slabs; ++s) {
                const float4 u = __ldg(reinterpret_cast<const float4*>(a0 + s * slab_stride));
                const float4 v = __ldg(reinterpret_cast<const float4*>(a1 + s * slab_stride));
                x.x += u.x; x.y += u.y; x.z += u.z; x.w += u.w;
                y.x += v.x; y.y += v.y; y.z += v.z; y.w += v.w;
            }
            dst[i0] = x;
            if (two) dst[i1] = y;
        }
    }
    if (!peer_barrier(p, p.epoch, t_start)) return;

    // ---- phase A: my slice, summed over the ranks in rank order ----
    const int64_t r_lo = p.rows * p.rank / p.world, r_hi = p.rows * (p.rank + 1) / p.world;
    {
        float4* res = reinterpret_cast<float4*>(p.result[p.rank]);
        for (int64_t i = r_lo * q4 + tid; i < r_hi * q4; i += nthreads) {
            float4 acc = ld_peer_f4(p.partial[0] + i * 4);
            for (int k0 = 1; k0 < p.world; k0 += 4) {                      // up to four peer loads in flight, added in rank order
                float4 v[4];
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (k0 + k < p.world) v[k] = ld_peer_f4(p.partial[k0 + k] + i * 4);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (k0 + k < p.world) { acc.x += v[k].x; acc.y += v[k].y; acc.z += v[k].z; acc.w += v[k].w; }
            }
            res[i] = acc;
            const int64_t r = i / q4;
            peer_emit(p, r, (int)(i - r * q4), acc);
        }
    }
    if (!peer_barrier(p, p.epoch + 1, t_start)) return;

    // ---- phase B: the other ranks' slices from their owners ----
    for (int step = 1; step < p.world; ++step) {
        const int owner = (p.rank + step) % p.world;        // every rank starts on a different peer
        const int64_t o_lo = p.rows * owner / p.world, o_hi = p.rows * (owner + 1) / p.world;
        const float* src = p.result[owner];
        for (int64_t i = o_lo * q4 + tid; i < o_hi * q4; i += nthreads) {
            const float4 v = ld_peer_f4(src + i * 4);
            const int64_t r = i / q4;
            peer_emit(p, r, (int)(i - r * q4), v);
        }
    }
}

}  // namespace wsage
