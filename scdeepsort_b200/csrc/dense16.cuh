// The popular genes' share of the bipartite aggregation on the 5th-generation tensor cores (sm_100a).
//
// For a gene expressed in more than a few percent of the cells, walking its edges one by one (agg_tiled.cuh:
// 14 shared-memory wavefronts per edge) costs more than multiplying through its zeros on tcgen05.  The graph
// builder therefore splits  X = X_sparse + X_dense : the dense part is stored zero-filled as 16-bit tiles and
//
//   side 0 (destinations = cells):  acc[c,:] = SUM_s  Xd[c,s] * H[gene(s),:]     A = Xd (K-major),  B = H^T (K-major)
//   side 1 (destinations = genes):  acc[s,:] = SUM_c  Xd[c,s] * H[c,:]           A = Xd (MN-major), B = H   (MN-major)
//
// run as GEMMs from ONE copy of Xd.  This replaces, for those entries, exactly what agg_tiled_kernel / the
// reference's message_func + fn.mean do (/root/reference/models/gnn.py:47-56,65).
//
// fp32 parity.  Operands are split into fp16 hi + fp16 lo after a power-of-two scaling into fp16's range
// (x: static, at build time; H: per pass, from its amax), and hi*hi + lo*hi + hi*lo is accumulated in fp32 TMEM:
// products of two 11-bit significands are exact in fp32, the dropped lo*lo term is 2^-22 relative — the same
// grade as the tf32x3 scheme of dense_tc.cuh at twice the MMA rate and half the bytes.  The tensor core adds into
// its accumulator with truncation, so a chain is cut every `chunk_kb` k-blocks (2048 rows): the accumulator is
// drained, scaled and ADDED to the output tile in global memory by a TMA reduce (cp.reduce.async.bulk.tensor .add,
// fp32 round-to-nearest in L2; one pending add per address, so the result is order-independent and bitwise
// reproducible).  fmt = bf16: one product, no scaling (BASELINE configs[2]).
//
// Storage of Xd (hi and lo planes alike): [cell tile of 128][gene block of 32][128 cells][32 slots] 16-bit, i.e.
// rows of 64 bytes.  Seen as a 2-D tensor of 64-byte rows, a TMA box {32, 128} is one contiguous 8 KB K-major
// A tile for side 0 and a box {32, 32} is a contiguous 2 KB piece of the MN-major A tile for side 1: with the
// 64-byte swizzle both are canonical UMMA layouts (K-major SW64: 8-row groups 512 B apart; MN-major SW64: 32
// contiguous MN elements, 8-row K atoms 512 B apart, MN blocks one box apart), so HBM is streamed in whole
// 8-32 KB pieces in both directions.
//
// CTA = 6 warps: warp 0 TMA producer, warp 1 TMEM allocator + single-thread tcgen05.mma issuer, warps 2-5
// drain (tcgen05.ld -> scale -> swizzled shared-memory staging -> TMA store / reduce-add).  Persistent over
// (destination tile, k-range) work items, static round-robin.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "dense_tc.cuh"

namespace wsage {

constexpr int kD16TileM = 128;                 // destinations per tile = MMA M = cells per storage tile
constexpr int kD16BlockK = 32;                 // 16-bit elements per 64-byte row = k-block
constexpr int kD16UmmaK = 16;                  // K of one tcgen05.mma.kind::f16
constexpr int kD16ABytes = kD16TileM * 64;     // 8 KB: one A tile (hi or lo)
constexpr int kD16BoxMN = 32 * 64;             // 2 KB: one {32, 32} box of the MN-major operands
constexpr int kD16DrainWarps = 8;                     // two per TMEM lane quarter: each takes every other 16-column block
constexpr int kD16Threads = 32 * (2 + kD16DrainWarps);
constexpr int kD16OutCols = 16;                // fp32 columns per TMA store box (64 bytes)
constexpr int kD16BufBytes = 32 * 64;                  // one staging buffer of a drain warp: [32 rows][64 B]
constexpr int kD16MaxBufs = 6;                         // staging buffers per drain warp (TMA stores in flight)
constexpr int kD16DefaultChunkRows = 2048;
constexpr int kD16MaxStages = 6;
constexpr int kD16SmemBudget = 227 * 1024;

// ------------------------------------------------------------------------------------------------
// dynamic power-of-two scale of the per-pass operand: amax * scale in [2^13, 2^14)
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ int d16_scale_exp(float amax) {
    if (!(amax > 0.f) || amax > 3.0e38f) return 0;
    int e;
    frexpf(amax, &e);                           // amax = m * 2^e, m in [0.5, 1)
    int k = 14 - e;
    return k > 100 ? 100 : (k < -100 ? -100 : k);
}

__global__ void __launch_bounds__(256)
amax_kernel(const float* __restrict__ x, int64_t ld, const int32_t* __restrict__ row_ids, const float* __restrict__ rowscale,
            int64_t rows, int cols, unsigned* __restrict__ out) {
    const int64_t n4 = cols / 4;
    const int64_t total = rows * n4;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (int64_t)gridDim.x * blockDim.x;
    float m = 0.f;
    if (!row_ids && !rowscale && ld == cols) {
        // a contiguous matrix: a flat sweep, four independent 16-byte loads per thread and iteration
        const float4* p = reinterpret_cast<const float4*>(x);
        int64_t i = tid;
        for (; i + 3 * nthreads < total; i += 4 * nthreads) {
            const float4 a = __ldg(p + i), b = __ldg(p + i + nthreads), c = __ldg(p + i + 2 * nthreads), d = __ldg(p + i + 3 * nthreads);
            const float ma = fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w)));
            const float mb = fmaxf(fmaxf(fabsf(b.x), fabsf(b.y)), fmaxf(fabsf(b.z), fabsf(b.w)));
            const float mc = fmaxf(fmaxf(fabsf(c.x), fabsf(c.y)), fmaxf(fabsf(c.z), fabsf(c.w)));
            const float md = fmaxf(fmaxf(fabsf(d.x), fabsf(d.y)), fmaxf(fabsf(d.z), fabsf(d.w)));
            m = fmaxf(m, fmaxf(fmaxf(ma, mb), fmaxf(mc, md)));
        }
        for (; i < total; i += nthreads) {
            const float4 a = __ldg(p + i);
            m = fmaxf(m, fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))));
        }
    } else {
        for (int64_t i = tid; i < total; i += nthreads) {
            const int64_t r = i / n4;
            const int c = (int)(i - r * n4) * 4;
            const int64_t src = row_ids ? (int64_t)__ldg(row_ids + r) : r;
            const float4 v = __ldg(reinterpret_cast<const float4*>(x + src * ld + c));
            const float s = rowscale ? fabsf(__ldg(rowscale + src)) : 1.f;
            m = fmaxf(m, s * fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    __shared__ float wm[8];
    if ((threadIdx.x & 31) == 0) wm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) m = fmaxf(m, wm[w]);
        atomicMax(out, __float_as_uint(m));     // non-negative floats order like their bit patterns
    }
}

__device__ __forceinline__ void d16_split(float v, int fmt, unsigned short& hi, unsigned short& lo) {
    if (fmt == 0) {
        const __half h = __float2half_rn(v);
        hi = __half_as_ushort(h);
        lo = __half_as_ushort(__float2half_rn(v - __half2float(h)));
    } else {
        hi = __bfloat16_as_ushort(__float2bfloat16_rn(v));
        lo = 0;
    }
}

// rows stay rows: hi/lo = split(v[r, c] * rowscale[r] * 2^k) with v = x, or x * (mask_src > 0) (ReLU backward), in one of
//   layout 0 (ROWS)       planes [rows][ld_o]
//   layout 2 (COLBLOCKS)  planes [ceil(cols / 32)][ld_o rows][32]: the K-major B operand of dense16_kernel when the k index
//                         is the COLUMN of x (a weight matrix [n_out][n_in]: no transposition); the columns past `cols` of
//                         the last block are written as zeros
//   layout 4 (BLOCKED)    planes [ceil(rows / 128)][ld_o / 32][128][32]: the A-operand layout of dense16_kernel (include/wsage.h)
//                         for an activation / gradient matrix, ld_o = slot padding; rows and columns past the matrix are zeros
__global__ void __launch_bounds__(256)
split16_kernel(const float* __restrict__ x, int64_t ld, const float* __restrict__ mask_src, int64_t ld_m, const float* __restrict__ rowscale,
               int64_t rows, int cols, const float* __restrict__ amax, int fmt,
               unsigned short* __restrict__ hi, unsigned short* __restrict__ lo, int64_t ld_o, int layout) {
    const float scale = (fmt == 0 && amax) ? ldexpf(1.f, d16_scale_exp(*amax)) : 1.f;
    const int64_t rows_p = layout == 4 ? (rows + 127) / 128 * 128 : rows;
    const int cols_p = layout == 4 ? (int)ld_o : (layout == 2 ? (cols + 31) / 32 * 32 : cols);
    const int nb = (int)(ld_o / 32);
    const int64_t n4 = cols_p / 4;
    const int64_t total = rows_p * n4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / n4;
        const int c = (int)(i - r * n4) * 4;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < rows && c < cols) {
            v = __ldg(reinterpret_cast<const float4*>(x + r * ld + c));
            if (mask_src) {
                const float4 y = __ldg(reinterpret_cast<const float4*>(mask_src + r * ld_m + c));
                v.x = y.x > 0.f ? v.x : 0.f; v.y = y.y > 0.f ? v.y : 0.f;
                v.z = y.z > 0.f ? v.z : 0.f; v.w = y.w > 0.f ? v.w : 0.f;
            }
            const float s = scale * (rowscale ? __ldg(rowscale + r) : 1.f);
            v.x *= s; v.y *= s; v.z *= s; v.w *= s;
        }
        unsigned short h[4], l[4];
        d16_split(v.x, fmt, h[0], l[0]); d16_split(v.y, fmt, h[1], l[1]);
        d16_split(v.z, fmt, h[2], l[2]); d16_split(v.w, fmt, h[3], l[3]);
        int64_t o;
        if (layout == 4) o = ((((r >> 7) * nb + (c >> 5)) << 7) + (r & 127)) * 32 + (c & 31);
        else if (layout == 2) o = ((int64_t)(c >> 5) * ld_o + r) * 32 + (c & 31);
        else o = r * ld_o + c;
        *reinterpret_cast<uint2*>(hi + o) = make_uint2(h[0] | ((unsigned)h[1] << 16), h[2] | ((unsigned)h[3] << 16));
        if (fmt == 0) *reinterpret_cast<uint2*>(lo + o) = make_uint2(l[0] | ((unsigned)l[1] << 16), l[2] | ((unsigned)l[3] << 16));
    }
}

// amax over v = x * (mask_src > 0) is bounded by amax over x: the ReLU-backward split uses the unmasked amax (a valid scale).

// partial[b, :] = SUM_{r in block b's rows} x[r, :] * (mask_src[r, :] > 0): the bias gradient of a Linear + ReLU layer (column sums
// of the masked output gradient); the per-block partials are added in block order by sum_slabs_kernel (deterministic).
// The k-blocked transposition for matrices with many rows: a block owns one k-block (32 source rows), each of its warps a
// tile of 32 columns (lane <-> column).  A thread reads its column of all 32 rows (32 loads in flight, each a warp-wide
// 128-byte row segment) and then owns one whole 64-byte output row per plane: the warp writes 2 KB contiguous per plane
// (split16_transpose_kernel's 16-byte pieces at a 64-byte stride: 0.65 ms for 780 k x 400; ideal 0.40).
__global__ void __launch_bounds__(128)
split16_kblocks_wide_kernel(const float* __restrict__ x, int64_t ld, const int32_t* __restrict__ row_ids, const float* __restrict__ rowscale,
                            int64_t rows, int cols, const float* __restrict__ amax, int fmt,
                            unsigned short* __restrict__ hi, unsigned short* __restrict__ lo, int64_t ld_o) {
    __shared__ int64_t s_src[32];
    __shared__ float s_rs[32];
    const float scale = (fmt == 0 && amax) ? ldexpf(1.f, d16_scale_exp(*amax)) : 1.f;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x < 32) {
        const int64_t sidx = (int64_t)blockIdx.x * 32 + threadIdx.x;
        const int64_t src = sidx < rows ? (row_ids ? (int64_t)__ldg(row_ids + sidx) : sidx) : -1;
        s_src[threadIdx.x] = src;
        s_rs[threadIdx.x] = (src >= 0 && rowscale) ? scale * __ldg(rowscale + src) : scale;
    }
    __syncthreads();
    for (int c0 = w * 32; c0 < cols; c0 += 128) {
        const int c = c0 + lane;
        if (c >= cols) continue;
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const int64_t src = s_src[i];
            v[i] = src >= 0 ? __ldg(x + src * ld + c) * s_rs[i] : 0.f;
        }
        const int64_t o = ((int64_t)blockIdx.x * ld_o + c) * 32;
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            unsigned short h[8], l[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) d16_split(v[g * 8 + i], fmt, h[i], l[i]);
            *reinterpret_cast<uint4*>(hi + o + g * 8) = make_uint4(h[0] | ((unsigned)h[1] << 16), h[2] | ((unsigned)h[3] << 16),
                                                                   h[4] | ((unsigned)h[5] << 16), h[6] | ((unsigned)h[7] << 16));
            if (fmt == 0)
                *reinterpret_cast<uint4*>(lo + o + g * 8) = make_uint4(l[0] | ((unsigned)l[1] << 16), l[2] | ((unsigned)l[3] << 16),
                                                                       l[4] | ((unsigned)l[5] << 16), l[6] | ((unsigned)l[7] << 16));
        }
    }
}

// Layout 4 (BLOCKED) on its own: a work item is one (tile of 128 rows, block of 32 columns) = 8 KB contiguous per plane.
// Thread <-> (row, 8 columns): two float4 loads (+ two of the mask), one 16-byte store per plane; a warp writes 512 contiguous
// bytes.  gridDim.x = (column blocks in use) x groups: a CTA keeps its column block and walks the tiles group, group + groups, ...
// Blocks entirely past `cols` are not written (side 0 never loads them, on side 1 they only feed output rows past the matrix).
// COLSUM: also partial[group][c] = SUM over the group's rows of v[r, c] (the bias gradient of the layer whose output gradient
// is being split: one pass over dy and the mask instead of two); fixed assignment and order, so the sums are reproducible.
template <bool COLSUM>
__global__ void __launch_bounds__(256)
split16_blocked_kernel(const float* __restrict__ x, int64_t ld, const float* __restrict__ mask_src, int64_t ld_m,
                       const float* __restrict__ rowscale, int64_t rows, int cols, const float* __restrict__ amax, int fmt,
                       unsigned short* __restrict__ hi, unsigned short* __restrict__ lo, int64_t ld_o, float* __restrict__ partial) {
    const float scale = (fmt == 0 && amax) ? ldexpf(1.f, d16_scale_exp(*amax)) : 1.f;
    const int nb = (int)(ld_o / 32), nb_used = (cols + 31) / 32;
    const int64_t n_tiles = (rows + 127) / 128;
    const int cb = blockIdx.x % nb_used, group = blockIdx.x / nb_used, groups = gridDim.x / nb_used;
    const int c8 = (threadIdx.x & 3) * 8, rr = threadIdx.x >> 2;            // 4 threads per row of the block, 64 rows per sweep
    const int c = cb * 32 + c8;
    float sum[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) sum[i] = 0.f;
    for (int64_t t = group; t < n_tiles; t += groups) {
        float4 v[2][2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int64_t r = t * 128 + j * 64 + rr;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                v[j][k] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (r < rows && c + 4 * k < cols) v[j][k] = __ldg(reinterpret_cast<const float4*>(x + r * ld + c + 4 * k));
            }
        }
        if (mask_src) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int64_t r = t * 128 + j * 64 + rr;
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    if (r < rows && c + 4 * k < cols) {
                        const float4 y = __ldg(reinterpret_cast<const float4*>(mask_src + r * ld_m + c + 4 * k));
                        v[j][k].x = y.x > 0.f ? v[j][k].x : 0.f; v[j][k].y = y.y > 0.f ? v[j][k].y : 0.f;
                        v[j][k].z = y.z > 0.f ? v[j][k].z : 0.f; v[j][k].w = y.w > 0.f ? v[j][k].w : 0.f;
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int64_t r = t * 128 + j * 64 + rr;
            if (COLSUM) {
                sum[0] += v[j][0].x; sum[1] += v[j][0].y; sum[2] += v[j][0].z; sum[3] += v[j][0].w;
                sum[4] += v[j][1].x; sum[5] += v[j][1].y; sum[6] += v[j][1].z; sum[7] += v[j][1].w;
            }
            const float sc = scale * ((rowscale && r < rows) ? __ldg(rowscale + r) : 1.f);
            const float f[8] = {v[j][0].x * sc, v[j][0].y * sc, v[j][0].z * sc, v[j][0].w * sc,
                                v[j][1].x * sc, v[j][1].y * sc, v[j][1].z * sc, v[j][1].w * sc};
            unsigned short h[8], l[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) d16_split(f[i], fmt, h[i], l[i]);
            const int64_t o = (((t * nb + cb) << 7) + (j * 64 + rr)) * 32 + c8;
            *reinterpret_cast<uint4*>(hi + o) = make_uint4(h[0] | ((unsigned)h[1] << 16), h[2] | ((unsigned)h[3] << 16),
                                                           h[4] | ((unsigned)h[5] << 16), h[6] | ((unsigned)h[7] << 16));
            if (fmt == 0)
                *reinterpret_cast<uint4*>(lo + o) = make_uint4(l[0] | ((unsigned)l[1] << 16), l[2] | ((unsigned)l[3] << 16),
                                                               l[4] | ((unsigned)l[5] << 16), l[6] | ((unsigned)l[7] << 16));
        }
    }
    if (COLSUM) {
        // lanes with equal (lane & 3) hold the same 8 columns: butterfly over the other lane bits, then the 8 warps in order
        __shared__ float red[8][32];
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            sum[i] += __shfl_xor_sync(0xffffffffu, sum[i], 4);
            sum[i] += __shfl_xor_sync(0xffffffffu, sum[i], 8);
            sum[i] += __shfl_xor_sync(0xffffffffu, sum[i], 16);
        }
        if (lane < 4) {
#pragma unroll
            for (int i = 0; i < 8; ++i) red[warp][lane * 8 + i] = sum[i];
        }
        __syncthreads();
        if (threadIdx.x < 32 && cb * 32 + threadIdx.x < cols) {
            float a = red[0][threadIdx.x];
#pragma unroll
            for (int w = 1; w < 8; ++w) a += red[w][threadIdx.x];
            partial[(int64_t)group * cols + cb * 32 + threadIdx.x] = a;
        }
    }
}

__global__ void __launch_bounds__(256)
colsum_masked_kernel(const float* __restrict__ x, int64_t ld, const float* __restrict__ mask_src, int64_t ld_m,
                     int64_t rows, int cols, float* __restrict__ partial) {
    __shared__ float4 red[256];
    const int ncg = cols / 4;                              // column groups of 4 (host guarantees ncg <= 256)
    const int lanes = 256 / ncg;                           // rows in flight per block
    const int cg = threadIdx.x % ncg, rl = threadIdx.x / ncg;
    const int64_t per = (rows + gridDim.x - 1) / gridDim.x;
    const int64_t r0 = (int64_t)blockIdx.x * per, r1 = min(rows, r0 + per);
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    if (rl < lanes) {
        for (int64_t r = r0 + rl; r < r1; r += lanes) {
            float4 v = __ldg(reinterpret_cast<const float4*>(x + r * ld + cg * 4));
            if (mask_src) {
                const float4 y = __ldg(reinterpret_cast<const float4*>(mask_src + r * ld_m + cg * 4));
                v.x = y.x > 0.f ? v.x : 0.f; v.y = y.y > 0.f ? v.y : 0.f;
                v.z = y.z > 0.f ? v.z : 0.f; v.w = y.w > 0.f ? v.w : 0.f;
            }
            a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        }
    }
    red[threadIdx.x] = a;
    __syncthreads();
    if (rl == 0) {
        for (int l = 1; l < lanes; ++l) {
            const float4 t = red[l * ncg + cg];
            a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
        }
        *reinterpret_cast<float4*>(partial + (size_t)blockIdx.x * cols + cg * 4) = a;
    }
}

// out[r, :] = SUM_k slabs[k][r, :]  (fixed order): the split-K slabs of a weight gradient
__global__ void __launch_bounds__(256)
sum_slabs_kernel(const float* __restrict__ slabs, int n_slabs, int64_t slab_stride, int64_t rows, int cols,
                 float* __restrict__ out, int64_t ld_out) {
    const int64_t n4 = cols / 4;
    const int64_t total = rows * n4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / n4;
        const int c = (int)(i - r * n4) * 4;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int k = 0; k < n_slabs; ++k) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(slabs + (size_t)k * slab_stride + r * cols + c));
            a.x += t.x; a.y += t.y; a.z += t.z; a.w += t.w;
        }
        *reinterpret_cast<float4*>(out + r * ld_out + c) = a;
    }
}

// gathered and transposed: hi/lo[c, s] = split(x[row_ids[s], c] * rowscale[row_ids[s]] * 2^k).
// kblocks == 0: planes [cols][ld_o] (row c = the rows of x as columns).
// kblocks != 0: planes [ceil(rows / 32)][ld_o][32] — the K-major B operand of dense16_kernel cut into k-blocks of 32 source
//   rows, so that one k-block of all `cols` output columns is ONE contiguous piece of memory (a [cols][rows] matrix would be
//   read as 64-byte granules a whole row pitch apart: 1.5 MB at atlas scale).  Entries of the last block past `rows` are
//   written as zeros (they meet real entries of X); rows cols .. ld_o of a block are left unwritten (never stored columns).
__global__ void __launch_bounds__(128)
split16_transpose_kernel(const float* __restrict__ x, int64_t ld, const int32_t* __restrict__ row_ids, const float* __restrict__ rowscale,
                         int64_t rows, int cols, const float* __restrict__ amax, int fmt,
                         unsigned short* __restrict__ hi, unsigned short* __restrict__ lo, int64_t ld_o, int kblocks) {
    // A block owns 32 source rows (one k-block) and walks the column tiles; lane <-> column, warp q <-> source rows 8q .. 8q+7.
    // Every load is a warp-wide 128-byte row segment; every thread then owns 8 consecutive k entries of one output row and
    // writes them as ONE 16-byte store per plane — no shared memory, no 2-byte stores (the first version: 1.0 ms for 780 k x 400,
    // 2.5 TB/s).
    const float scale = (fmt == 0 && amax) ? ldexpf(1.f, d16_scale_exp(*amax)) : 1.f;
    const int lane = threadIdx.x & 31, q = threadIdx.x >> 5;          // 32 x 4
    const int64_t s0 = (int64_t)blockIdx.x * 32 + q * 8;
    int64_t src[8];
    float rs[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t sidx = s0 + i;
        src[i] = sidx < rows ? (row_ids ? (int64_t)__ldg(row_ids + sidx) : sidx) : -1;
        rs[i] = (src[i] >= 0 && rowscale) ? scale * __ldg(rowscale + src[i]) : scale;
    }
    for (int c0 = blockIdx.y * 32; c0 < cols; c0 += gridDim.y * 32) {
        const int c = c0 + lane;
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = (src[i] >= 0 && c < cols) ? __ldg(x + src[i] * ld + c) * rs[i] : 0.f;
        if (c >= cols) continue;
        unsigned short h[8], l[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) d16_split(v[i], fmt, h[i], l[i]);
        const uint4 ph = make_uint4(h[0] | ((unsigned)h[1] << 16), h[2] | ((unsigned)h[3] << 16), h[4] | ((unsigned)h[5] << 16), h[6] | ((unsigned)h[7] << 16));
        const uint4 pl = make_uint4(l[0] | ((unsigned)l[1] << 16), l[2] | ((unsigned)l[3] << 16), l[4] | ((unsigned)l[5] << 16), l[6] | ((unsigned)l[7] << 16));
        if (kblocks) {                                       // [k-block][ld_o][32]: zeros past `rows` are part of the layout
            const int64_t o = ((int64_t)blockIdx.x * ld_o + c) * 32 + q * 8;
            *reinterpret_cast<uint4*>(hi + o) = ph;
            if (fmt == 0) *reinterpret_cast<uint4*>(lo + o) = pl;
        } else {                                             // [cols][ld_o]
            const int64_t o = (int64_t)c * ld_o + s0;
            if (s0 + 8 <= rows) {
                *reinterpret_cast<uint4*>(hi + o) = ph;
                if (fmt == 0) *reinterpret_cast<uint4*>(lo + o) = pl;
            } else {
                for (int i = 0; i < 8 && s0 + i < rows; ++i) {
                    hi[o + i] = h[i];
                    if (fmt == 0) lo[o + i] = l[i];
                }
            }
        }
    }
}

// out[r] = <a[r, :], b[r, :]>: warp per row (the self-loop terms of d-alpha: sum_c s_c <h_c, dn_c>)
__global__ void __launch_bounds__(256)
rowdot_kernel(const float* __restrict__ a, int64_t ld_a, const float* __restrict__ b, int64_t ld_b, int64_t rows, int cols,
              float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    for (int64_t r = warp0; r < rows; r += (int64_t)gridDim.x * 8) {
        float acc = 0.f;
        for (int c = lane * 4; c < cols; c += 128) {
            const float4 u = __ldg(reinterpret_cast<const float4*>(a + r * ld_a + c));
            const float4 v = __ldg(reinterpret_cast<const float4*>(b + r * ld_b + c));
            acc += u.x * v.x + u.y * v.y + u.z * v.z + u.w * v.w;
        }
        acc = warp_sum(acc);
        if (lane == 0) out[r] = acc;
    }
}

// ------------------------------------------------------------------------------------------------
// PTX wrappers (TMA store side; the load / MMA / TMEM wrappers are dense_tc.cuh's)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int x, int y, int z, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)) : "memory");
}
// L2 eviction policies: the X planes are read exactly once per pass (evict first), the per-pass operand is re-read by
// every CTA and the output tiles are re-read by the next chain's reduce-add (evict last)
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ void tma_load_2d_hint(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar, uint64_t pol) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ void tma_load_3d_hint(void* dst, const CUtensorMap* map, int x, int y, int z, uint64_t* bar, uint64_t pol) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)), "l"(pol) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int x, int y, uint64_t pol) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;"
                 ::"l"(map), "r"(smem_u32(src)), "r"(x), "r"(y), "l"(pol) : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, const void* src, int x, int y, uint64_t pol) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;"
                 ::"l"(map), "r"(smem_u32(src)), "r"(x), "r"(y), "l"(pol) : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// waits until at most n of this thread's bulk groups still read their shared-memory source
__device__ __forceinline__ void bulk_wait_read(int n) {
    switch (n) {
        case 1: asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); break;
        case 2: asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory"); break;
        case 3: asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory"); break;
        case 4: asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory"); break;
        case 5: asm volatile("cp.async.bulk.wait_group.read 5;" ::: "memory"); break;
        default: asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); break;
    }
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// MN-major operand, 64-byte swizzle: 32 contiguous MN elements per row, 8-row K atoms 512 B apart (SBO), the next
// 32-element MN block one box further (LBO).
__device__ __forceinline__ uint64_t umma_desc_mn_sw64(const void* smem) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_u32(smem) & 0x3FFFF) >> 4);
    d |= (uint64_t)(kD16BoxMN >> 4) << 16;
    d |= (uint64_t)(512 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)4 << 61;                                    // SWIZZLE_64B
    return d;
}
// kind::f16 instruction descriptor: D fp32, A/B fp16 (0) or bf16 (1), K-major or MN-major (both operands alike).
__device__ __forceinline__ uint32_t umma_idesc_f16(int n, int bf16, int mn_major) {
    return (1u << 4) | ((uint32_t)bf16 << 7) | ((uint32_t)bf16 << 10) | ((uint32_t)mn_major << 15) | ((uint32_t)mn_major << 16) |
           ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kD16TileM >> 4) << 24);
}
// issue only: the registers are valid after tmem_ld_wait()
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
// the registers are tied to the wait ("+r"), so no use of them can be scheduled above it
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :: "memory");
}

struct D16Params {
    int side;                    // 0: cell destinations (A K-major), 1: gene-slot destinations (A MN-major); B is K-major on both sides
    int terms;                   // 3: hi*hi + lo*hi + hi*lo;  1: single product
    int bf16;
    int n, n_pad, n1, n2;        // output width, rounded up to 16, MMA N halves (n1 <= 256)
    int stages, stage_bytes, tx_bytes, b_bytes;      // stage_bytes: ring pitch (1 KB multiple); tx_bytes: bytes TMA delivers per stage and CTA
    int nbuf;                    // staging buffers per drain warp
    int drain_diag;              // -DWSAGE_TUNING only
    int m_tiles;                 // destination tiles of 128 (PAIR: an even number of them is processed, the last one may be empty)
    int nb;                      // 32-slot gene blocks per cell tile of the storage
    int ld_hb;                   // rows per k-block of the B planes
    int num_kb, chunk_kb;        // k-blocks in all / per accumulation chain
    int n_splits, kb_per_split;  // side 0: 1, num_kb
    // side 0, last (partial) round of destination tiles: its tail_tiles tiles are cut along k into tail_splits items of tail_kb
    // k-blocks each, which reduce-add into rows the host zeroed — the round then costs 1 / tail_splits of a tile.  Items below
    // full_items are decoded the plain way (all items when there is no tail).
    int full_items, tail_tiles, tail_splits, tail_kb;
    int64_t rows_per_split;      // side 1: rows of the output map per split (padded slots)
    int64_t m_total;             // destination rows
    const float* amax;           // device scalar the per-pass operand was scaled by, or null
    const float* x_amax;         // device scalar the X planes were scaled by (activations as A operand), or null: x_scale_inv
    float x_scale_inv;
    const float* bias;           // side 0, per output column, added on a tile's first chain; or null
    int relu;                    // side 0, single chain only: out = max(out, 0)
    const float* dscale;         // side 0 epilogue: out = dscale * acc + selfcoef * hself
    const float* selfcoef;
    const float* hself;
    int64_t ld_hself;
};

struct D16Item {
    int m;            // destination tile (pair of tiles) of the item
    int split;        // side 1: slab
    int kb0, kb1;     // k-blocks
    bool add_only;    // the destination rows were zeroed by the host: every chain reduce-adds
};
__device__ __forceinline__ D16Item d16_item(const D16Params& p, int item, int m_items) {
    D16Item w;
    if (item < p.full_items) {
        w.m = item % m_items; w.split = item / m_items;
        w.kb0 = w.split * p.kb_per_split; w.kb1 = min(p.num_kb, w.kb0 + p.kb_per_split);
        w.add_only = false;
    } else {
        const int j = item - p.full_items;
        w.m = p.full_items + j % p.tail_tiles; w.split = 0;
        w.kb0 = (j / p.tail_tiles) * p.tail_kb; w.kb1 = min(p.num_kb, w.kb0 + p.tail_kb);
        w.add_only = true;
    }
    return w;
}

// ---- CTA-pair (cta_group::2) variants of the TMA / MMA / commit instructions -------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {       // same offset in CTA `rank` of the cluster
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
// loads into THIS CTA's shared memory, transaction bytes counted on an mbarrier that may live in the peer CTA (cluster address)
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, int x, int y, uint32_t bar_cluster_addr, uint64_t pol) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(bar_cluster_addr), "l"(pol) : "memory");
}
__device__ __forceinline__ void tma_load_3d_pair(void* dst, const CUtensorMap* map, int x, int y, int z, uint32_t bar_cluster_addr, uint64_t pol) {
    asm volatile("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(z), "r"(bar_cluster_addr), "l"(pol) : "memory");
}
// D[tmem of both CTAs] (+)= A * B^T with M = 256 over the pair: each CTA contributes its 128 rows of A and its half of B's N rows
__device__ __forceinline__ void tc_mma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
// arrives (once all prior MMAs of this thread have retired) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
// kind::f16 instruction descriptor with a_major / b_major given separately and M = m
__device__ __forceinline__ uint32_t umma_idesc_f16_ab(int m, int n, int bf16, int a_mn_major, int b_mn_major) {
    return (1u << 4) | ((uint32_t)bf16 << 7) | ((uint32_t)bf16 << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// PAIR = false: one CTA per 128-row destination tile (cta_group::1).
// PAIR = true : clusters of two CTAs on the SMs of one TPC work on two neighbouring destination tiles as ONE M = 256 MMA
//               (cta_group::2): every CTA stages its own 128 rows of A but only HALF of B's N rows, and the tensor cores of both
//               SMs read both halves.  Per SM that is 41 KB instead of 67 KB of operands per k-block through L2 -> SM and through
//               shared memory, which is what bounds the single-CTA form (ncu: 9.0 TB/s L2 -> SM, ~123 of 128 B/clk/SM of shared
//               memory at 77 % tensor-pipe activity).  Rank 0 of the pair issues the MMAs; both ranks run a TMA producer
//               (transaction bytes land on rank 0's barrier) and four drain warps for their own half of the accumulator.
template <bool PAIR>
__global__ void __launch_bounds__(kD16Threads, 1)
dense16_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
               const __grid_constant__ CUtensorMap map_b1_hi, const __grid_constant__ CUtensorMap map_b1_lo,
               const __grid_constant__ CUtensorMap map_b2_hi, const __grid_constant__ CUtensorMap map_b2_lo,
               const __grid_constant__ CUtensorMap map_out, const D16Params p) {
    extern __shared__ unsigned char dsmem_raw[];
    unsigned char* ring = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(dsmem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char* staging = ring + (size_t)p.stages * p.stage_bytes;          // 1024-byte aligned (stage_bytes % 1024 == 0)
    uint64_t* bars = reinterpret_cast<uint64_t*>(staging + kD16DrainWarps * p.nbuf * kD16BufBytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + kD16MaxStages;
    uint64_t* tmem_full = bars + 2 * kD16MaxStages;
    uint64_t* tmem_empty = bars + 2 * kD16MaxStages + 1;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kD16MaxStages + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int rank = PAIR ? (int)cluster_ctarank() : 0;
    const int group = PAIR ? blockIdx.x >> 1 : blockIdx.x;                    // work-item stream of this CTA (pair)
    const int n_groups = PAIR ? gridDim.x >> 1 : gridDim.x;
    const int tiles_per_item = PAIR ? 2 : 1;
    const int m_items = (p.m_tiles + tiles_per_item - 1) / tiles_per_item;     // destination tiles (pairs of tiles) per split
    const int n_items = p.full_items + p.tail_tiles * p.tail_splits;
    const int a_planes = p.terms == 3 ? 2 : 1;
    // rows of B (= output columns) this CTA stages for the two MMA column groups
    const int h1 = PAIR ? p.n1 / 2 : p.n1, h2 = PAIR ? p.n2 / 2 : p.n2;

    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(tmem_full, 1);
        mbar_init(tmem_empty, PAIR ? 2 * kD16DrainWarps : kD16DrainWarps);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(kTcMaxN));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(kTcMaxN));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
    }
    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ------------------------------------ TMA producer ------------------------------------
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a_hi) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b1_hi) : "memory");
            const uint64_t pol_a = l2_policy_evict_first(), pol_b = l2_policy_evict_last();
            int it = 0;
            for (int item = group; item < n_items; item += n_groups) {
                const D16Item w = d16_item(p, item, m_items);
                const int mt = w.m * tiles_per_item + rank, kb0 = w.kb0, kb1 = w.kb1;
                for (int kb = kb0; kb < kb1; ++kb, ++it) {
                    const int s = it % p.stages;
                    const uint32_t ph = (it / p.stages) & 1;
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    unsigned char* st = ring + (size_t)s * p.stage_bytes;
                    unsigned char* sb = st + a_planes * kD16ABytes;
                    // the stage's bytes of BOTH CTAs are counted on rank 0's barrier
                    const uint32_t bar = PAIR ? mapa_u32(smem_u32(&full_bar[s]), 0) : smem_u32(&full_bar[s]);
                    if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], (uint32_t)(PAIR ? 2 * p.tx_bytes : p.tx_bytes));
                    // A: this CTA's 128 destinations
                    if (p.side == 0) {
                        const int row = (mt * p.nb + kb) * kD16TileM;          // cell tile mt, gene block kb: 8 KB contiguous
                        if (PAIR) {
                            tma_load_2d_pair(st, &map_a_hi, 0, row, bar, pol_a);
                            if (p.terms == 3) tma_load_2d_pair(st + kD16ABytes, &map_a_lo, 0, row, bar, pol_a);
                        } else {
                            tma_load_2d_hint(st, &map_a_hi, 0, row, &full_bar[s], pol_a);
                            if (p.terms == 3) tma_load_2d_hint(st + kD16ABytes, &map_a_lo, 0, row, &full_bar[s], pol_a);
                        }
                    } else {
                        // 32 cells of cell tile kb / 4, gene blocks 4 mt .. 4 mt + 3: four 2 KB pieces of one 32 KB region,
                        // fetched by ONE 3-D box {32 slots, 32 cells, 4 blocks} (the producer thread is a scarce resource:
                        // 34 two-dimensional boxes per stage held this side at 0.69 of the other one)
                        const int row = ((kb >> 2) * p.nb + 4 * mt) * kD16TileM + (kb & 3) * 32;
                        if (PAIR) {
                            tma_load_3d_pair(st, &map_a_hi, 0, row, 0, bar, pol_a);
                            if (p.terms == 3) tma_load_3d_pair(st + kD16ABytes, &map_a_lo, 0, row, 0, bar, pol_a);
                        } else {
                            tma_load_3d_hint(st, &map_a_hi, 0, row, 0, &full_bar[s], pol_a);
                            if (p.terms == 3) tma_load_3d_hint(st + kD16ABytes, &map_a_lo, 0, row, 0, &full_bar[s], pol_a);
                        }
                    }
                    // B = H^T in k-blocks [kb][ld_hb rows][32]: this CTA's rows of the two MMA column groups (all of them
                    // without a pair), each box one contiguous piece of h x 64 bytes
                    const int r1 = kb * p.ld_hb + rank * h1, r2 = kb * p.ld_hb + p.n1 + rank * h2;
                    if (PAIR) {
                        tma_load_2d_pair(sb, &map_b1_hi, 0, r1, bar, pol_b);
                        if (p.terms == 3) tma_load_2d_pair(sb + p.b_bytes, &map_b1_lo, 0, r1, bar, pol_b);
                        if (h2 > 0) {
                            tma_load_2d_pair(sb + h1 * 64, &map_b2_hi, 0, r2, bar, pol_b);
                            if (p.terms == 3) tma_load_2d_pair(sb + p.b_bytes + h1 * 64, &map_b2_lo, 0, r2, bar, pol_b);
                        }
                    } else {
                        tma_load_2d_hint(sb, &map_b1_hi, 0, r1, &full_bar[s], pol_b);
                        if (p.terms == 3) tma_load_2d_hint(sb + p.b_bytes, &map_b1_lo, 0, r1, &full_bar[s], pol_b);
                        if (h2 > 0) {
                            tma_load_2d_hint(sb + h1 * 64, &map_b2_hi, 0, r2, &full_bar[s], pol_b);
                            if (p.terms == 3) tma_load_2d_hint(sb + p.b_bytes + h1 * 64, &map_b2_lo, 0, r2, &full_bar[s], pol_b);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------ MMA issuer (rank 0 of a pair) -------------------
        if (lane == 0 && rank == 0) {
            const int mma_m = PAIR ? 2 * kD16TileM : kD16TileM;
            const uint32_t idesc1 = umma_idesc_f16_ab(mma_m, p.n1, p.bf16, p.side, 0);
            const uint32_t idesc2 = umma_idesc_f16_ab(mma_m, p.n2 > 0 ? p.n2 : 16, p.bf16, p.side, 0);
            // advance along K by one MMA (16 elements): A: 32 bytes inside the 64-byte row (K-major) / two 8-row atoms (MN-major);
            // B: 32 bytes inside its 64-byte row
            const uint64_t a_step = p.side == 0 ? (uint64_t)(32 >> 4) : (uint64_t)(1024 >> 4);
            const uint64_t b_step = (uint64_t)(32 >> 4);
            const uint64_t n2_off = (uint64_t)((h1 * 64) >> 4);
            int it = 0, chunk_no = 0;
            for (int item = group; item < n_items; item += n_groups) {
                const D16Item w = d16_item(p, item, m_items);
                const int kb0 = w.kb0, kb1 = w.kb1;
                for (int c0 = kb0; c0 < kb1; c0 += p.chunk_kb, ++chunk_no) {
                    const int c1 = min(kb1, c0 + p.chunk_kb);
                    mbar_wait(tmem_empty, (chunk_no & 1) ^ 1);             // the drain warps (of both CTAs) have emptied the accumulators
                    tc_fence_after();
                    for (int kb = c0; kb < c1; ++kb, ++it) {
                        const int s = it % p.stages;
                        const uint32_t ph = (it / p.stages) & 1;
                        mbar_wait(&full_bar[s], ph);
                        tc_fence_after();
                        unsigned char* st = ring + (size_t)s * p.stage_bytes;
                        unsigned char* sb = st + a_planes * kD16ABytes;
                        uint64_t a_hi, a_lo;
                        if (p.side == 0) { a_hi = umma_desc_sw64(st); a_lo = umma_desc_sw64(st + kD16ABytes); }
                        else { a_hi = umma_desc_mn_sw64(st); a_lo = umma_desc_mn_sw64(st + kD16ABytes); }
                        const uint64_t b_hi = umma_desc_sw64(sb), b_lo = umma_desc_sw64(sb + p.b_bytes);
#pragma unroll
                        for (int kk = 0; kk < kD16BlockK / kD16UmmaK; ++kk) {
                            const uint64_t ka = (uint64_t)kk * a_step, kbo = (uint64_t)kk * b_step;
                            const uint32_t acc0 = (kb != c0 || kk != 0) ? 1u : 0u;
                            if (PAIR) {
                                tc_mma_f16_pair(tmem_base, a_hi + ka, b_hi + kbo, idesc1, acc0);
                                if (p.terms == 3) {
                                    tc_mma_f16_pair(tmem_base, a_lo + ka, b_hi + kbo, idesc1, 1);
                                    tc_mma_f16_pair(tmem_base, a_hi + ka, b_lo + kbo, idesc1, 1);
                                }
                                if (p.n2 > 0) {
                                    tc_mma_f16_pair(tmem_base + p.n1, a_hi + ka, b_hi + n2_off + kbo, idesc2, acc0);
                                    if (p.terms == 3) {
                                        tc_mma_f16_pair(tmem_base + p.n1, a_lo + ka, b_hi + n2_off + kbo, idesc2, 1);
                                        tc_mma_f16_pair(tmem_base + p.n1, a_hi + ka, b_lo + n2_off + kbo, idesc2, 1);
                                    }
                                }
                            } else {
                                tc_mma_f16(tmem_base, a_hi + ka, b_hi + kbo, idesc1, acc0);
                                if (p.terms == 3) {
                                    tc_mma_f16(tmem_base, a_lo + ka, b_hi + kbo, idesc1, 1);
                                    tc_mma_f16(tmem_base, a_hi + ka, b_lo + kbo, idesc1, 1);
                                }
                                if (p.n2 > 0) {
                                    tc_mma_f16(tmem_base + p.n1, a_hi + ka, b_hi + n2_off + kbo, idesc2, acc0);
                                    if (p.terms == 3) {
                                        tc_mma_f16(tmem_base + p.n1, a_lo + ka, b_hi + n2_off + kbo, idesc2, 1);
                                        tc_mma_f16(tmem_base + p.n1, a_hi + ka, b_lo + n2_off + kbo, idesc2, 1);
                                    }
                                }
                            }
                        }
                        if (PAIR) tc_commit_pair(&empty_bar[s]); else tc_commit(&empty_bar[s]);
                    }
                    if (PAIR) tc_commit_pair(tmem_full); else tc_commit(tmem_full);
                }
            }
        }
    } else {
        // ------------------------------------ drain warps -------------------------------------
        const int q = warp & 3;                                   // TMEM lane quarter of this warp
        constexpr int kShare = kD16DrainWarps / 4;                // warps per quarter
        const int cb_first = (warp - 2) / 4;                      // this warp's column blocks: cb_first, cb_first + kShare, ...
        unsigned char* my_stage = staging + (warp - 2) * (p.nbuf * kD16BufBytes);
        const int sw = (lane >> 1) & 3;                           // 64-byte swizzle: 16-byte chunk j of row r sits at j ^ ((r >> 1) & 3)
        const uint32_t tmem_empty_addr = PAIR ? mapa_u32(smem_u32(tmem_empty), 0) : smem_u32(tmem_empty);
        float h_inv = p.x_amax ? ldexpf(1.f, -d16_scale_exp(*p.x_amax)) : p.x_scale_inv;
        if (p.amax) h_inv *= ldexpf(1.f, -d16_scale_exp(*p.amax));
        const uint64_t pol_out = l2_policy_evict_last();
        int chunk_no = 0;
        for (int item = group; item < n_items; item += n_groups) {
            const D16Item w = d16_item(p, item, m_items);
            const int mt = w.m * tiles_per_item + rank, split = w.split, kb0 = w.kb0, kb1 = w.kb1;
            const int64_t grow = (int64_t)mt * kD16TileM + q * 32 + lane;        // destination row of this thread
            const int out_row = (int)(split * p.rows_per_split + (int64_t)mt * kD16TileM + q * 32);
            const bool tile_ok = mt < p.m_tiles;                                  // the odd tile out of a pair computes nothing useful
            const bool row_ok = grow < p.m_total;
            float scale = h_inv, sc = 0.f;
            if (p.side == 0 && row_ok) {
                if (p.dscale) scale *= __ldg(p.dscale + grow);
                if (p.selfcoef) sc = __ldg(p.selfcoef + grow);
            }
            for (int c0 = kb0; c0 < kb1; c0 += p.chunk_kb, ++chunk_no) {
                const bool first = c0 == kb0 && (!w.add_only || kb0 == 0);     // carries the self-loop / bias terms
                const bool store = c0 == kb0 && !w.add_only;                     // nothing to add to yet
                mbar_wait(tmem_full, chunk_no & 1);
                tc_fence_after();
                // the previous chain's adds to these rows have been performed (and the staging buffers are free)
                if (lane == 0) bulk_wait_all();
                __syncwarp();
                // One 16-column block: scale, epilogue terms, swizzled staging, TMA store / reduce-add.
                auto emit = [&](int cb, int nth, const uint32_t (&raw)[16]) {       // nth: count of this warp's blocks so far
                    float v[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(raw[i]) * scale;
                    if (nth >= p.nbuf) {                                 // the store issued nbuf blocks ago has read this buffer
                        if (lane == 0) bulk_wait_read(p.nbuf - 1);
                        __syncwarp();
                    }
                    if (first && p.selfcoef != nullptr && p.side == 0 && row_ok) {
                        const float* hrow = p.hself + grow * p.ld_hself + cb * kD16OutCols;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            if (cb * kD16OutCols + j * 4 < p.n) {
                                const float4 h = __ldg(reinterpret_cast<const float4*>(hrow + j * 4));
                                v[j * 4 + 0] = fmaf(sc, h.x, v[j * 4 + 0]); v[j * 4 + 1] = fmaf(sc, h.y, v[j * 4 + 1]);
                                v[j * 4 + 2] = fmaf(sc, h.z, v[j * 4 + 2]); v[j * 4 + 3] = fmaf(sc, h.w, v[j * 4 + 3]);
                            }
                        }
                    }
                    if (first && p.bias != nullptr) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            if (cb * kD16OutCols + j * 4 < p.n) {
                                const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + cb * kD16OutCols + j * 4));
                                v[j * 4 + 0] += b.x; v[j * 4 + 1] += b.y; v[j * 4 + 2] += b.z; v[j * 4 + 3] += b.w;
                            }
                        }
                    }
                    if (p.relu) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
                    }
                    unsigned char* buf = my_stage + (nth % p.nbuf) * kD16BufBytes;
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        *reinterpret_cast<float4*>(buf + lane * 64 + ((j ^ sw) << 4)) = make_float4(v[j * 4], v[j * 4 + 1], v[j * 4 + 2], v[j * 4 + 3]);
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
#ifdef WSAGE_TUNING     // drain diagnostics (results wrong): 1 = no TMA at all, 2 = plain stores instead of reduce-adds
                        if (p.drain_diag == 1) return;
                        if (p.drain_diag == 2) { tma_store_2d(&map_out, buf, cb * kD16OutCols, out_row, pol_out); bulk_commit_group(); return; }
#endif
                        if (store) tma_store_2d(&map_out, buf, cb * kD16OutCols, out_row, pol_out);
                        else tma_reduce_add_2d(&map_out, buf, cb * kD16OutCols, out_row, pol_out);
                        bulk_commit_group();
                    }
                };
                // software pipeline over this warp's column blocks: the tcgen05.ld of the next block is in flight while a block is
                // staged and stored.  (With four drain warps a drain cost ~6.8 us per 128 x 400 tile — 10 % / 5 % of a cell- /
                // gene-destination pass at 2048-row chains, 80 % of a K = 400 dense-layer tile — during which the tensor pipe idles:
                // a second accumulator would need 800 TMEM columns.  More stores in flight did not shorten it and without any TMA
                // traffic it still cost 2/3 of that: the per-block latency chain of one warp.  Hence two warps per lane quarter.)
                if (tile_ok) {
                    const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16);
                    const int n_cb = p.n_pad / kD16OutCols;
                    uint32_t ra[16], rb[16];
                    if (cb_first < n_cb) tmem_ld16_issue(t0 + (uint32_t)(cb_first * kD16OutCols), ra);
                    int nth = 0;
                    for (int cb = cb_first; cb < n_cb; cb += 2 * kShare, nth += 2) {
                        const int cb1 = cb + kShare, cb2 = cb + 2 * kShare;
                        tmem_ld_wait(ra);
                        if (cb1 < n_cb) tmem_ld16_issue(t0 + (uint32_t)(cb1 * kD16OutCols), rb);
                        emit(cb, nth, ra);
                        if (cb1 < n_cb) {
                            tmem_ld_wait(rb);
                            if (cb2 < n_cb) tmem_ld16_issue(t0 + (uint32_t)(cb2 * kD16OutCols), ra);
                            emit(cb1, nth + 1, rb);
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if (PAIR) mbar_arrive_cluster(tmem_empty_addr); else mbar_arrive(tmem_empty);
                }
            }
        }
        if (lane == 0) bulk_wait_all();
    }

    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();       // the peer may still read this CTA's shared memory until its last MMA
    if (warp == 1) {
        tc_fence_after();
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTcMaxN));
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTcMaxN));
    }
}

// Sum of the side-1 split partials (+ the tiled kernel's epilogue) when the CSR remainder of the pass is empty:
// one warp per destination row; slot = map[v] (< 0: the row has no dense entries) or v.
__global__ void __launch_bounds__(256)
dense16_finalize_kernel(const TiledParams p) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * 8;
    for (int64_t v = warp0; v < p.n_dst; v += nwarps) {
        RowAcc<0> acc;
        acc.zero();
        const int64_t slot = p.init_map ? (int64_t)__ldg(p.init_map + v) : v;
        if (slot >= 0) {
            for (int s = 0; s < p.init_slabs; ++s) {
                const float* src = p.init + ((size_t)s * p.init_rows + slot) * p.dim;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int c = (j * 32 + lane) * 4;
                    if (c < p.dim) {
                        const float4 t = *reinterpret_cast<const float4*>(src + c);
                        acc.v4[j].x += t.x; acc.v4[j].y += t.y; acc.v4[j].z += t.z; acc.v4[j].w += t.w;
                    }
                }
            }
        }
        tiled_row_epilogue<0>(p, v, acc, lane);
    }
}

// ------------------------------------------ host side ------------------------------------------
inline int make_map_2d(CUtensorMap* map, CUtensorMapDataType dt, int elem_bytes, const void* base, uint64_t inner, uint64_t rows,
                       uint64_t pitch_bytes, uint32_t box_inner, uint32_t box_rows, const char* what) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return fail(WSAGE_ECUDA, "%s: %s", what, "cuTensorMapEncodeTiled not available");
    const cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)pitch_bytes};
    const cuuint32_t box[2] = {box_inner, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    (void)elem_bytes;
    const CUresult r = enc(map, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(WSAGE_ECUDA, "%s: %s", what, "cuTensorMapEncodeTiled failed");
    return WSAGE_OK;
}

// Three-dimensional view {32 elements, rows (pitch row_pitch), blocks (pitch block_pitch)} of a 16-bit matrix, box
// {32, 32, box_blocks}: lands in shared memory as [block][row][32 elements], i.e. box_blocks MN-major SW64 operand blocks.
inline int make_map_3d(CUtensorMap* map, CUtensorMapDataType dt, const void* base, uint64_t rows, uint64_t blocks,
                       uint64_t row_pitch_bytes, uint64_t block_pitch_bytes, uint32_t box_blocks, const char* what) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return fail(WSAGE_ECUDA, "%s: %s", what, "cuTensorMapEncodeTiled not available");
    const cuuint64_t dims[3] = {32, (cuuint64_t)rows, (cuuint64_t)blocks};
    const cuuint64_t strides[2] = {(cuuint64_t)row_pitch_bytes, (cuuint64_t)block_pitch_bytes};
    const cuuint32_t box[3] = {32, 32, box_blocks};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = enc(map, dt, 3, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(WSAGE_ECUDA, "%s: %s", what, "cuTensorMapEncodeTiled (3-D) failed");
    return WSAGE_OK;
}

// slots are padded to 256: a CTA pair works on two whole tiles of 128 gene slots
inline int d16_slots_pad(int gene_slots) { return (gene_slots + 2 * kD16TileM - 1) / (2 * kD16TileM) * (2 * kD16TileM); }

struct D16Plan {
    bool pair;
    int m_tiles, nb, num_kb, chunk_kb, n_splits, kb_per_split;
    int full_items, tail_tiles, tail_splits, tail_kb;      // see D16Params
    int n_pad, n1, n2, h1, h2, b_bytes, stage_bytes, tx_bytes, stages, nbuf;
    size_t smem_bytes;
};

// WSAGE_D16_PAIR=0 (env, tuning only) selects the single-CTA form
inline bool d16_pair_env() {
    static const bool v = [] { const char* e = getenv("WSAGE_D16_PAIR"); return !(e && atoi(e) == 0); }();
    return v;
}

// WSAGE_D16_TAIL=0 (env, tuning only) keeps the last partial round of side-0 tiles whole
inline bool d16_tail_env() {
    static const bool v = [] { const char* e = getenv("WSAGE_D16_TAIL"); return !(e && atoi(e) == 0); }();
    return v;
}

// cut the k range into splits of whole chains so that the (tile, split) work items fill the SMs (or SM pairs) evenly
inline void d16_choose_splits(int m_items, int units, int num_kb, int chunk_kb, int& n_splits, int& kb_per_split) {
    const int chunks = (num_kb + chunk_kb - 1) / chunk_kb;
    int best = 1;
    double best_eff = -1.0;
    const int max_splits = chunks < 64 ? chunks : 64;
    for (int s = 1; s <= max_splits; ++s) {
        const int cps = (chunks + s - 1) / s;
        const int ns = (chunks + cps - 1) / cps;
        if (ns != s) continue;
        const int64_t items = (int64_t)m_items * ns;
        const int64_t rounds = (items + units - 1) / units;
        const double eff = (double)m_items * chunks / ((double)units * rounds * cps);
        if (eff > best_eff + 0.02) { best_eff = eff; best = s; }           // fewer splits unless clearly better
    }
    const int cps = (chunks + best - 1) / best;
    n_splits = (chunks + cps - 1) / cps;
    kb_per_split = cps * chunk_kb;
}

inline int d16_plan(const wsage_dense16_args* a, D16Plan& pl) {
    const int terms = a->fmt == WSAGE_D16_F16X2 ? 3 : 1;
    pl.pair = d16_pair_env();
    pl.nb = d16_slots_pad(a->gene_slots) / kD16BlockK;
    int chunk_rows = a->chunk_rows > 0 ? a->chunk_rows : kD16DefaultChunkRows;
    pl.chunk_kb = (chunk_rows + kD16BlockK - 1) / kD16BlockK;
    pl.n_pad = (a->dim + 15) & ~15;
    pl.n1 = pl.n_pad < 256 ? pl.n_pad : 256;
    pl.n2 = pl.n_pad - pl.n1;
    pl.h1 = pl.pair ? pl.n1 / 2 : pl.n1;          // rows of B a CTA stages for each MMA column group
    pl.h2 = pl.pair ? pl.n2 / 2 : pl.n2;
    pl.b_bytes = (pl.h1 + pl.h2) * 64;
    const int tiles_per_item = pl.pair ? 2 : 1;
    const int units = pl.pair ? kNumSMs / 2 : kNumSMs;
    pl.full_items = -1; pl.tail_tiles = pl.tail_splits = pl.tail_kb = 0;
    if (a->side == 0) {
        pl.m_tiles = (int)((a->n_dst + kD16TileM - 1) / kD16TileM);
        pl.num_kb = (a->gene_slots + kD16BlockK - 1) / kD16BlockK;
        pl.n_splits = 1;
        pl.kb_per_split = pl.num_kb;
        // the last round of tiles is partial (c4: 2969 pairs of tiles over 74 SM pairs = 40.1 rounds, a shard of it at N = 8:
        // 5.02): cut its tiles along k so that the round ends after 1 / tail_splits of a tile instead of a whole one
        const int m_items = (pl.m_tiles + tiles_per_item - 1) / tiles_per_item;
        const int chunks = (pl.num_kb + pl.chunk_kb - 1) / pl.chunk_kb;
        const int tail = m_items % units;
        if (d16_tail_env() && !a->deterministic && tail > 0 && chunks >= 2 && !a->relu) {
            int ts = units / tail;
            if (ts > chunks) ts = chunks;
            const int cps = (chunks + ts - 1) / ts;
            ts = (chunks + cps - 1) / cps;
            if (ts >= 2) {
                pl.full_items = m_items - tail; pl.tail_tiles = tail; pl.tail_splits = ts; pl.tail_kb = cps * pl.chunk_kb;
            }
        }
    } else {
        pl.m_tiles = d16_slots_pad(a->gene_slots) / kD16TileM;
        pl.num_kb = (int)((a->n_src_cells + kD16BlockK - 1) / kD16BlockK);
        d16_choose_splits((pl.m_tiles + tiles_per_item - 1) / tiles_per_item, units, pl.num_kb, pl.chunk_kb, pl.n_splits, pl.kb_per_split);
    }
    if (pl.full_items < 0) pl.full_items = (pl.m_tiles + tiles_per_item - 1) / tiles_per_item * pl.n_splits;
    pl.tx_bytes = (terms == 3 ? 2 : 1) * (kD16ABytes + pl.b_bytes);
    pl.stage_bytes = (pl.tx_bytes + 1023) & ~1023;
    // staging buffers of the drain warps.  A drain of a 128 x 400 tile costs ~6.8 us (10 % of a cell-destination pass at
    // 2048-row chains); more TMA stores in flight do not shorten it (WSAGE_D16_BUFS = 2 / 4 / 6: 3.61 / 3.61 / 3.60 ms at c3), so
    // the default stays at two buffers and the ring keeps its fifth stage.
    static const int forced = [] { const char* e = getenv("WSAGE_D16_BUFS"); return e ? atoi(e) : 0; }();
    pl.nbuf = (forced >= 2 && forced <= kD16MaxBufs) ? forced : 2;
    const int fixed = kD16DrainWarps * pl.nbuf * kD16BufBytes + 256 + 1024;
    pl.stages = (kD16SmemBudget - fixed) / pl.stage_bytes;
    if (pl.stages > kD16MaxStages) pl.stages = kD16MaxStages;
    pl.smem_bytes = (size_t)pl.stages * pl.stage_bytes + fixed;
    return WSAGE_OK;
}

}  // namespace wsage
