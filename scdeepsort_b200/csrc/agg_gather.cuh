// "gather" aggregation kernels: one warp per destination row, source rows gathered through
// L2 with 16-byte loads, fp32 accumulation in registers, no per-edge message tensor.
//
// Serves (a) generic NodeFlow blocks with the on-device alpha cascade (forward + atomic
// backward) and (b) the full-graph bipartite pass when source-row reuse per CTA is too low
// for the tiled kernel (agg_tiled.cuh).
#pragma once
#include "common.cuh"

namespace wsage {

struct GatherParams {
    const int64_t* rowptr;
    const void* col;
    const float* w;
    // alpha cascade (generic blocks only)
    const int32_t* src_id;
    const int32_t* dst_id;
    const float* alpha;
    int gene_num;
    const float* hs;
    int64_t ld_hs;
    int64_t n_dst;
    int dim;
    int mean;                 // out *= 1/max(deg,1)
    const float* dscale;
    const float* selfcoef;
    const float* hself;
    int64_t ld_hself;
    float* out;
    int64_t ld_out;
    float* raw;
    int64_t ld_raw;
    const float* q;
    int64_t ld_q;
    float* dot;
    const int32_t* row_perm;
};

constexpr int kGatherWarps = 8;   // 256 threads per CTA

// VEC: floats per lane per load (4 = 16-byte path, 1 = scalar fallback for odd dims / alignment)
// J:   column chunks per lane held in registers; one pass covers 32*VEC*J columns
template <int VEC, int J, typename ColT, bool CASCADE>
__global__ void __launch_bounds__(kGatherWarps * 32)
agg_gather_fwd_kernel(const GatherParams p) {
    using V = Vec<VEC>;
    using T = typename V::T;
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * kGatherWarps + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * kGatherWarps;
    const ColT* __restrict__ col = static_cast<const ColT*>(p.col);
    constexpr int kTile = 32 * VEC * J;

    for (int64_t r = warp0; r < p.n_dst; r += nwarps) {
        const int64_t v = p.row_perm ? (int64_t)p.row_perm[r] : r;
        const int64_t beg = p.rowptr[v], end = p.rowptr[v + 1];
        int did = -1;
        if (CASCADE) did = p.dst_id[v];
        float scale = 1.f;
        if (p.mean) scale = 1.f / (float)max((int64_t)1, end - beg);
        if (p.dscale) scale *= p.dscale[v];
        const float sc = p.selfcoef ? p.selfcoef[v] : 0.f;
        float dot = 0.f;

        for (int c0 = 0; c0 < p.dim; c0 += kTile) {
            T acc[J];
#pragma unroll
            for (int j = 0; j < J; ++j) acc[j] = V::zero();
            int cidx[J];
#pragma unroll
            for (int j = 0; j < J; ++j) cidx[j] = c0 + (j * 32 + lane) * VEC;

            for (int64_t base = beg; base < end; base += 32) {
                const int64_t e = base + lane;
                int64_t my_c = 0;
                float my_w = 0.f;
                if (e < end) {
                    my_c = (int64_t)col[e];
                    my_w = p.w[e];
                    if (CASCADE) my_w *= p.alpha[alpha_index(p.src_id[my_c], did, p.gene_num)];
                }
                const int n = (int)min((int64_t)32, end - base);
                for (int k = 0; k < n; k += 4) {
                    const float* rowp[4];
                    float wk[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {   // k + u <= 31; edges past the row end are skipped
                        const int64_t cu = __shfl_sync(0xffffffffu, my_c, k + u);
                        wk[u] = __shfl_sync(0xffffffffu, my_w, k + u);
                        rowp[u] = p.hs + cu * p.ld_hs;
                    }
                    T val[4][J];
#pragma unroll
                    for (int u = 0; u < 4; ++u)
#pragma unroll
                        for (int j = 0; j < J; ++j)
                            val[u][j] = (k + u < n && cidx[j] < p.dim) ? V::ldg(rowp[u] + cidx[j]) : V::zero();
#pragma unroll
                    for (int u = 0; u < 4; ++u)
#pragma unroll
                        for (int j = 0; j < J; ++j) V::fma(acc[j], wk[u], val[u][j]);
                }
            }
            // epilogue for this column tile
#pragma unroll
            for (int j = 0; j < J; ++j) {
                if (cidx[j] >= p.dim) continue;
                if (p.raw) V::st(p.raw + v * p.ld_raw + cidx[j], acc[j]);
                if (p.dot) dot += V::dot(acc[j], V::ldg(p.q + v * p.ld_q + cidx[j]));
                if (p.out) {
                    T o = V::scale(scale, acc[j]);
                    if (p.selfcoef) V::fma(o, sc, V::ldg(p.hself + v * p.ld_hself + cidx[j]));
                    V::st(p.out + v * p.ld_out + cidx[j], o);
                }
            }
        }
        if (p.dot) {
            dot = warp_sum(dot);
            if (lane == 0) p.dot[v] = dot;
        }
    }
}

struct GatherBwdParams {
    const int64_t* rowptr;
    const int32_t* col;
    const float* w;
    const int32_t* src_id;
    const int32_t* dst_id;
    const float* alpha;
    int gene_num;
    const float* hs;
    int64_t ld_hs;
    const float* dout;
    int64_t ld_dout;
    int64_t n_dst;
    int dim;
    float* dh;
    int64_t ld_dh;
    float* dalpha;
};

// Backward of the generic block pass.  Per edge: dh[col] += s_v*w*alpha[k]*dout[v] (vector
// atomics) and dalpha[k] += s_v*w*<hs[col], dout[v]> (warp reduction, one atomic per edge).
template <int VEC, int J>
__global__ void __launch_bounds__(kGatherWarps * 32)
agg_gather_bwd_kernel(const GatherBwdParams p) {
    using V = Vec<VEC>;
    using T = typename V::T;
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * kGatherWarps + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * kGatherWarps;
    constexpr int kTile = 32 * VEC * J;

    for (int64_t v = warp0; v < p.n_dst; v += nwarps) {
        const int64_t beg = p.rowptr[v], end = p.rowptr[v + 1];
        if (end == beg) continue;
        const int did = p.dst_id[v];
        const float sv = 1.f / (float)(end - beg);
        for (int c0 = 0; c0 < p.dim; c0 += kTile) {
            T g[J];
            int cidx[J];
#pragma unroll
            for (int j = 0; j < J; ++j) {
                cidx[j] = c0 + (j * 32 + lane) * VEC;
                g[j] = (cidx[j] < p.dim) ? V::scale(sv, V::ldg(p.dout + v * p.ld_dout + cidx[j])) : V::zero();
            }
            for (int64_t base = beg; base < end; base += 32) {
                const int64_t e = base + lane;
                int my_c = 0, my_k = 0;
                float my_w = 0.f;
                if (e < end) {
                    my_c = p.col[e];
                    my_w = p.w[e];
                    my_k = alpha_index(p.src_id[my_c], did, p.gene_num);
                }
                const int n = (int)min((int64_t)32, end - base);
                for (int k = 0; k < n; ++k) {
                    const int64_t cu = __shfl_sync(0xffffffffu, my_c, k);
                    const float wu = __shfl_sync(0xffffffffu, my_w, k);
                    const int ku = __shfl_sync(0xffffffffu, my_k, k);
                    const float au = p.alpha[ku];
                    float dot = 0.f;
#pragma unroll
                    for (int j = 0; j < J; ++j) {
                        if (cidx[j] >= p.dim) continue;
                        if (p.dalpha) dot += V::dot(g[j], V::ldg(p.hs + cu * p.ld_hs + cidx[j]));
                        if (p.dh) V::atomic_add(p.dh + cu * p.ld_dh + cidx[j], V::scale(wu * au, g[j]));
                    }
                    if (p.dalpha) {
                        dot = warp_sum(dot);
                        if (lane == 0) atomicAdd(p.dalpha + ku, wu * dot);
                    }
                }
            }
        }
    }
}

inline int gather_grid(int64_t n_rows) {
    int64_t ctas = (n_rows + kGatherWarps - 1) / kGatherWarps;
    const int64_t cap = (int64_t)kNumSMs * 8;   // 8 resident CTAs of 256 threads per SM
    if (ctas > cap) ctas = cap;
    return (int)(ctas < 1 ? 1 : ctas);
}

template <typename ColT, bool CASCADE>
int launch_gather_fwd(const GatherParams& p, bool vec4, cudaStream_t st) {
    const int grid = gather_grid(p.n_dst);
    const int block = kGatherWarps * 32;
    if (vec4) {
        if (p.dim <= 128) agg_gather_fwd_kernel<4, 1, ColT, CASCADE><<<grid, block, 0, st>>>(p);
        else if (p.dim <= 256) agg_gather_fwd_kernel<4, 2, ColT, CASCADE><<<grid, block, 0, st>>>(p);
        else agg_gather_fwd_kernel<4, 4, ColT, CASCADE><<<grid, block, 0, st>>>(p);
    } else {
        if (p.dim <= 64) agg_gather_fwd_kernel<1, 2, ColT, CASCADE><<<grid, block, 0, st>>>(p);
        else agg_gather_fwd_kernel<1, 4, ColT, CASCADE><<<grid, block, 0, st>>>(p);
    }
    return check_launch("agg_gather_fwd");
}

inline int launch_gather_bwd(const GatherBwdParams& p, bool vec4, cudaStream_t st) {
    const int grid = gather_grid(p.n_dst);
    const int block = kGatherWarps * 32;
    if (vec4) {
        if (p.dim <= 256) agg_gather_bwd_kernel<4, 2><<<grid, block, 0, st>>>(p);
        else agg_gather_bwd_kernel<4, 4><<<grid, block, 0, st>>>(p);
    } else {
        agg_gather_bwd_kernel<1, 4><<<grid, block, 0, st>>>(p);
    }
    return check_launch("agg_gather_bwd");
}

}  // namespace wsage
