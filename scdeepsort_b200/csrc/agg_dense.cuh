// Dense-block companion of the tiled aggregation kernel (sm_100a).
//
// Gene popularity is heavily skewed (housekeeping genes are expressed in nearly every cell; in the
// synthetic atlas 9 % of the genes hold 43 % of the edges at >= 20 % density).  For those genes the
// CSR walk of agg_tiled_kernel pays 14 shared-memory wavefronts per EDGE for a source row that almost
// every destination row of the tile wants anyway.  The graph builder therefore splits the expression
// matrix into  X = X_sparse + X_dense : the popular genes' entries are stored as a dense, zero-filled
// block and handled here; agg_tiled_kernel walks only the sparse remainder and starts its accumulators
// from this kernel's sums (TiledParams::init), so the epilogue and every output stay where they were.
//
//   dout[t, :] = SUM_k xd[k, t] * hs[src(k), :]            t < t_total destinations, k < K sources
//
// Same CTA organisation as the tiled kernel — NW consumer warps x TM destination rows with register-resident
// accumulators, one producer warp streaming source rows through a shared-memory ring with bulk-async
// copies — but there is no edge list: for every source row of the window a warp reads the row once
// (13 wavefronts for 400 floats) and its TM weights with broadcast loads, then issues TM x 13 FMAs.
// With TM = 6 that is 16 wavefronts (16 clk of the LSU pipe per SM) against 78 FMA instructions
// (19.5 clk on four sub-partitions): the kernel is bound by the fp32 FMA pipe, not by shared memory.
//
// xd is TILE-BLOCKED by the caller: [n_tiles][K][T] with T = kDenseT destinations per tile, zero padded,
// so a CTA's weights for a window are one contiguous bulk copy.  Two uses (wsage_spmm picks by which of
// dense_src_ids / dense_dst_map is given):
//   destinations = popular genes, sources = all cells  (gene<-cell passes; split over the source range,
//                  partial sums per split, added in fixed order by the tiled kernel's prologue)
//   destinations = all cells, sources = popular genes   (cell<-gene passes; src(k) = dense_src_ids[k])
#pragma once
#include "agg_tiled.cuh"

namespace wsage {

constexpr int kDenseNW = 14;                       // consumer warps (+ producer = 15 warps at <= 128 registers); T must be a multiple of 4
constexpr int kDenseTM = 6;                        // destination rows per warp
constexpr int kDenseT = kDenseNW * kDenseTM;       // destinations per tile (the blocking of xd)
constexpr int kDenseStages = 3;                    // default ring depth (WSAGE_DENSE_STAGES = 2 | 3 | 4 for tuning)
constexpr int kDenseMaxSplits = 64;

struct DenseParams {
    const float* xd;            // [n_tiles][K][T]
    const int32_t* src_ids;     // [K] rows of hs, or null (source k = row k)
    const float* hs;            // contiguous [*, dim]
    int64_t K;
    int64_t t_total;
    int dim;
    int pitch;                  // floats between staged source rows (dim rounded up to 32)
    int win_rows;
    int n_windows;
    int n_tiles;
    int n_splits;
    int win_per_split;
    float* dout;                // [n_splits][t_total][dim]
};

template <int DIM, int STG>
__global__ void __launch_bounds__((kDenseNW + 1) * 32, 1)
agg_dense_kernel(const DenseParams p) {
    using S = RowShape<DIM>;
    constexpr int NW = kDenseNW, TM = kDenseTM, T = kDenseT;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t full_bar[STG];
    __shared__ uint64_t empty_bar[STG];

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int tile = blockIdx.x % p.n_tiles;          // split-major: resident CTAs share the source slice
    const int split = blockIdx.x / p.n_tiles;
    const int w_begin = split * p.win_per_split;
    const int w_end = min(p.n_windows, w_begin + p.win_per_split);
    const int dim = DIM > 0 ? DIM : p.dim;
    const int pitch = DIM > 0 ? ((DIM + 31) & ~31) : p.pitch;
    const size_t stage_floats = (size_t)p.win_rows * (pitch + T);      // source rows, then the weights [W][T]
    float* stages = reinterpret_cast<float*>(smem_raw);

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < STG; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], NW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == NW) {
        // ------------------------------- producer -------------------------------------------
        const uint32_t row_bytes = (uint32_t)(dim * sizeof(float));
        const float* xd_tile = p.xd + (size_t)tile * p.K * T;
        for (int w = w_begin, it = 0; w < w_end; ++w, ++it) {
            const int s = it % STG;
            const uint32_t ph = (it / STG) & 1;
            mbar_wait(&empty_bar[s], ph ^ 1);
            const int64_t k0 = (int64_t)w * p.win_rows;
            const int rows = (int)min((int64_t)p.win_rows, p.K - k0);
            float* st = stages + s * stage_floats;
            if (lane == 0) {
                mbar_arrive_expect_tx(&full_bar[s], rows * (row_bytes + T * (uint32_t)sizeof(float)));
                bulk_g2s(st + (size_t)p.win_rows * pitch, xd_tile + k0 * T, rows * T * (uint32_t)sizeof(float), &full_bar[s]);
            }
            __syncwarp();
            for (int r = lane; r < rows; r += 32) {
                const int64_t src = p.src_ids ? (int64_t)__ldg(p.src_ids + k0 + r) : k0 + r;
                bulk_g2s(st + (size_t)r * pitch, p.hs + src * dim, row_bytes, &full_bar[s]);
            }
        }
        return;
    }

    // ----------------------------------- consumers ------------------------------------------
    RowAcc<DIM> acc[TM];
#pragma unroll
    for (int t = 0; t < TM; ++t) acc[t].zero();

    for (int w = w_begin, it = 0; w < w_end; ++w, ++it) {
        const int s = it % STG;
        const uint32_t ph = (it / STG) & 1;
        const int rows = (int)min((int64_t)p.win_rows, p.K - (int64_t)w * p.win_rows);
        const float* hrow = stages + s * stage_floats + lane * 4;
        const float* xrow = stages + s * stage_floats + (size_t)p.win_rows * pitch + warp * TM;
        mbar_wait(&full_bar[s], ph);
#pragma unroll 2
        for (int k = 0; k < rows; ++k) {
            const float* h = hrow + (size_t)k * pitch;
            float4 a4[S::N4];
            float a1 = 0.f;
#pragma unroll
            for (int j = 0; j < S::N4; ++j)
                if (S::on4(j, lane, dim)) a4[j] = *reinterpret_cast<const float4*>(h + j * 128);
            if (S::TAIL1 && S::on1(lane)) a1 = h[S::J4 * 128 - lane * 3];
            float x[TM];
#pragma unroll
            for (int t = 0; t < TM; t += 2) {            // warp-uniform address: one broadcast wavefront each
                const float2 v = *reinterpret_cast<const float2*>(xrow + (size_t)k * T + t);
                x[t] = v.x; x[t + 1] = v.y;
            }
#pragma unroll
            for (int t = 0; t < TM; ++t) {
#pragma unroll
                for (int j = 0; j < S::N4; ++j)
                    if (S::on4(j, lane, dim)) Vec<4>::fma(acc[t].v4[j], x[t], a4[j]);
                if (S::TAIL1) acc[t].v1 = fmaf(x[t], a1, acc[t].v1);
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty_bar[s]);
    }

#pragma unroll
    for (int t = 0; t < TM; ++t) {
        const int64_t slot = (int64_t)tile * T + warp * TM + t;
        if (slot >= p.t_total) continue;
        float* dst = p.dout + ((size_t)split * p.t_total + slot) * dim;
#pragma unroll
        for (int j = 0; j < S::N4; ++j)
            if (S::on4(j, lane, dim)) *reinterpret_cast<float4*>(dst + (j * 32 + lane) * 4) = acc[t].v4[j];
        if (S::TAIL1 && S::on1(lane)) dst[S::J4 * 128 + lane] = acc[t].v1;
    }
}

// ------------------------------------------ host side ------------------------------------------
inline int dense_stages() {
    static const int v = [] {
        const char* e = getenv("WSAGE_DENSE_STAGES");
        const int i = e ? atoi(e) : kDenseStages;
        return (i >= 2 && i <= 4) ? i : kDenseStages;
    }();
    return v;
}

struct DensePlan {
    int stages, win_rows, pitch, n_windows, n_tiles, n_splits, win_per_split;
    size_t smem_bytes, out_bytes;
};

inline bool dense_requested(const wsage_spmm_args* a) { return a->dense_x != nullptr; }

inline DensePlan dense_plan(const wsage_spmm_args* a) {
    DensePlan pl{};
    pl.pitch = (a->dim + 31) & ~31;
    const size_t per_row = (size_t)(pl.pitch + kDenseT) * sizeof(float);
    pl.stages = dense_stages();
    int w = (int)(kTiledSmemBudget / pl.stages / per_row);
    w &= ~1;
    if ((int64_t)w > a->dense_k) w = (int)a->dense_k;
    if (w < 1) w = 1;
    pl.win_rows = w;
    pl.n_windows = (int)((a->dense_k + w - 1) / w);
    pl.n_tiles = (int)((a->dense_t + kDenseT - 1) / kDenseT);
    // few destination tiles (popular genes as destinations): cut the long source range so that the grid
    // covers the chip ~4 times over; many tiles (cells as destinations) need no split
    int splits = 1;
    // (rounded DOWN: 14 tiles x 43 splits = 602 CTAs would leave a fifth, nearly empty wave — measured 20.5 ms
    // against 15.3 ms for 4.0 waves)
    if (pl.n_tiles < 2 * kNumSMs) splits = (4 * kNumSMs) / pl.n_tiles;
    if (splits < 1) splits = 1;
    const int max_splits = pl.n_windows / 4 > 0 ? pl.n_windows / 4 : 1;
    if (splits > max_splits) splits = max_splits;
    if (splits > kDenseMaxSplits) splits = kDenseMaxSplits;
    pl.win_per_split = (pl.n_windows + splits - 1) / splits;
    pl.n_splits = (pl.n_windows + pl.win_per_split - 1) / pl.win_per_split;
    pl.smem_bytes = (size_t)pl.stages * w * per_row;
    pl.out_bytes = (size_t)pl.n_splits * a->dense_t * a->dim * sizeof(float);
    pl.out_bytes = (pl.out_bytes + 255) & ~(size_t)255;
    return pl;
}

template <int DIM, int STG>
int launch_dense_shape(const DenseParams& p, const DensePlan& pl, cudaStream_t st) {
    auto kern = agg_dense_kernel<DIM, STG>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bytes);
    if (e != cudaSuccess) return fail(WSAGE_ECUDA, "%s: %s", "cudaFuncSetAttribute(agg_dense)", cudaGetErrorString(e));
    kern<<<pl.n_tiles * pl.n_splits, (kDenseNW + 1) * 32, pl.smem_bytes, st>>>(p);
    return check_launch("agg_dense");
}

template <int DIM>
int launch_dense_dim(const DenseParams& p, const DensePlan& pl, cudaStream_t st) {
    if (DIM == 400) {                      // the bench width carries the tuning variants
        if (pl.stages == 2) return launch_dense_shape<DIM, 2>(p, pl, st);
        if (pl.stages == 4) return launch_dense_shape<DIM, 4>(p, pl, st);
    }
    return launch_dense_shape<DIM, kDenseStages>(p, pl, st);
}

// Runs the dense block into `dout` (dense_plan(a).out_bytes bytes of workspace).
inline int launch_dense(const wsage_spmm_args* a, const DensePlan& pl, float* dout, cudaStream_t st) {
    DenseParams p{};
    p.xd = a->dense_x; p.src_ids = a->dense_src_ids; p.hs = a->hs;
    p.K = a->dense_k; p.t_total = a->dense_t; p.dim = a->dim; p.pitch = pl.pitch;
    p.win_rows = pl.win_rows; p.n_windows = pl.n_windows; p.n_tiles = pl.n_tiles;
    p.n_splits = pl.n_splits; p.win_per_split = pl.win_per_split; p.dout = dout;
    switch (a->dim) {
        case 400: return launch_dense_dim<400>(p, pl, st);
        case 200: return launch_dense_dim<200>(p, pl, st);
        case 128: return launch_dense_dim<128>(p, pl, st);
        default:  return launch_dense_dim<0>(p, pl, st);
    }
}

}  // namespace wsage
