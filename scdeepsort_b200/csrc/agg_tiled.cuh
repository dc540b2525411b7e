// "tiled" aggregation kernel for the full-graph bipartite pass (sm_100a).
//
//   acc[v,:] = SUM_{e in row v} x[e] * hs[col[e],:]        (+ epilogue, see wsage_spmm)
//
// The gather kernel re-reads a source row from L2 once per EDGE.  Here a CTA owns a tile of
// NW*R destination rows whose accumulators stay in registers, and streams the source table
// through shared memory in windows of W consecutive rows: columns are sorted inside each CSR
// row, so the rows a window needs are exactly one contiguous [W, dim] slab of hs, fetched by a
// single bulk-async copy (cp.async.bulk, completion on an mbarrier; SASS UBLKCP) into a ring of
// S stages.  Every source row is thus read from L2 once per TILE and then served to all its
// edges from shared memory (16-byte conflict-free LDS).
//
//   warp NW (producer): one elected lane waits empty[s], arms full[s] with the byte count and
//                       issues the bulk copy of window w into stage s.
//   warps 0..NW-1     : each owns R destination rows; per window waits full[s], consumes the
//                       edges of its rows whose column falls in the window, arrives on empty[s].
//
// Scheduling.  A tile is NW*R CONSECUTIVE rows of row_perm (rows sorted by degree), so the
// warps of a CTA see the same edge density and stay in step on the stage ring; the degree skew
// (gene rows: 3 % .. 100 % dense) is absorbed across CTAs instead: the source dimension is cut
// into n_splits window ranges, CTAs are ordered split-major / heavy-tile-first, partial sums of
// split tiles go to the workspace and tiled_reduce_kernel adds them in fixed order
// (deterministic).  Split-major order also means the ~148 resident CTAs stream the SAME slice
// of the source table (<= ~40 MB, L2-resident) at the same time, so HBM sees each source row
// about once even when the table (cells: 1.2 GB) is far larger than L2.
//
// DIM is a compile-time feature width (0 = runtime width, predicated loads) so the hot loop has
// no per-chunk predicates: a 400-float row is 3 full float4 chunks per lane plus one scalar
// chunk on lanes 0..15 (13 shared-memory wavefronts per edge, the minimum for 1600 B).
#pragma once
#include <stdlib.h>

#include "common.cuh"

namespace wsage {

constexpr int kTiledStages = 4;                // deepest window ring among the compiled shapes
constexpr int kTiledSmemBudget = 224 * 1024;   // bytes of window ring per CTA (227 KB max per CTA)
constexpr int kTiledDefaultCluster = 1;        // CTAs per cluster sharing window loads (-DWSAGE_TUNING: WSAGE_TILED_CLUSTER overrides)

struct TiledParams {
    const int64_t* rowptr;
    const void* col;
    const float* x;
    const float* hs;          // contiguous [n_src, dim]
    int64_t n_src;
    int64_t n_dst;
    int dim;
    int win_rows;             // W
    int pitch;                // floats between consecutive rows of a stage: dim rounded up to 32 (128 B),
                              // so every 512-byte warp load touches exactly 4 shared-memory lines
    int n_windows;            // ceil(n_src / W)
    int n_tiles;
    int n_splits;
    int win_per_split;
    const int32_t* row_perm;
    int l2_prefetch;          // > 0: some CTAs prefetch the window this many windows ahead into L2
    // accumulators start from the dense block's sums (agg_dense.cuh): init[slab][slot][dim], slabs added in
    // index order by the split-0 CTA of the row; slot = init_map[row] (< 0: none) or the row itself
    const float* init;
    int init_slabs;
    int64_t init_rows;
    const int32_t* init_map;
    // epilogue (n_splits == 1) ...
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
    // ... or partial sums (n_splits > 1): partial[split][row][dim]
    float* partial;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// Same, on precomputed 32-bit shared-window addresses (keeps the generic->shared conversion out of hot loops).
__device__ __forceinline__ void mbar_wait_a(uint32_t bar_addr, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar_addr), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar_addr) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_addr) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// Bulk prefetch of a contiguous global range into L2 (no completion tracking).
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
// 1-D bulk async copy global -> shared, completion counted in bytes on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// --- thread-block-cluster helpers (CL = 2: two CTAs on neighbouring SMs share every window load) -----------
// The same bulk copy, written into the shared memory of every CTA in cta_mask at the same offset; each
// destination CTA's mbarrier (same offset) receives the complete_tx.
__device__ __forceinline__ void bulk_g2s_multicast(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint16_t cta_mask) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint64_t* bar, uint32_t cta_rank) {     // arrive on the peer CTA's barrier
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(smem_u32(bar)), "r"(cta_rank) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// How a DIM-float row is spread over the 32 lanes of a warp.
template <int DIM>
struct RowShape {
    static constexpr int J4 = DIM > 0 ? DIM / 128 : 0;            // float4 chunks used by all 32 lanes
    static constexpr int REM = DIM > 0 ? DIM - J4 * 128 : 0;      // floats left over
    static constexpr bool TAIL4 = REM > 32;                       // partial float4 chunk: lanes < REM/4
    static constexpr bool TAIL1 = REM > 0 && REM <= 32;           // scalar chunk: lanes < REM
    static constexpr int N4 = DIM > 0 ? J4 + (TAIL4 ? 1 : 0) : 4;  // runtime width: 4 predicated chunks (<= 512)
    static_assert(DIM % 4 == 0, "feature width must be a multiple of 4");

    static __device__ __forceinline__ bool on4(int j, int lane, int dim) {
        if (DIM > 0) return j < J4 || lane < REM / 4;
        return (j * 32 + lane) * 4 < dim;
    }
    static __device__ __forceinline__ bool on1(int lane) { return lane < REM; }
};

template <int DIM>
struct RowAcc {
    using S = RowShape<DIM>;
    float4 v4[S::N4];
    float v1;
    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int j = 0; j < S::N4; ++j) v4[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        v1 = 0.f;
    }
};

// Final per-row epilogue shared by the tiled kernel (n_splits == 1) and the split reducer.
template <int DIM>
__device__ __forceinline__ void tiled_row_epilogue(const TiledParams& p, int64_t v, const RowAcc<DIM>& a, int lane) {
    using S = RowShape<DIM>;
    const float scale = p.dscale ? p.dscale[v] : 1.f;
    const float sc = p.selfcoef ? p.selfcoef[v] : 0.f;
    float dot = 0.f;
#pragma unroll
    for (int j = 0; j < S::N4; ++j) {
        if (!S::on4(j, lane, p.dim)) continue;
        const int c = (j * 32 + lane) * 4;
        if (p.raw) *reinterpret_cast<float4*>(p.raw + v * p.ld_raw + c) = a.v4[j];
        if (p.dot) dot += Vec<4>::dot(a.v4[j], __ldg(reinterpret_cast<const float4*>(p.q + v * p.ld_q + c)));
        if (p.out) {
            float4 o = Vec<4>::scale(scale, a.v4[j]);
            if (p.selfcoef) Vec<4>::fma(o, sc, __ldg(reinterpret_cast<const float4*>(p.hself + v * p.ld_hself + c)));
            *reinterpret_cast<float4*>(p.out + v * p.ld_out + c) = o;
        }
    }
    if (S::TAIL1 && S::on1(lane)) {
        const int c = S::J4 * 128 + lane;
        if (p.raw) p.raw[v * p.ld_raw + c] = a.v1;
        if (p.dot) dot += a.v1 * __ldg(p.q + v * p.ld_q + c);
        if (p.out) {
            float o = scale * a.v1;
            if (p.selfcoef) o = fmaf(sc, __ldg(p.hself + v * p.ld_hself + c), o);
            p.out[v * p.ld_out + c] = o;
        }
    }
    if (p.dot) {
        dot = warp_sum(dot);
        if (lane == 0) p.dot[v] = dot;
    }
}

template <typename ColT, int DIM, int NW, int R, int STG, bool ESM, int CL = 1, int DIAG = 0>
// one CTA per SM; ptxas derives the register cap from the warp count (16K registers per SM sub-partition,
// so 13 warps -> 128 registers/thread, 12 warps -> 168).
// CL = 2: launched as clusters of two CTAs (neighbouring tiles of the same split).  Each producer copies every
// other row of the window and multicasts it into both CTAs' rings, so L2 and the crossbar deliver every source
// row once per PAIR of tiles; a stage is refilled only after the consumers of BOTH CTAs have released it.
__global__ void __launch_bounds__((NW + 1) * 32, 1)
agg_tiled_kernel(const TiledParams p) {
    using S = RowShape<DIM>;
    const uint32_t cta_rank = CL > 1 ? (blockIdx.x % CL) : 0;
    constexpr int kTiledStages = STG;       // depth of the window ring
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t full_bar[STG];
    __shared__ uint64_t empty_bar[STG];

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int tile = blockIdx.x % p.n_tiles;          // split-major launch order: CTAs that are resident
    const int split = blockIdx.x / p.n_tiles;         // together stream the same slice of hs
    const int w_begin = split * p.win_per_split;
    const int w_end = min(p.n_windows, w_begin + p.win_per_split);
    const int dim = DIM > 0 ? DIM : p.dim;
    const int pitch = DIM > 0 ? ((DIM + 31) & ~31) : p.pitch;
    const size_t stage_floats = (size_t)p.win_rows * pitch;
    float* stages = reinterpret_cast<float*>(smem_raw);
    // ESM: the current 32-edge chunk of every row is also kept in shared memory so that an edge's
    // (column, value) pair is fetched with ONE broadcast LDS.64 instead of two SHFLs (both go
    // through the same LSU data pipe that bounds this kernel)
    int2* edge_buf = reinterpret_cast<int2*>(stages + (size_t)STG * stage_floats) + (size_t)(warp < NW ? warp : 0) * R * 32;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < kTiledStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], NW * CL);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (CL > 1) cluster_sync_all(); else __syncthreads();       // peers must not touch uninitialised barriers

    if (warp == NW) {
        // ------------------------------- producer -------------------------------------------
        // one bulk copy per source row (lane l copies rows l, l+32, ...) into the padded stage
        for (int w = w_begin, it = 0; w < w_end; ++w, ++it) {
            const int s = it % kTiledStages;
            const uint32_t ph = (it / kTiledStages) & 1;
            mbar_wait(&empty_bar[s], ph ^ 1);
            const int64_t row0 = (int64_t)w * p.win_rows;
            const int rows = (int)min((int64_t)p.win_rows, p.n_src - row0);
            const uint32_t row_bytes = (uint32_t)(dim * sizeof(float));
            if (p.l2_prefetch > 0 && (tile & 7) == 0 && lane == 0 && w + p.l2_prefetch < w_end) {
                // tables larger than L2 (cell rows): one CTA in eight pulls a window further ahead into L2,
                // so the ring's own copies (all resident CTAs stream the same slice) find it there
                const int64_t prow0 = (int64_t)(w + p.l2_prefetch) * p.win_rows;
                const int prows = (int)min((int64_t)p.win_rows, p.n_src - prow0);
                bulk_prefetch_l2(p.hs + prow0 * dim, prows * row_bytes);
            }
            if (DIAG == 1) {
                if (lane == 0) mbar_arrive(&full_bar[s]);
                continue;
            }
            if (lane == 0) mbar_arrive_expect_tx(&full_bar[s], rows * row_bytes);     // bytes of BOTH producers land here
            __syncwarp();
            if (CL > 1) {
                for (int r = lane * CL + (int)cta_rank; r < rows; r += 32 * CL)
                    bulk_g2s_multicast(stages + s * stage_floats + (size_t)r * pitch, p.hs + (row0 + r) * dim, row_bytes,
                                       &full_bar[s], (uint16_t)((1u << CL) - 1));
            } else {
                for (int r = lane; r < rows; r += 32)
                    bulk_g2s(stages + s * stage_floats + (size_t)r * pitch, p.hs + (row0 + r) * dim, row_bytes, &full_bar[s]);
            }
        }
        if (CL > 1) cluster_sync_all();      // stay resident while the peer may still write into this CTA
        return;
    }

    // ----------------------------------- consumers ------------------------------------------
    const ColT* __restrict__ col = static_cast<const ColT*>(p.col);
    int64_t row[R];      // destination row (or -1)
    int64_t beg[R];      // first edge of the row
    int len[R];          // edges in the row
    int cur[R];          // next unconsumed edge (relative to beg)
    int ccol[R], ncol[R];   // this lane's column in the current / prefetched 32-edge chunk
    float cx[R], nx[R];     // this lane's value   in the current / prefetched chunk
    RowAcc<DIM> acc[R];

    // lane-private fetch of edge (chunk_base + lane); sentinel past the row end.  (ncu attributes 11 % of the
    // stall samples to the select that follows these loads; a clamped-index variant without the select removed
    // the stall but not a microsecond of run time — the kernel is throughput-bound on the shared-memory pipe —
    // and its extra ballot predicate cost 5 %, so the sentinel form stays.)
    auto fetch = [&](int r, int chunk_base, int& c_out, float& x_out) {
        const int e = chunk_base + lane;
        if (e < len[r]) {
            c_out = (int)col[beg[r] + e];
            x_out = p.x[beg[r] + e];
        } else {
            c_out = 0x7fffffff;
            x_out = 0.f;
        }
    };

#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int64_t i = ((int64_t)tile * NW + warp) * R + r;      // consecutive rows of row_perm
        row[r] = -1; beg[r] = 0; len[r] = 0; cur[r] = 0;
        if (i < p.n_dst) {
            row[r] = p.row_perm ? (int64_t)p.row_perm[i] : i;
            beg[r] = p.rowptr[row[r]];
            len[r] = (int)(p.rowptr[row[r] + 1] - beg[r]);
        }
        acc[r].zero();
        if (p.init != nullptr && split == 0 && row[r] >= 0) {
            const int64_t slot = p.init_map ? (int64_t)__ldg(p.init_map + row[r]) : row[r];
            if (slot >= 0) {
                for (int k = 0; k < p.init_slabs; ++k) {
                    const float* src = p.init + ((size_t)k * p.init_rows + slot) * dim;
#pragma unroll
                    for (int j = 0; j < S::N4; ++j) {
                        if (!S::on4(j, lane, dim)) continue;
                        const float4 t = __ldg(reinterpret_cast<const float4*>(src + (j * 32 + lane) * 4));
                        acc[r].v4[j].x += t.x; acc[r].v4[j].y += t.y; acc[r].v4[j].z += t.z; acc[r].v4[j].w += t.w;
                    }
                    if (S::TAIL1 && S::on1(lane)) acc[r].v1 += __ldg(src + S::J4 * 128 + lane);
                }
            }
        }
    }
    if (w_begin > 0) {
        // first edge with col >= first column of this split: lane r searches row r, then broadcast
        const int64_t first_col = (int64_t)w_begin * p.win_rows;
        int found = 0;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (lane == r) {
                int lo = 0, hi = len[r];
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if ((int64_t)col[beg[r] + mid] < first_col) lo = mid + 1; else hi = mid;
                }
                found = lo;
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) cur[r] = __shfl_sync(0xffffffffu, found, r);
    }
    auto publish = [&](int r) {       // current chunk of row r -> shared memory (warp-private region)
        if (ESM) {
            __syncwarp();
            edge_buf[r * 32 + lane] = make_int2(ccol[r], __float_as_int(cx[r]));
            __syncwarp();
        }
    };
    auto edge = [&](int r, int idx, int& c_out, float& x_out) {     // (column, value) of edge idx of the chunk
        if (ESM) {
            const int2 e = edge_buf[r * 32 + idx];
            c_out = e.x;
            x_out = __int_as_float(e.y);
        } else {
            c_out = __shfl_sync(0xffffffffu, ccol[r], idx);
            x_out = __shfl_sync(0xffffffffu, cx[r], idx);
        }
    };
#pragma unroll
    for (int r = 0; r < R; ++r) {
        fetch(r, cur[r] & ~31, ccol[r], cx[r]);
        fetch(r, (cur[r] & ~31) + 32, ncol[r], nx[r]);
        publish(r);
    }

    // Per-window bookkeeping is kept off the critical path after the barrier: stage index, phase, window bounds and
    // the stage pointer are running values (the pointer of window w in slot s is stages + (s - w)·W·pitch, which
    // only changes when the ring wraps), barrier addresses are precomputed shared-window offsets.
    const uint32_t full_a0 = smem_u32(&full_bar[0]), empty_a0 = smem_u32(&empty_bar[0]);
    // scalar tail chunk (columns J4*128 + lane): when the padded row has room for 32 lanes (400 -> pitch 416) every
    // lane loads, lanes past the row read the never-written padding into an accumulator that is never stored —
    // same single wavefront, no predicate / zero-fill instructions in the hot loop
    constexpr bool kTailAllLanes = DIM > 0 && S::TAIL1 && (((DIM + 31) & ~31) - S::J4 * 128 >= 32);
    int s = 0;
    uint32_t ph = 0;
    const int win_rows = p.win_rows;
    int win_end = (w_begin + 1) * win_rows;          // advanced at the END of an iteration: ready before the next wait
    const float* stage = stages + lane * 4 - (size_t)w_begin * win_rows * pitch;
    for (int w = w_begin; w < w_end; ++w) {
        mbar_wait_a(full_a0 + 8u * s, ph);
        if (DIAG == 2) {
            __syncwarp();
            if (lane == 0) {
                mbar_arrive_a(empty_a0 + 8u * s);
                if (CL > 1) mbar_arrive_remote(&empty_bar[s], cta_rank ^ 1);
            }
            win_end += win_rows;
            if (++s == kTiledStages) { s = 0; ph ^= 1; stage -= (size_t)kTiledStages * stage_floats; }
            continue;
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            while (true) {
                const int l0 = cur[r] & 31;
                const unsigned m = __ballot_sync(0xffffffffu, lane >= l0 && ccol[r] < win_end);
                const int cnt = __popc(m);         // columns ascend: a contiguous run starting at lane l0
                int k = l0;                            // index of the next edge inside the chunk
                const int2* eb = edge_buf + r * 32 + l0;      // ESM: running pointer into the row's edge mirror
                for (int pairs = cnt >> 1; pairs > 0; --pairs) {      // two edges per iteration: 2x the loads in flight
                    int c0, c1;
                    float x0, x1;
                    if (ESM) {
                        const int2 e0 = eb[0], e1 = eb[1];
                        eb += 2;
                        c0 = e0.x; x0 = __int_as_float(e0.y);
                        c1 = e1.x; x1 = __int_as_float(e1.y);
                    } else {
                        edge(r, k, c0, x0);
                        edge(r, k + 1, c1, x1);
                        k += 2;
                    }
                    const float* s0 = stage + (size_t)c0 * pitch;
                    const float* s1 = stage + (size_t)c1 * pitch;
                    float4 a4[S::N4], b4[S::N4];
                    float a1 = 0.f, b1 = 0.f;
#pragma unroll
                    for (int j = 0; j < S::N4; ++j) {
                        if (S::on4(j, lane, dim)) {
                            a4[j] = *reinterpret_cast<const float4*>(s0 + j * 128);
                            b4[j] = *reinterpret_cast<const float4*>(s1 + j * 128);
                        }
                    }
                    if (S::TAIL1 && (kTailAllLanes || S::on1(lane))) {
                        a1 = s0[S::J4 * 128 - lane * 3];           // column J4*128 + lane (s0 already has +4*lane)
                        b1 = s1[S::J4 * 128 - lane * 3];
                    }
#pragma unroll
                    for (int j = 0; j < S::N4; ++j) {
                        if (S::on4(j, lane, dim)) {
                            Vec<4>::fma(acc[r].v4[j], x0, a4[j]);
                            Vec<4>::fma(acc[r].v4[j], x1, b4[j]);
                        }
                    }
                    if (S::TAIL1) acc[r].v1 = fmaf(x1, b1, fmaf(x0, a1, acc[r].v1));
                }
                if (cnt & 1) {                           // odd edge left over
                    int c0;
                    float x0;
                    if (ESM) {
                        const int2 e0 = eb[0];
                        c0 = e0.x; x0 = __int_as_float(e0.y);
                    } else {
                        edge(r, k, c0, x0);
                    }
                    const float* s0 = stage + (size_t)c0 * pitch;
                    float4 a4[S::N4];
                    float a1 = 0.f;
#pragma unroll
                    for (int j = 0; j < S::N4; ++j)
                        if (S::on4(j, lane, dim)) a4[j] = *reinterpret_cast<const float4*>(s0 + j * 128);
                    if (S::TAIL1 && (kTailAllLanes || S::on1(lane))) a1 = s0[S::J4 * 128 - lane * 3];
#pragma unroll
                    for (int j = 0; j < S::N4; ++j)
                        if (S::on4(j, lane, dim)) Vec<4>::fma(acc[r].v4[j], x0, a4[j]);
                    if (S::TAIL1) acc[r].v1 = fmaf(x0, a1, acc[r].v1);
                }
                cur[r] += cnt;
                if (cnt == 0 || (cur[r] & 31) != 0) break;
                // chunk used up: the prefetched one becomes current (all-sentinel past the row
                // end) and the one after it is requested; the window may continue in it
                ccol[r] = ncol[r];
                cx[r] = nx[r];
                fetch(r, cur[r] + 32, ncol[r], nx[r]);
                publish(r);
            }
        }
        __syncwarp();
        if (lane == 0) {
            mbar_arrive_a(empty_a0 + 8u * s);
            if (CL > 1) mbar_arrive_remote(&empty_bar[s], cta_rank ^ 1);
        }
        win_end += win_rows;
        if (++s == kTiledStages) { s = 0; ph ^= 1; stage -= (size_t)kTiledStages * stage_floats; }
    }

    if (CL > 1) cluster_sync_all();      // the peer may still multicast into / arrive on this CTA until its last window

    // ------------------------------------ epilogue -------------------------------------------
#pragma unroll
    for (int r = 0; r < R; ++r) {
        if (row[r] < 0) continue;
        if (p.n_splits == 1) {
            tiled_row_epilogue<DIM>(p, row[r], acc[r], lane);
        } else {
            float* dst = p.partial + ((size_t)split * p.n_dst + row[r]) * dim;
#pragma unroll
            for (int j = 0; j < S::N4; ++j)
                if (S::on4(j, lane, dim)) *reinterpret_cast<float4*>(dst + (j * 32 + lane) * 4) = acc[r].v4[j];
            if (S::TAIL1 && S::on1(lane)) dst[S::J4 * 128 + lane] = acc[r].v1;
        }
    }
}

// Sums the split partials in fixed order and applies the epilogue: one warp per row.
__global__ void __launch_bounds__(256)
tiled_reduce_kernel(const TiledParams p) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int64_t nwarps = (int64_t)gridDim.x * 8;
    for (int64_t v = warp0; v < p.n_dst; v += nwarps) {
        RowAcc<0> acc;
        acc.zero();
        for (int s = 0; s < p.n_splits; ++s) {
            const float* src = p.partial + ((size_t)s * p.n_dst + v) * p.dim;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = (j * 32 + lane) * 4;
                if (c < p.dim) {
                    const float4 t = *reinterpret_cast<const float4*>(src + c);
                    acc.v4[j].x += t.x; acc.v4[j].y += t.y; acc.v4[j].z += t.z; acc.v4[j].w += t.w;
                }
            }
        }
        tiled_row_epilogue<0>(p, v, acc, lane);
    }
}

// ------------------------------------------ host side ------------------------------------------
// Kernel shape: consumer warps per CTA, destination rows per warp, depth of the window ring.
//   widths 400 / 200 / 64 (BASELINE's configs and the reference's hidden_dim): 15 consumer warps + producer = 16 warps at
//     128 registers fill the register file exactly, 60 rows per tile (c4 cell<-gene 94.7 -> 87.1 ms against 12 warps,
//     profiles/r01_summary.md);
//   every other width (128, run-time width): 11 consumer warps + producer = 12 warps at 168 registers — the 16-warp shape
//     (128 registers) spills there.
// -DWSAGE_TUNING additionally compiles the shapes that lost the round-1 sweeps (ring depth 2 / 4, 12 / 14 / 16 warps, no
// edge mirror, 2-CTA clusters, the fill-only / walk-only diagnostics) and lets WSAGE_TILED_VARIANT / _DIAG / _CLUSTER pick them
// for dim == 400.
struct TiledVariant { int nw, r, stages; bool esm; };
constexpr TiledVariant kTiledWide = {15, 4, 3, true};
constexpr TiledVariant kTiledSafe = {11, 4, 3, true};      // 11 + producer = 12 warps: 3 per scheduler, 168 registers
inline bool tiled_wide_dim(int dim) { return dim == 400 || dim == 200 || dim == 64; }

#ifdef WSAGE_TUNING
inline int tiled_diag_env() { const char* e = getenv("WSAGE_TILED_DIAG"); const int i = e ? atoi(e) : 0; return (i == 1 || i == 2) ? i : 0; }
inline int tiled_cluster_env() { const char* e = getenv("WSAGE_TILED_CLUSTER"); return (e && atoi(e) == 2) ? 2 : 1; }
constexpr int kTiledDefault400 = 6;
constexpr TiledVariant kTiledVariants[] = {{12, 4, 3, true}, {12, 4, 3, false}, {12, 4, 4, true}, {16, 3, 4, true}, {12, 4, 2, true}, {14, 4, 3, true}, {15, 4, 3, true}, {15, 4, 4, true}, {15, 4, 2, true}};
constexpr int kNumTiledVariants = sizeof(kTiledVariants) / sizeof(kTiledVariants[0]);
inline int tiled_variant_index() {
    static const int v = [] {
        const char* e = getenv("WSAGE_TILED_VARIANT");
        int i = e ? atoi(e) : kTiledDefault400;
        if (tiled_diag_env() != 0 || tiled_cluster_env() == 2) i = e ? i : 0;     // those variants exist for shape [0] only
        return (i >= 0 && i < kNumTiledVariants) ? i : kTiledDefault400;
    }();
    return v;
}
#endif

inline TiledVariant tiled_variant(const wsage_spmm_args* a) {
#ifdef WSAGE_TUNING
    if (a->dim == 400) return kTiledVariants[tiled_variant_index()];
#endif
    return tiled_wide_dim(a->dim) ? kTiledWide : kTiledSafe;
}

struct TiledPlan {
    TiledVariant v;
    int win_rows, pitch, n_windows, n_tiles, n_splits, win_per_split;
    size_t smem_bytes, workspace_bytes;
};

inline bool tiled_supported(const wsage_spmm_args* a, bool vec4) {
    return vec4 && a->dim <= 512 && a->ld_hs == a->dim && a->n_src < (int64_t)0x7fffffff &&
           a->n_dst < (int64_t)0x7fffffff && (int64_t)((a->dim + 31) & ~31) * 4 * 8 <= kTiledSmemBudget / kTiledStages;
}

inline TiledPlan tiled_plan(const wsage_spmm_args* a) {
    TiledPlan pl{};
    pl.v = tiled_variant(a);
    const int rows_per_tile = pl.v.nw * pl.v.r;
    const size_t row_bytes = (size_t)a->dim * sizeof(float);
    pl.pitch = (a->dim + 31) & ~31;
    const size_t pitch_bytes = (size_t)pl.pitch * sizeof(float);
    const size_t edge_bytes = pl.v.esm ? (size_t)pl.v.nw * pl.v.r * 32 * sizeof(int2) : 0;
    int w = (int)((kTiledSmemBudget - edge_bytes) / pl.v.stages / pitch_bytes);
    if ((int64_t)w > a->n_src) w = (int)(a->n_src > 0 ? a->n_src : 1);
    pl.win_rows = w;
    pl.n_windows = (int)((a->n_src + w - 1) / w);
    pl.n_tiles = (int)((a->n_dst + rows_per_tile - 1) / rows_per_tile);
    // (a) each split's slice of hs should stay L2-resident while the resident CTAs stream it
    const double table_bytes = (double)a->n_src * row_bytes;
    int splits = (int)((table_bytes + 40.0 * (1 << 20) - 1) / (40.0 * (1 << 20)));
    // (b) few tiles means few, long, degree-skewed rows (gene destinations): cut them into >= ~48
    //     waves of work units so the heaviest tile cannot dominate; many tiles balance by themselves
    if (pl.n_tiles < 8 * kNumSMs) {
        const int balance = (48 * kNumSMs + pl.n_tiles - 1) / pl.n_tiles;
        if (balance > splits) splits = balance;
    }
    const int max_splits = pl.n_windows / 8 > 0 ? pl.n_windows / 8 : 1;
    {   // WSAGE_TILED_SPLITS (env, tuning only) overrides the heuristic
        static const int forced = [] { const char* e = getenv("WSAGE_TILED_SPLITS"); return e ? atoi(e) : 0; }();
        if (forced > 0) splits = forced;
    }
    if (splits > max_splits) splits = max_splits;
    if (splits > 256) splits = 256;
    if (splits < 1) splits = 1;
    pl.win_per_split = (pl.n_windows + splits - 1) / splits;
    pl.n_splits = (pl.n_windows + pl.win_per_split - 1) / pl.win_per_split;
    pl.smem_bytes = (size_t)pl.v.stages * w * pitch_bytes + (pl.v.esm ? (size_t)pl.v.nw * pl.v.r * 32 * sizeof(int2) : 0);
    pl.workspace_bytes = pl.n_splits > 1 ? (size_t)pl.n_splits * a->n_dst * a->dim * sizeof(float) : 0;
    return pl;
}

inline size_t tiled_workspace_bytes(const wsage_spmm_args* a, bool vec4) {
    if (a->algo == 1 || !tiled_supported(a, vec4)) return 0;
    return tiled_plan(a).workspace_bytes;
}

// Worth it when a source row is referenced by >= ~1.5 rows of a tile on average (then staging it
// once per tile moves fewer bytes out of L2 than gathering it once per edge) and there is enough
// work to amortise the pipeline; otherwise the L2 gather kernel wins.
inline bool tiled_profitable(const wsage_spmm_args* a, bool vec4) {
    if (!tiled_supported(a, vec4) || a->n_src == 0 || a->n_dst == 0) return false;
    const TiledVariant v = tiled_variant(a);
    const double avg_deg = (double)a->nnz / (double)a->n_dst;
    const double reuse = avg_deg * v.nw * v.r / (double)a->n_src;
    return reuse >= 1.5 && a->nnz >= (int64_t)1 << 20;
}

struct TiledInit { const float* init; int slabs; int64_t rows; const int32_t* map; };

#ifdef WSAGE_TUNING
// Profiling aid (WSAGE_TILED_DIAG, dim 400 shape [0] only; compile-time variants so that the production kernel carries no
// extra branch — a run-time flag in the window loop cost 3.7 % of the step): 1 = the producer copies nothing (edge-walk
// time only, results meaningless), 2 = the consumers skip the edge walk (window-fill time only).
inline int tiled_diag() {
    static const int v = tiled_diag_env();
    return v;
}
inline int tiled_cluster() {      // WSAGE_TILED_CLUSTER = 1 | 2 (dim 400, shape [0] only)
    static const int v = kTiledDefaultCluster == 2 ? 2 : tiled_cluster_env();
    return v;
}
#endif

template <typename ColT, int DIM, int NW, int R, int STG, bool ESM, int CL = 1, int DIAG = 0>
int launch_tiled_shape(const wsage_spmm_args* a, const TiledPlan& pl, const TiledInit& ini, cudaStream_t st) {
    TiledParams p{};
    p.init = ini.init; p.init_slabs = ini.slabs; p.init_rows = ini.rows; p.init_map = ini.map;
    p.rowptr = a->rowptr; p.col = a->col; p.x = a->x; p.hs = a->hs;
    p.n_src = a->n_src; p.n_dst = a->n_dst; p.dim = a->dim;
    p.win_rows = pl.win_rows; p.pitch = pl.pitch; p.n_windows = pl.n_windows; p.n_tiles = pl.n_tiles;
    p.n_splits = pl.n_splits; p.win_per_split = pl.win_per_split; p.row_perm = a->row_perm;
    {   // WSAGE_TILED_L2PF (env, tuning only): prefetch distance in windows
        static const int forced = [] { const char* e = getenv("WSAGE_TILED_L2PF"); return e ? atoi(e) : -1; }();
        const bool big = (double)a->n_src * a->dim * sizeof(float) > 64.0 * (1 << 20);
        p.l2_prefetch = forced >= 0 ? forced : 0;      // measured: no gain at c3/c4 (the slice is L2-resident), off
        (void)big;
    }
    p.dscale = a->dscale; p.selfcoef = a->selfcoef; p.hself = a->hself; p.ld_hself = a->ld_hself;
    p.out = a->out; p.ld_out = a->ld_out; p.raw = a->raw; p.ld_raw = a->ld_raw;
    p.q = a->q; p.ld_q = a->ld_q; p.dot = a->dot;
    p.partial = static_cast<float*>(a->workspace);
    auto kern = agg_tiled_kernel<ColT, DIM, NW, R, STG, ESM, CL, DIAG>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bytes);
    if (e != cudaSuccess) return fail(WSAGE_ECUDA, "%s: %s", "cudaFuncSetAttribute(agg_tiled)", cudaGetErrorString(e));
    if (CL > 1) {
        // clusters pair neighbouring tiles of one split: pad the tile count to a multiple of CL (the extra tile has
        // no rows but takes part in the ring protocol)
        p.n_tiles = (pl.n_tiles + CL - 1) / CL * CL;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(p.n_tiles * pl.n_splits));
        cfg.blockDim = dim3((NW + 1) * 32);
        cfg.dynamicSmemBytes = pl.smem_bytes;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        e = cudaLaunchKernelEx(&cfg, kern, p);
        if (e != cudaSuccess) return fail(WSAGE_ECUDA, "%s: %s", "cudaLaunchKernelEx(agg_tiled cluster)", cudaGetErrorString(e));
    } else {
        kern<<<pl.n_tiles * pl.n_splits, (NW + 1) * 32, pl.smem_bytes, st>>>(p);
    }
    int rc = check_launch("agg_tiled");
    if (rc != WSAGE_OK || pl.n_splits == 1) return rc;
    tiled_reduce_kernel<<<gather_grid(a->n_dst), 256, 0, st>>>(p);
    return check_launch("tiled_reduce");
}

template <typename ColT>
int launch_tiled_col(const wsage_spmm_args* a, const TiledPlan& pl, const TiledInit& ini, cudaStream_t st) {
    switch (a->dim) {       // widths of the reference (dense_dim 400, hidden 200) and of BASELINE's configs (400 / 400, 64)
        case 400:
#ifdef WSAGE_TUNING
            switch (tiled_variant_index()) {
                case 1: return launch_tiled_shape<ColT, 400, 12, 4, 3, false>(a, pl, ini, st);
                case 2: return launch_tiled_shape<ColT, 400, 12, 4, 4, true>(a, pl, ini, st);
                case 3: return launch_tiled_shape<ColT, 400, 16, 3, 4, true>(a, pl, ini, st);
                case 4: return launch_tiled_shape<ColT, 400, 12, 4, 2, true>(a, pl, ini, st);
                case 5: return launch_tiled_shape<ColT, 400, 14, 4, 3, true>(a, pl, ini, st);
                case 6: break;
                case 7: return launch_tiled_shape<ColT, 400, 15, 4, 4, true>(a, pl, ini, st);
                case 8: return launch_tiled_shape<ColT, 400, 15, 4, 2, true>(a, pl, ini, st);
                default:
                    if (tiled_diag() == 1) return launch_tiled_shape<ColT, 400, 12, 4, 3, true, 1, 1>(a, pl, ini, st);
                    if (tiled_diag() == 2) return launch_tiled_shape<ColT, 400, 12, 4, 3, true, 1, 2>(a, pl, ini, st);
                    if (tiled_cluster() == 2) return launch_tiled_shape<ColT, 400, 12, 4, 3, true, 2>(a, pl, ini, st);
                    return launch_tiled_shape<ColT, 400, 12, 4, 3, true>(a, pl, ini, st);
            }
#endif
            return launch_tiled_shape<ColT, 400, 15, 4, 3, true>(a, pl, ini, st);
        case 200: return launch_tiled_shape<ColT, 200, 15, 4, 3, true>(a, pl, ini, st);
        case 64:  return launch_tiled_shape<ColT, 64, 15, 4, 3, true>(a, pl, ini, st);
        case 128: return launch_tiled_shape<ColT, 128, 11, 4, 3, true>(a, pl, ini, st);
        default:  return launch_tiled_shape<ColT, 0, 11, 4, 3, true>(a, pl, ini, st);
    }
}

// The split partials occupy the first tiled_plan(a).workspace_bytes of the workspace (the dense block's
// sums, if any, follow; wsage_spmm checks the size).
inline int launch_tiled(const wsage_spmm_args* a, const TiledInit& ini, cudaStream_t st) {
    const TiledPlan pl = tiled_plan(a);
    return a->col_bits == WSAGE_COL_U16 ? launch_tiled_col<uint16_t>(a, pl, ini, st) : launch_tiled_col<int32_t>(a, pl, ini, st);
}

}  // namespace wsage
