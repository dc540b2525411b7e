// Post-aggregate dense layer on the 5th-generation tensor cores (sm_100a):
//
//   out[M, N] = act( A[M, K] * B[N, K]^T + bias[N] )         fp32 in / fp32 out
//
// replaces NodeUpdate.forward (/root/reference/models/gnn.py:18-25: fc_neigh + activation) and its
// input-gradient GEMM.  fp32 parity (1e-4) rules out plain bf16/tf32, so every operand is split into
// tf32 hi + tf32 lo (split_tf32_kernel: hi = rn_tf32(x), lo = rn_tf32(x - hi), both stored as fp32
// words whose low 13 bits are zero, so the tensor core reads them exactly) and three MMAs are
// accumulated per k-step in TMEM:  hi*hi + lo*hi + hi*lo   (residual ~2^-21 relative per product,
// i.e. fp32-grade; a bf16 split would leave ~2^-17).
//
// One CTA per 128-row tile (persistent, static stride), all N <= 512 columns in TMEM:
//   warp 0      TMA producer: cp.async.bulk.tensor.2d (SASS UTMALDG) of A_hi/A_lo [128 x 16] and
//               B_hi/B_lo [N x 16] tf32 tiles, 64-byte swizzle, 3-stage mbarrier ring
//   warp 1      TMEM allocator + single-thread tcgen05.mma.kind::tf32 issuer (SASS UTCHMMA), commits to the ring
//   warps 2..5  epilogue: tcgen05.ld (LDTM) 32 lanes x 32 columns -> +bias, ReLU -> shared-memory
//               transpose -> coalesced 128-byte global stores
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace wsage {

constexpr int kTcBlockM = 128;
constexpr int kTcBlockK = 16;          // tf32 elements per k-block = 64 bytes = one 64B-swizzle row
constexpr int kTcUmmaK = 8;            // K of one tcgen05.mma.kind::tf32 (32 bytes)
constexpr int kTcStages = 3;
constexpr int kTcThreads = 192;        // 6 warps
constexpr int kTcMaxN = 512;           // TMEM columns

// ------------------------------------------------------------------------------------------------
// fp32 -> (tf32 hi, tf32 lo) split, optionally fused with the ReLU-backward mask (g = x * (y > 0))
// and an fp32 copy of the masked value (for the weight-gradient GEMM).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float rn_tf32(float v) {       // round to nearest tf32, returned as an fp32 word
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}

__global__ void __launch_bounds__(256)
split_tf32_kernel(const float* __restrict__ x, int64_t ld_x, const float* __restrict__ mask_src, int64_t ld_m,
                  float* __restrict__ hi, float* __restrict__ lo, int64_t ld_o,
                  float* __restrict__ masked, int64_t ld_mk, int64_t rows, int cols) {
    const int64_t n4 = cols / 4;
    const int64_t total = rows * n4;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / n4;
        const int c = (int)(i - r * n4) * 4;
        float4 v = *reinterpret_cast<const float4*>(x + r * ld_x + c);
        if (mask_src) {
            const float4 y = *reinterpret_cast<const float4*>(mask_src + r * ld_m + c);
            v.x = y.x > 0.f ? v.x : 0.f; v.y = y.y > 0.f ? v.y : 0.f;
            v.z = y.z > 0.f ? v.z : 0.f; v.w = y.w > 0.f ? v.w : 0.f;
            if (masked) *reinterpret_cast<float4*>(masked + r * ld_mk + c) = v;
        }
        float4 h, l;
        h.x = rn_tf32(v.x); h.y = rn_tf32(v.y); h.z = rn_tf32(v.z); h.w = rn_tf32(v.w);
        l.x = rn_tf32(v.x - h.x); l.y = rn_tf32(v.y - h.y); l.z = rn_tf32(v.z - h.z); l.w = rn_tf32(v.w - h.w);
        *reinterpret_cast<float4*>(hi + r * ld_o + c) = h;
        if (lo) *reinterpret_cast<float4*>(lo + r * ld_o + c) = l;
    }
}

// ------------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int x, int y, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {      // arrives on bar when all prior MMAs of this thread retire
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, tf32 inputs, fp32 accumulate; issued by ONE thread for the CTA.
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// K-major operand tile, 64-byte swizzle: rows of 64 B, 8-row groups 512 B apart (SBO), LBO unused (=1).
__device__ __forceinline__ uint64_t umma_desc_sw64(const void* smem) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_u32(smem) & 0x3FFFF) >> 4);          // start address      bits [0,14)
    d |= (uint64_t)1 << 16;                                    // leading byte offset bits [16,30) (ignored when swizzled)
    d |= (uint64_t)(512 >> 4) << 32;                           // stride byte offset  bits [32,46)
    d |= (uint64_t)1 << 46;                                    // descriptor version 1 (Blackwell)
    d |= (uint64_t)4 << 61;                                    // layout type SWIZZLE_64B
    return d;
}
// kind::tf32 instruction descriptor: D fp32 (1), A/B tf32 (2), both K-major, M = 128, N = n.
__device__ __forceinline__ uint32_t umma_idesc_tf32(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTcBlockM >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

struct LinearTcParams {
    int terms;               // 3: hi*hi + lo*hi + hi*lo (fp32-grade); 1: hi*hi only (plain tf32, the bf16 configuration)
    int64_t m;
    int n, n_pad, k;         // n_pad = N rounded up to 16 (MMA granularity); columns >= N are never stored
    int n1, n2;              // MMA N halves: n1 = min(N, 256), n2 = N - n1 (multiples of 16)
    int b_boxes, b_box_rows; // TMA boxes covering the N rows of B (a box is at most 256 rows)
    const float* bias;       // [N] or null
    int relu;
    float* out;
    int64_t ld_out;
    int num_tiles;
};

struct LinearTcSmem {       // carved from dynamic shared memory (1024-byte aligned)
    static constexpr int a_bytes = kTcBlockM * kTcBlockK * 4;            // 8 KB per hi / lo tile
    static __host__ __device__ int b_bytes(int n) { return n * kTcBlockK * 4; }
    static __host__ __device__ int stage_bytes(int n) { return 2 * a_bytes + 2 * b_bytes(n); }
    static __host__ __device__ size_t total(int n) {
        return (size_t)kTcStages * stage_bytes(n) + 4 * 32 * 33 * sizeof(float) + 1024 /*align slack*/ + 256 /*barriers*/;
    }
};

__global__ void __launch_bounds__(kTcThreads, 1)
linear_tc_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                 const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                 const LinearTcParams p) {
    extern __shared__ unsigned char dsmem_raw[];
    unsigned char* dsmem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(dsmem_raw) + 1023) & ~uintptr_t(1023));
    const int stage_bytes = LinearTcSmem::stage_bytes(p.n_pad);
    const int b_bytes = LinearTcSmem::b_bytes(p.n_pad);
    unsigned char* ring = dsmem;
    float* xpose = reinterpret_cast<float*>(ring + (size_t)kTcStages * stage_bytes);     // [4 warps][32][33]
    uint64_t* bars = reinterpret_cast<uint64_t*>(xpose + 4 * 32 * 33);
    uint64_t* full_bar = bars;                       // [kTcStages]  TMA -> MMA
    uint64_t* empty_bar = bars + kTcStages;          // [kTcStages]  MMA -> TMA
    uint64_t* tmem_full = bars + 2 * kTcStages;      //              MMA -> epilogue
    uint64_t* tmem_empty = bars + 2 * kTcStages + 1; //              epilogue -> MMA
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kTcStages + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_kb = (p.k + kTcBlockK - 1) / kTcBlockK;       // K tail is zero-filled by TMA

    if (threadIdx.x == 0) {
        for (int s = 0; s < kTcStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(tmem_full, 1);
        mbar_init(tmem_empty, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {        // whole warp: allocate all 512 TMEM columns (1 CTA per SM)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(kTcMaxN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ------------------------------------ TMA producer ------------------------------------
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a_hi) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b_hi) : "memory");
            int it = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                const int m0 = tile * kTcBlockM;
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % kTcStages;
                    const uint32_t ph = (it / kTcStages) & 1;
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    unsigned char* st = ring + (size_t)s * stage_bytes;
                    mbar_arrive_expect_tx(&full_bar[s], (uint32_t)(p.terms == 3 ? stage_bytes : stage_bytes / 2));
                    const int k0 = kb * kTcBlockK;
                    tma_load_2d(st, &map_a_hi, k0, m0, &full_bar[s]);
                    if (p.terms == 3) tma_load_2d(st + LinearTcSmem::a_bytes, &map_a_lo, k0, m0, &full_bar[s]);
                    unsigned char* sb = st + 2 * LinearTcSmem::a_bytes;
                    for (int j = 0; j < p.b_boxes; ++j) {
                        const int r0 = j * p.b_box_rows;
                        tma_load_2d(sb + r0 * kTcBlockK * 4, &map_b_hi, k0, r0, &full_bar[s]);
                        if (p.terms == 3) tma_load_2d(sb + b_bytes + r0 * kTcBlockK * 4, &map_b_lo, k0, r0, &full_bar[s]);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------ MMA issuer --------------------------------------
        if (lane == 0) {
            const uint32_t idesc1 = umma_idesc_tf32(p.n1);
            const uint32_t idesc2 = umma_idesc_tf32(p.n2 > 0 ? p.n2 : 16);
            int it = 0, local_tile = 0;
            for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++local_tile) {
                mbar_wait(tmem_empty, (local_tile & 1) ^ 1);       // epilogue has drained the accumulators
                tc_fence_after();
                for (int kb = 0; kb < num_kb; ++kb, ++it) {
                    const int s = it % kTcStages;
                    const uint32_t ph = (it / kTcStages) & 1;
                    mbar_wait(&full_bar[s], ph);
                    tc_fence_after();
                    unsigned char* st = ring + (size_t)s * stage_bytes;
                    const uint64_t a_hi = umma_desc_sw64(st), a_lo = umma_desc_sw64(st + LinearTcSmem::a_bytes);
                    unsigned char* sb = st + 2 * LinearTcSmem::a_bytes;
                    const uint64_t b_hi = umma_desc_sw64(sb), b_lo = umma_desc_sw64(sb + b_bytes);
                    const uint64_t n2_off = (uint64_t)((p.n1 * kTcBlockK * 4) >> 4);
#pragma unroll
                    for (int kk = 0; kk < kTcBlockK / kTcUmmaK; ++kk) {      // UMMA_K = 8 tf32 = 32 bytes along K
                        const uint64_t ko = (uint64_t)(kk * 32 >> 4);
                        const uint32_t acc0 = (kb | kk) != 0;
                        // hi*hi, lo*hi, hi*lo  (smallest terms last does not matter: fp32 accumulate)
                        tc_mma_tf32(tmem_base, a_hi + ko, b_hi + ko, idesc1, acc0);
                        if (p.terms == 3) {
                            tc_mma_tf32(tmem_base, a_lo + ko, b_hi + ko, idesc1, 1);
                            tc_mma_tf32(tmem_base, a_hi + ko, b_lo + ko, idesc1, 1);
                        }
                        if (p.n2 > 0) {
                            tc_mma_tf32(tmem_base + p.n1, a_hi + ko, b_hi + n2_off + ko, idesc2, acc0);
                            if (p.terms == 3) {
                                tc_mma_tf32(tmem_base + p.n1, a_lo + ko, b_hi + n2_off + ko, idesc2, 1);
                                tc_mma_tf32(tmem_base + p.n1, a_hi + ko, b_lo + n2_off + ko, idesc2, 1);
                            }
                        }
                    }
                    tc_commit(&empty_bar[s]);          // stage reusable once these MMAs have read it
                }
                tc_commit(tmem_full);                  // accumulators complete
            }
        }
    } else {
        // ------------------------------------ epilogue ----------------------------------------
        const int q = warp & 3;                        // TMEM lane quarter this warp may access
        float* xp = xpose + (warp - 2) * 32 * 33;
        int local_tile = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++local_tile) {
            mbar_wait(tmem_full, local_tile & 1);
            tc_fence_after();
            const int64_t row0 = (int64_t)tile * kTcBlockM + q * 32;
            for (int c0 = 0; c0 < p.n; c0 += 32) {
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
                for (int j = 0; j < 32; ++j) xp[lane * 33 + j] = v[j];       // thread = row, j = column
                __syncwarp();
                const int c = c0 + lane;
                if (c < p.n) {
                    const float b = p.bias ? p.bias[c] : 0.f;
                    for (int r = 0; r < 32; ++r) {                            // lanes = 32 consecutive columns of row r
                        const int64_t row = row0 + r;
                        if (row >= p.m) break;
                        float o = xp[r * 33 + lane] + b;
                        if (p.relu) o = fmaxf(o, 0.f);
                        p.out[row * p.ld_out + c] = o;
                    }
                }
                __syncwarp();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTcMaxN));
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void* f = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            f = nullptr;
        return reinterpret_cast<EncodeTiledFn>(f);
    }();
    return fn;
}

// fp32 (tf32-valued) matrix [rows, cols] with row pitch ld (elements): box = [box_rows, 16 cols], 64-byte swizzle.
inline int make_tf32_map(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return fail(WSAGE_ECUDA, "%s: %s", "wsage_linear_tc", "cuTensorMapEncodeTiled not available");
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    const cuuint32_t box[2] = {(cuuint32_t)kTcBlockK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(WSAGE_ECUDA, "%s: %s", "wsage_linear_tc", "cuTensorMapEncodeTiled failed");
    return WSAGE_OK;
}


// ================================================================================================
// Weight gradient on the tensor cores:  dW[n_out, n_in] = g[rows, n_out]^T * x[rows, n_in]
//
// The reduction runs over the ROWS (all cells and genes, 7.8e5 at atlas scale), so both operands are
// "MN-major" for the MMA: in memory the M (resp. N) index is contiguous and K strides by a row.  For 32-bit
// operands the tensor core takes MN-major tiles only in the SWIZZLE_128B_BASE32B layout (CUTLASS sm100_common.inl:
// "for mn-major tf32 operands, SW128_32B is the only available smem layout"): 32 MN elements (128 B) contiguous,
// K atoms of 4 rows, 32-byte chunks of a row XOR-swizzled by (row mod 4) — which is what a TMA box of
// [16 k-rows x 32 floats] with CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B writes.  Descriptor: LBO = 2 KB (next
// 32-element MN block = next box), SBO = 512 B (next 4-row K atom); one tf32 MMA (K = 8) spans two atoms.
// Same tf32 hi/lo three-product scheme as linear_tc_kernel.  One CTA per
// (128-row tile of n_out, split of the row range); partial tiles go to the workspace and
// grad_w_reduce_kernel adds them in split order (deterministic).
// ================================================================================================
constexpr int kGwBlockK = 16;                     // rows (K) per pipeline stage = two 8-row K atoms
constexpr int kGwMnBlock = 32;                    // floats per swizzle row (128 B)
constexpr int kGwBoxBytes = kGwBlockK * kGwMnBlock * 4;     // 2 KB: one TMA box = one MN block of a stage
constexpr int kGwStages = 3;
constexpr int kGwABlocks = kTcBlockM / kGwMnBlock;          // 4 MN blocks cover the 128-row M tile

struct GradWParams {
    int terms;               // 3 | 1, as LinearTcParams
    int64_t rows;            // K extent
    int n_out, n_in;
    int n_pad;               // n_in rounded up to 16
    int n1, n2;              // MMA N halves
    int b_blocks;            // ceil(n_pad / 32) MN blocks of B per stage
    int num_kb, kb_per_split, n_splits;
    float* partial;          // [n_splits][n_out][n_in]
};

__device__ __forceinline__ uint64_t umma_desc_mn_sw128(const void* smem) {     // MN-major, 128B swizzle with 32B atoms
    uint64_t d = 0;
    d |= (uint64_t)((smem_u32(smem) & 0x3FFFF) >> 4);          // start address
    d |= (uint64_t)(kGwBoxBytes >> 4) << 16;                   // leading byte offset: next 32-element MN block
    d |= (uint64_t)(512 >> 4) << 32;                           // stride byte offset: next 4-row K atom
    d |= (uint64_t)1 << 46;                                    // descriptor version 1 (Blackwell)
    d |= (uint64_t)1 << 61;                                    // layout type SWIZZLE_128B_BASE32B
    return d;
}
// kind::tf32, D fp32, A and B tf32, both MN-major (bits 15 / 16), M = 128, N = n.
__device__ __forceinline__ uint32_t umma_idesc_tf32_mn(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTcBlockM >> 4) << 24);
}

struct GradWSmem {
    static __host__ __device__ int a_bytes() { return kGwABlocks * kGwBoxBytes; }                  // 8 KB per hi / lo
    static __host__ __device__ int b_bytes(int b_blocks) { return b_blocks * kGwBoxBytes; }
    static __host__ __device__ int stage_bytes(int b_blocks) { return 2 * a_bytes() + 2 * b_bytes(b_blocks); }
    static __host__ __device__ size_t total(int b_blocks) {
        return (size_t)kGwStages * stage_bytes(b_blocks) + 4 * 32 * 33 * sizeof(float) + 1024 + 256;
    }
};

__global__ void __launch_bounds__(kTcThreads, 1)
grad_w_tc_kernel(const __grid_constant__ CUtensorMap map_g_hi, const __grid_constant__ CUtensorMap map_g_lo,
                 const __grid_constant__ CUtensorMap map_x_hi, const __grid_constant__ CUtensorMap map_x_lo,
                 const GradWParams p) {
    extern __shared__ unsigned char dsmem_raw[];
    unsigned char* ring = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(dsmem_raw) + 1023) & ~uintptr_t(1023));
    const int stage_bytes = GradWSmem::stage_bytes(p.b_blocks);
    const int a_bytes = GradWSmem::a_bytes();
    const int b_bytes = GradWSmem::b_bytes(p.b_blocks);
    float* xpose = reinterpret_cast<float*>(ring + (size_t)kGwStages * stage_bytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(xpose + 4 * 32 * 33);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + kGwStages;
    uint64_t* tmem_full = bars + 2 * kGwStages;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kGwStages + 1);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int m_tiles = (p.n_out + kTcBlockM - 1) / kTcBlockM;
    const int m_tile = blockIdx.x % m_tiles;         // the m tiles of one split are neighbours: they share x in L2
    const int split = blockIdx.x / m_tiles;
    const int kb_begin = split * p.kb_per_split;
    const int kb_end = min(p.num_kb, kb_begin + p.kb_per_split);
    const int m0 = m_tile * kTcBlockM;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kGwStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "n"(kTcMaxN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_g_hi) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x_hi) : "memory");
            for (int kb = kb_begin, it = 0; kb < kb_end; ++kb, ++it) {
                const int s = it % kGwStages;
                const uint32_t ph = (it / kGwStages) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);
                unsigned char* st = ring + (size_t)s * stage_bytes;
                mbar_arrive_expect_tx(&full_bar[s], (uint32_t)(p.terms == 3 ? stage_bytes : stage_bytes / 2));
                const int k0 = kb * kGwBlockK;
                for (int j = 0; j < kGwABlocks; ++j) {       // blocks past n_out are zero-filled by TMA
                    tma_load_2d(st + j * kGwBoxBytes, &map_g_hi, m0 + j * kGwMnBlock, k0, &full_bar[s]);
                    if (p.terms == 3) tma_load_2d(st + a_bytes + j * kGwBoxBytes, &map_g_lo, m0 + j * kGwMnBlock, k0, &full_bar[s]);
                }
                unsigned char* sb = st + 2 * a_bytes;
                for (int j = 0; j < p.b_blocks; ++j) {
                    tma_load_2d(sb + j * kGwBoxBytes, &map_x_hi, j * kGwMnBlock, k0, &full_bar[s]);
                    if (p.terms == 3) tma_load_2d(sb + b_bytes + j * kGwBoxBytes, &map_x_lo, j * kGwMnBlock, k0, &full_bar[s]);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc1 = umma_idesc_tf32_mn(p.n1);
            const uint32_t idesc2 = umma_idesc_tf32_mn(p.n2 > 0 ? p.n2 : 16);
            for (int kb = kb_begin, it = 0; kb < kb_end; ++kb, ++it) {
                const int s = it % kGwStages;
                const uint32_t ph = (it / kGwStages) & 1;
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                unsigned char* st = ring + (size_t)s * stage_bytes;
                const uint64_t a_hi = umma_desc_mn_sw128(st), a_lo = umma_desc_mn_sw128(st + a_bytes);
                unsigned char* sb = st + 2 * a_bytes;
                const uint64_t b_hi = umma_desc_mn_sw128(sb), b_lo = umma_desc_mn_sw128(sb + b_bytes);
                const uint64_t n2_off = (uint64_t)(((p.n1 / kGwMnBlock) * kGwBoxBytes) >> 4);     // n1 = 256 -> 8 MN blocks
#pragma unroll
                for (int kk = 0; kk < kGwBlockK / kTcUmmaK; ++kk) {          // one 8-row K atom (1 KB) per MMA
                    const uint64_t ko = (uint64_t)((kk * 1024) >> 4);
                    const uint32_t acc0 = (it | kk) != 0;
                    tc_mma_tf32(tmem_base, a_hi + ko, b_hi + ko, idesc1, acc0);
                    if (p.terms == 3) {
                        tc_mma_tf32(tmem_base, a_lo + ko, b_hi + ko, idesc1, 1);
                        tc_mma_tf32(tmem_base, a_hi + ko, b_lo + ko, idesc1, 1);
                    }
                    if (p.n2 > 0) {
                        tc_mma_tf32(tmem_base + p.n1, a_hi + ko, b_hi + n2_off + ko, idesc2, acc0);
                        if (p.terms == 3) {
                            tc_mma_tf32(tmem_base + p.n1, a_lo + ko, b_hi + n2_off + ko, idesc2, 1);
                            tc_mma_tf32(tmem_base + p.n1, a_hi + ko, b_lo + n2_off + ko, idesc2, 1);
                        }
                    }
                }
                tc_commit(&empty_bar[s]);
            }
            tc_commit(tmem_full);
        }
    } else {
        const int q = warp & 3;
        float* xp = xpose + (warp - 2) * 32 * 33;
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        const int row0 = m0 + q * 32;
        float* dst = p.partial + (size_t)split * p.n_out * p.n_in;
        for (int c0 = 0; c0 < p.n_in; c0 += 32) {
            float v[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) xp[lane * 33 + j] = v[j];
            __syncwarp();
            const int c = c0 + lane;
            if (c < p.n_in) {
                for (int r = 0; r < 32; ++r) {
                    const int row = row0 + r;
                    if (row >= p.n_out) break;
                    dst[(size_t)row * p.n_in + c] = xp[r * 33 + lane];
                }
            }
            __syncwarp();
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kTcMaxN));
    }
}

__global__ void __launch_bounds__(256)
grad_w_reduce_kernel(const float* __restrict__ partial, int n_splits, int64_t n, float* __restrict__ out, int64_t ld_out, int n_in) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int k = 0; k < n_splits; ++k) s += partial[(size_t)k * n + i];
        out[(i / n_in) * ld_out + (i % n_in)] = s;
    }
}

// fp32 (tf32-valued) matrix [rows, cols], row pitch ld: box = [16 rows, 32 cols], 128-byte swizzle (MN-major operand).
inline int make_tf32_mn_map(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t ld) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return fail(WSAGE_ECUDA, "%s: %s", "wsage_grad_w_tc", "cuTensorMapEncodeTiled not available");
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    const cuuint32_t box[2] = {(cuuint32_t)kGwMnBlock, (cuuint32_t)kGwBlockK};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(WSAGE_ECUDA, "%s: %s", "wsage_grad_w_tc", "cuTensorMapEncodeTiled failed");
    return WSAGE_OK;
}

// The tensor core adds into its fp32 accumulator with truncation, so the error of one accumulation chain
// grows linearly with its length (measured: ~7e-9 relative per row, 1.5e-4 at 21k rows, 3e-6 at 400).  A CTA
// therefore never accumulates more than kGwMaxChainKb k-blocks (1024 rows); the partial tiles are added in
// fp32 round-to-nearest by grad_w_reduce_kernel.
constexpr int kGwMaxChainKb = 64;

inline int grad_w_splits(int64_t rows, int n_out) {
    const int m_tiles = (n_out + kTcBlockM - 1) / kTcBlockM;
    const int64_t num_kb = (rows + kGwBlockK - 1) / kGwBlockK;
    int64_t splits = kNumSMs / m_tiles;
    if (splits > num_kb / 32) splits = num_kb / 32;        // at least 32 k-blocks (512 rows) per CTA when filling the chip
    const int64_t prec = (num_kb + kGwMaxChainKb - 1) / kGwMaxChainKb;
    if (splits < prec) splits = prec;
    if (splits < 1) splits = 1;
    return (int)splits;
}

}  // namespace wsage
