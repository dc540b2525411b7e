// Microbenchmarks behind the TMEM-sourced aggregation kernel (sm_100a, standalone: nvcc only, no torch).
//
//   layout : what tcgen05.cp.32x128b.warpx4 does to a contiguous 512-byte shared-memory chunk
//            (expected: lane l of every 32-lane quarter receives floats 4l..4l+3 in 4 consecutive columns)
//   ld     : tcgen05.ld (LDTM) throughput per SM for the access pattern of the kernel (warp-uniform dynamic
//            column, 12 columns per source row, two rows in flight), against the LDS.128 equivalent, and both
//            together; optionally with a warp issuing tcgen05.cp (UTCCP) at the same time
//   cp     : tcgen05.cp throughput (clk per 512-byte chunk replicated to the four lane quarters)
//
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/tmem_probe.bin tools/tmem_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

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
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {     // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {    // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// shared-memory matrix descriptor, no swizzle: 8-row x 16-byte core matrices, 128 B apart (SBO), LBO unused
__device__ __forceinline__ uint64_t cp_desc(const void* smem) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_u32(smem) & 0x3FFFF) >> 4);
    d |= (uint64_t)(128 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__device__ __forceinline__ void tmem_cp_32x128b_x4(uint32_t taddr, uint64_t desc) {   // one thread
    asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(taddr), "l"(desc) : "memory");
}
__device__ __forceinline__ void ldtm_x8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
}
__device__ __forceinline__ void ldtm_x4(uint32_t taddr, uint32_t (&r)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr));
}
__device__ __forceinline__ void ldtm_x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr));
}
__device__ __forceinline__ void ldtm_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// layout probe
// ------------------------------------------------------------------------------------------------
__global__ void layout_kernel(float* out /*[4 warps][32 lanes][8]*/) {
    __shared__ __align__(128) float chunk[256];
    __shared__ uint64_t bar;
    __shared__ uint32_t tbase;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) chunk[i] = i < 128 ? (float)i : (float)(1000 + i - 128);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) tmem_alloc(&tbase, 32);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic st.shared -> async-proxy reader
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t t = tbase;
    if (threadIdx.x == 0) {
        tmem_cp_32x128b_x4(t + 0, cp_desc(chunk));
        tmem_cp_32x128b_x4(t + 4, cp_desc(chunk + 128));
        tc_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    uint32_t r[8];
    ldtm_x8(t + ((uint32_t)(warp * 32) << 16), r);
    ldtm_wait();
    for (int i = 0; i < 8; ++i) out[(warp * 32 + lane) * 8 + i] = __uint_as_float(r[i]);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(t, 32);
}

// ------------------------------------------------------------------------------------------------
// ld throughput.  MODE 0: 12 TMEM columns per "edge"; 1: the LDS.128 x3 equivalent; 2: TMEM 12 columns +
// LDS.64 broadcast + LDS.32 remainder (the planned inner loop); 3: 8 TMEM columns + one LDS.128 + LDS.64.
// CPW: an extra warp keeps issuing tcgen05.cp windows (63 chunks + commit + wait) while the others read.
// ------------------------------------------------------------------------------------------------
constexpr int kRows = 42, kColsPerRow = 12, kPitch = 416;   // floats per shared-memory row (1664 B)

template <int MODE, bool CPW>
__global__ void __launch_bounds__(544, 1)
ld_kernel(int nw, int iters, float xin, long long* cyc, float* sink, long long* cp_count) {
    extern __shared__ __align__(128) float rows[];     // [kRows][kPitch]
    __shared__ uint64_t bar;
    __shared__ uint32_t tbase;
    __shared__ volatile int done;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < kRows * kPitch; i += blockDim.x) rows[i] = (float)(i % 97) * 0.01f;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); done = 0; asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) tmem_alloc(&tbase, 512);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t t = tbase;
    if (threadIdx.x == 0) {      // fill all 42 rows once
        for (int r = 0; r < kRows; ++r)
            for (int c = 0; c < 3; ++c) tmem_cp_32x128b_x4(t + r * kColsPerRow + c * 4, cp_desc(rows + r * kPitch + c * 128));
        tc_commit(&bar);
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    __syncthreads();

    if (CPW && warp == nw) {
        long long n = 0;
        uint32_t ph = 1;
        if (lane == 0) {
            while (!done) {
                for (int r = 0; r < 21; ++r)
                    for (int c = 0; c < 3; ++c) tmem_cp_32x128b_x4(t + r * kColsPerRow + c * 4, cp_desc(rows + r * kPitch + c * 128));
                tc_commit(&bar);
                mbar_wait(&bar, ph);
                ph ^= 1;
                n += 63;
            }
            cp_count[blockIdx.x] = n + 1;
        }
        return;
    }
    if (warp >= nw) return;

    float acc[13];
#pragma unroll
    for (int i = 0; i < 13; ++i) acc[i] = 0.f;
    const uint32_t tq = t + ((uint32_t)((warp & 3) * 32) << 16);
    int row = (warp * 5) % kRows;
    float x0 = xin, x1 = xin * 0.5f;
    const long long c0 = clock64();
    for (int it = 0; it < iters; it += 2) {
        int ra = row, rb = row + 7;
        if (rb >= kRows) rb -= kRows;
        row += 11;
        if (row >= kRows) row -= kRows;
        if (MODE == 0 || MODE == 2) {
            uint32_t a8[8], a4[4], b8[8], b4[4];
            ldtm_x8(tq + ra * kColsPerRow, a8);
            ldtm_x4(tq + ra * kColsPerRow + 8, a4);
            ldtm_x8(tq + rb * kColsPerRow, b8);
            ldtm_x4(tq + rb * kColsPerRow + 8, b4);
            float ar = 0.f, br = 0.f;
            if (MODE == 2) {
                const float2 ea = *reinterpret_cast<const float2*>(rows + ra * kPitch + 400);      // broadcast LDS.64
                const float2 eb = *reinterpret_cast<const float2*>(rows + rb * kPitch + 402);
                x0 = ea.x + xin; x1 = eb.y + xin;
                if (lane < 16) { ar = rows[ra * kPitch + 384 + lane]; br = rows[rb * kPitch + 384 + lane]; }
            }
            ldtm_wait();
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = fmaf(x1, __uint_as_float(b8[i]), fmaf(x0, __uint_as_float(a8[i]), acc[i]));
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[8 + i] = fmaf(x1, __uint_as_float(b4[i]), fmaf(x0, __uint_as_float(a4[i]), acc[8 + i]));
            if (MODE == 2) acc[12] = fmaf(x1, br, fmaf(x0, ar, acc[12]));
        } else if (MODE == 1) {
            const float* sa = rows + ra * kPitch + lane * 4;
            const float* sb = rows + rb * kPitch + lane * 4;
            float4 a[3], b[3];
#pragma unroll
            for (int j = 0; j < 3; ++j) { a[j] = *reinterpret_cast<const float4*>(sa + j * 128); b[j] = *reinterpret_cast<const float4*>(sb + j * 128); }
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                acc[j * 4 + 0] = fmaf(x1, b[j].x, fmaf(x0, a[j].x, acc[j * 4 + 0]));
                acc[j * 4 + 1] = fmaf(x1, b[j].y, fmaf(x0, a[j].y, acc[j * 4 + 1]));
                acc[j * 4 + 2] = fmaf(x1, b[j].z, fmaf(x0, a[j].z, acc[j * 4 + 2]));
                acc[j * 4 + 3] = fmaf(x1, b[j].w, fmaf(x0, a[j].w, acc[j * 4 + 3]));
            }
        } else {   // MODE 3: 8 TMEM columns + one LDS.128 + broadcast LDS.64
            uint32_t a8[8], b8[8];
            ldtm_x8(tq + ra * kColsPerRow, a8);
            ldtm_x8(tq + rb * kColsPerRow, b8);
            const float4 a = *reinterpret_cast<const float4*>(rows + ra * kPitch + 256 + lane * 4);
            const float4 b = *reinterpret_cast<const float4*>(rows + rb * kPitch + 256 + lane * 4);
            const float2 ea = *reinterpret_cast<const float2*>(rows + ra * kPitch + 400);
            const float2 eb = *reinterpret_cast<const float2*>(rows + rb * kPitch + 402);
            x0 = ea.x + xin; x1 = eb.y + xin;
            ldtm_wait();
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = fmaf(x1, __uint_as_float(b8[i]), fmaf(x0, __uint_as_float(a8[i]), acc[i]));
            acc[8] = fmaf(x1, b.x, fmaf(x0, a.x, acc[8]));
            acc[9] = fmaf(x1, b.y, fmaf(x0, a.y, acc[9]));
            acc[10] = fmaf(x1, b.z, fmaf(x0, a.z, acc[10]));
            acc[11] = fmaf(x1, b.w, fmaf(x0, a.w, acc[11]));
        }
    }
    const long long c1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 13; ++i) s += acc[i];
    sink[(blockIdx.x * 32 + warp) * 32 + lane] = s;
    if (lane == 0) atomicMax((unsigned long long*)&cyc[blockIdx.x], (unsigned long long)(c1 - c0));
    // all readers done -> stop the cp warp, then free TMEM
    tc_fence_before();
    asm volatile("bar.sync 1, %0;" ::"r"(nw * 32));
    if (warp == 0) {
        if (CPW) { done = 1; __threadfence_block(); }
    }
    if (!CPW && warp == 0) tmem_dealloc(t, 512);
    // with CPW the allocation is released at CTA exit after the cp warp drains (dealloc needs no cp in flight):
    if (CPW && warp == 0) {
        // wait until the cp warp has written its count (it exits right after)
        while (atomicAdd((unsigned long long*)&cp_count[blockIdx.x], 0ULL) == 0ULL) { }
        tmem_dealloc(t, 512);
    }
}

// ------------------------------------------------------------------------------------------------
// cp throughput: one thread issues `windows` windows of `chunks` cps, committing each and waiting with
// a lag of `lag` windows (0 = wait for each window before the next).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64, 1)
cp_kernel(int windows, int chunks, int lag, long long* cyc) {
    extern __shared__ __align__(128) float rows[];
    __shared__ uint64_t bar[2];
    __shared__ uint32_t tbase;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < kRows * kPitch; i += blockDim.x) rows[i] = (float)i;
    if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) tmem_alloc(&tbase, 512);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t t = tbase;
    if (threadIdx.x == 0) {
        const long long c0 = clock64();
        for (int w = 0; w < windows; ++w) {
            const int half = w & 1;
            if (w >= 2) mbar_wait(&bar[half], ((w >> 1) - 1) & 1);      // slot's previous commit
            for (int c = 0; c < chunks; ++c)
                tmem_cp_32x128b_x4(t + half * 252 + c * 4, cp_desc(rows + (c % 64) * 128 + half * 8192));
            tc_commit(&bar[half]);
            if (lag == 0) { mbar_wait(&bar[half], (w >> 1) & 1); }
        }
        // drain
        if (lag != 0) {
            for (int w = (windows >= 2 ? windows - 2 : 0); w < windows; ++w) mbar_wait(&bar[w & 1], (w >> 1) & 1);
        }
        cyc[blockIdx.x] = clock64() - c0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(t, 512);
}

template <int MODE, bool CPW>
void run_ld(const char* name, int nw, int iters, int sm_khz) {
    const int grid = 148;
    long long *cyc, *cpc;
    float* sink;
    CK(cudaMalloc(&cyc, grid * sizeof(long long)));
    CK(cudaMalloc(&cpc, grid * sizeof(long long)));
    CK(cudaMalloc(&sink, grid * 32 * 32 * sizeof(float)));
    const size_t smem = (size_t)kRows * kPitch * sizeof(float);
    auto k = ld_kernel<MODE, CPW>;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best_ms = 1e30f;
    std::vector<long long> h(grid), hc(grid);
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaMemset(cyc, 0, grid * sizeof(long long)));
        CK(cudaMemset(cpc, 0, grid * sizeof(long long)));
        CK(cudaEventRecord(e0));
        k<<<grid, (nw + (CPW ? 1 : 0)) * 32, smem>>>(nw, iters, 1.0001f, cyc, sink, cpc);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best_ms) best_ms = ms;
    }
    CK(cudaMemcpy(h.data(), cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hc.data(), cpc, grid * sizeof(long long), cudaMemcpyDeviceToHost));
    long long mx = 0, cps = 0;
    for (int i = 0; i < grid; ++i) { if (h[i] > mx) mx = h[i]; cps += hc[i]; }
    const double edges = (double)nw * iters;
    printf("ld %-28s warps=%2d  clk/edge/SM=%7.3f  rowbytes/clk/SM=%7.1f  (max cyc %lld, kernel %.3f ms)",
           name, nw, mx / edges, edges * 1536.0 / mx, mx, best_ms);
    if (CPW) printf("  cp chunks/SM=%lld -> %.2f clk per 512B chunk", cps / grid, (double)mx / ((double)cps / grid));
    printf("\n");
    CK(cudaFree(cyc)); CK(cudaFree(cpc)); CK(cudaFree(sink));
}

void run_cp(int windows, int chunks, int lag) {
    const int grid = 148;
    long long* cyc;
    CK(cudaMalloc(&cyc, grid * sizeof(long long)));
    const size_t smem = (size_t)kRows * kPitch * sizeof(float);
    CK(cudaFuncSetAttribute(cp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    std::vector<long long> h(grid);
    for (int rep = 0; rep < 2; ++rep) {
        cp_kernel<<<grid, 64, smem>>>(windows, chunks, lag, cyc);
        CK(cudaDeviceSynchronize());
    }
    CK(cudaMemcpy(h.data(), cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost));
    long long mx = 0;
    for (int i = 0; i < grid; ++i) if (h[i] > mx) mx = h[i];
    printf("cp windows=%d chunks/window=%d lag=%d: %.2f clk per 512B chunk (x4 quarters), %.0f clk per window\n",
           windows, chunks, lag, (double)mx / ((double)windows * chunks), (double)mx / windows);
    CK(cudaFree(cyc));
}

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    printf("device %s, %d SMs, clock %d kHz\n", prop.name, prop.multiProcessorCount, prop.clockRate);

    {   // layout
        float* out;
        CK(cudaMalloc(&out, 4 * 32 * 8 * sizeof(float)));
        layout_kernel<<<1, 128>>>(out);
        CK(cudaDeviceSynchronize());
        std::vector<float> h(4 * 32 * 8);
        CK(cudaMemcpy(h.data(), out, h.size() * sizeof(float), cudaMemcpyDeviceToHost));
        int bad = 0;
        for (int w = 0; w < 4; ++w)
            for (int l = 0; l < 32; ++l)
                for (int i = 0; i < 8; ++i) {
                    const float want = i < 4 ? (float)(4 * l + i) : (float)(1000 + 4 * l + i - 4);
                    if (h[(w * 32 + l) * 8 + i] != want) ++bad;
                }
        printf("layout: %d mismatches vs 'lane l <- floats 4l..4l+3, replicated to 4 quarters'\n", bad);
        for (int w = 0; w < 4; ++w) {
            printf("  warp %d lanes 0,1,2,31:", w);
            for (int l : {0, 1, 2, 31}) { printf(" ["); for (int i = 0; i < 8; ++i) printf("%g ", h[(w * 32 + l) * 8 + i]); printf("]"); }
            printf("\n");
        }
        CK(cudaFree(out));
    }
    const int iters = 200000;
    for (int nw : {4, 8, 12, 16}) {
        run_ld<0, false>("tmem12", nw, iters, prop.clockRate);
        run_ld<1, false>("lds(3xLDS.128)", nw, iters, prop.clockRate);
        run_ld<2, false>("tmem12+lds64+lds32", nw, iters, prop.clockRate);
        run_ld<3, false>("tmem8+lds128+lds64", nw, iters, prop.clockRate);
    }
    run_ld<2, true>("tmem12+lds64+lds32 +cpwarp", 12, iters, prop.clockRate);
    run_ld<0, true>("tmem12 +cpwarp", 12, iters, prop.clockRate);
    run_ld<1, true>("lds +cpwarp", 12, iters, prop.clockRate);
    run_cp(2000, 63, 0);
    run_cp(2000, 63, 1);
    run_cp(2000, 21, 1);
    return 0;
}
