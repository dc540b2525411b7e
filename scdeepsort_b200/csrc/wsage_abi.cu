// extern "C" entry points of libwsage.so (declared in include/wsage.h).
#include "agg_gather.cuh"
#include "agg_tiled.cuh"
#include "dense_tc.cuh"
#include "dense16.cuh"
#include "sampler.cuh"
#include "loss_adam.cuh"
#include "peer_reduce.cuh"
#include <math.h>

using namespace wsage;

extern "C" {

int wsage_version(void) { return 2001; }

const char* wsage_last_error(void) { return g_err; }

int64_t wsage_launch_count(int reset) {
    return reset ? g_launches.exchange(0) : g_launches.load();
}

int wsage_block_agg_fwd(const int64_t* rowptr, const int32_t* col, const float* w,
                        const int32_t* src_id, const int32_t* dst_id,
                        const float* alpha, int32_t gene_num,
                        const float* h_src, int64_t ld_src, int64_t n_src,
                        float* out, int64_t ld_out, int64_t n_dst, int32_t dim,
                        void* stream) {
    WSAGE_REQUIRE(n_dst >= 0 && n_src >= 0 && dim > 0, "negative size or dim <= 0");
    if (n_dst == 0) return WSAGE_OK;
    WSAGE_REQUIRE(rowptr && out, "null rowptr/out");
    WSAGE_REQUIRE(ld_src >= dim && ld_out >= dim, "leading dimension < dim");
    WSAGE_REQUIRE((src_id && dst_id && alpha) || (!src_id && !dst_id), "src_id, dst_id and alpha go together");
    GatherParams p{};
    p.rowptr = rowptr; p.col = col; p.w = w;
    p.src_id = src_id; p.dst_id = dst_id; p.alpha = alpha; p.gene_num = gene_num;
    p.hs = h_src; p.ld_hs = ld_src; p.n_dst = n_dst; p.dim = dim; p.mean = 1;
    p.out = out; p.ld_out = ld_out;
    const bool vec4 = dim % 4 == 0 && ld_src % 4 == 0 && ld_out % 4 == 0 && aligned16(h_src) && aligned16(out);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return src_id ? launch_gather_fwd<int32_t, true>(p, vec4, st) : launch_gather_fwd<int32_t, false>(p, vec4, st);
}

int wsage_block_agg_bwd(const int64_t* rowptr, const int32_t* col, const float* w,
                        const int32_t* src_id, const int32_t* dst_id,
                        const float* alpha, int32_t gene_num,
                        const float* h_src, int64_t ld_src, int64_t n_src,
                        const float* d_out, int64_t ld_dout, int64_t n_dst, int32_t dim,
                        float* d_h_src, int64_t ld_dh, float* d_alpha,
                        void* stream) {
    WSAGE_REQUIRE(n_dst >= 0 && n_src >= 0 && dim > 0, "negative size or dim <= 0");
    if (n_dst == 0 || (!d_h_src && !d_alpha)) return WSAGE_OK;
    WSAGE_REQUIRE(rowptr && d_out && src_id && dst_id && alpha, "null argument");
    WSAGE_REQUIRE(!d_alpha || h_src, "d_alpha needs h_src");
    WSAGE_REQUIRE(ld_dout >= dim && (!d_h_src || ld_dh >= dim) && (!h_src || ld_src >= dim), "leading dimension < dim");
    GatherBwdParams p{};
    p.rowptr = rowptr; p.col = col; p.w = w;
    p.src_id = src_id; p.dst_id = dst_id; p.alpha = alpha; p.gene_num = gene_num;
    p.hs = h_src; p.ld_hs = ld_src; p.dout = d_out; p.ld_dout = ld_dout;
    p.n_dst = n_dst; p.dim = dim; p.dh = d_h_src; p.ld_dh = ld_dh; p.dalpha = d_alpha;
    const bool vec4 = dim % 4 == 0 && ld_src % 4 == 0 && ld_dout % 4 == 0 && ld_dh % 4 == 0 &&
                      aligned16(h_src) && aligned16(d_out) && aligned16(d_h_src);
    return launch_gather_bwd(p, vec4, static_cast<cudaStream_t>(stream));
}

static int spmm_validate(const wsage_spmm_args* a) {
    WSAGE_REQUIRE(a != nullptr, "null args");
    WSAGE_REQUIRE(a->n_dst >= 0 && a->n_src >= 0 && a->dim > 0, "negative size or dim <= 0");
    WSAGE_REQUIRE(a->col_bits == WSAGE_COL_I32 || a->col_bits == WSAGE_COL_U16, "col_bits must be 16 or 32");
    WSAGE_REQUIRE(a->col_bits == WSAGE_COL_I32 || a->n_src <= 65536, "uint16 columns need n_src <= 65536");
    WSAGE_REQUIRE(a->algo >= 0 && a->algo <= 2, "algo must be 0, 1 or 2");
    WSAGE_REQUIRE(a->nnz >= 0, "negative nnz");
    if (a->n_dst == 0) return WSAGE_OK;
    WSAGE_REQUIRE(a->rowptr && a->hs, "null rowptr/hs");
    WSAGE_REQUIRE(a->out || a->raw || a->dot, "no output requested");
    WSAGE_REQUIRE(a->ld_hs >= a->dim, "ld_hs < dim");
    WSAGE_REQUIRE(!a->selfcoef || (a->hself && a->ld_hself >= a->dim), "selfcoef needs hself");
    WSAGE_REQUIRE(!a->dot || (a->q && a->ld_q >= a->dim), "dot needs q");
    WSAGE_REQUIRE(!a->out || a->ld_out >= a->dim, "ld_out < dim");
    WSAGE_REQUIRE(!a->raw || a->ld_raw >= a->dim, "ld_raw < dim");
    if (a->init) {
        WSAGE_REQUIRE(a->init_slabs > 0 && a->init_rows > 0, "init needs init_slabs, init_rows > 0");
        WSAGE_REQUIRE(a->init_map || a->init_rows >= a->n_dst, "init_map == NULL needs init_rows >= n_dst");
        WSAGE_REQUIRE(aligned16(a->init), "init must be 16-byte aligned");
        WSAGE_REQUIRE(a->algo != 1, "init needs the tiled kernel (algo 0 or 2)");
    }
    return WSAGE_OK;
}

static bool spmm_vec4(const wsage_spmm_args* a) {
    return a->dim % 4 == 0 && a->ld_hs % 4 == 0 && aligned16(a->hs) &&
           (!a->out || (a->ld_out % 4 == 0 && aligned16(a->out))) &&
           (!a->raw || (a->ld_raw % 4 == 0 && aligned16(a->raw))) &&
           (!a->hself || (a->ld_hself % 4 == 0 && aligned16(a->hself))) &&
           (!a->q || (a->ld_q % 4 == 0 && aligned16(a->q)));
}

static bool spmm_use_tiled(const wsage_spmm_args* a, bool vec4) {
    if (a->init) return a->nnz > 0;
    return a->algo == 2 || (a->algo == 0 && tiled_profitable(a, vec4));
}

size_t wsage_spmm_workspace_bytes(const wsage_spmm_args* a) {
    if (!a || spmm_validate(a) != WSAGE_OK || a->n_dst == 0) return 0;
    const bool vec4 = spmm_vec4(a);
    if (!spmm_use_tiled(a, vec4) || !tiled_supported(a, vec4)) return 0;
    return tiled_plan(a).workspace_bytes;
}

int wsage_spmm_algo(const wsage_spmm_args* a) {
    if (!a || spmm_validate(a) != WSAGE_OK) return 0;
    if (a->algo != 0) return a->algo;
    if (a->init) return 2;
    return tiled_profitable(a, spmm_vec4(a)) ? 2 : 1;
}

static void fill_epilogue(TiledParams& p, const wsage_spmm_args* a) {
    p.n_dst = a->n_dst; p.dim = a->dim;
    p.dscale = a->dscale; p.selfcoef = a->selfcoef; p.hself = a->hself; p.ld_hself = a->ld_hself;
    p.out = a->out; p.ld_out = a->ld_out; p.raw = a->raw; p.ld_raw = a->ld_raw;
    p.q = a->q; p.ld_q = a->ld_q; p.dot = a->dot;
}

int wsage_spmm(const wsage_spmm_args* a, void* stream) {
    const int rc = spmm_validate(a);
    if (rc != WSAGE_OK) return rc;
    if (a->n_dst == 0) return WSAGE_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool vec4 = spmm_vec4(a);
    if (a->init && !(vec4 && a->dim <= 512))
        return fail(WSAGE_EUNSUPPORTED, "%s: %s", "wsage_spmm", "init needs dim % 4 == 0, dim <= 512 and 16-byte aligned rows");
    if (a->init && a->nnz == 0) {
        // every entry of the pass sits in the dense block: sum its slabs, apply the epilogue
        TiledParams p{};
        fill_epilogue(p, a);
        p.init = a->init; p.init_slabs = a->init_slabs; p.init_rows = a->init_rows; p.init_map = a->init_map;
        dense16_finalize_kernel<<<gather_grid(a->n_dst), 256, 0, st>>>(p);
        return check_launch("dense16_finalize");
    }
    int algo = a->algo;
    if (algo == 0) algo = (a->init || tiled_profitable(a, vec4)) ? 2 : 1;
    if (algo == 2) {
        if (!tiled_supported(a, vec4))
            return fail(WSAGE_EUNSUPPORTED, "%s: %s", "wsage_spmm", "tiled kernel needs dim % 4 == 0, contiguous 16-byte aligned hs and dim <= 512");
        const size_t need = tiled_plan(a).workspace_bytes;
        if (need > a->workspace_bytes || (need && !a->workspace))
            return fail(WSAGE_EINVAL, "%s: %s", "wsage_spmm", "workspace too small (see wsage_spmm_workspace_bytes)");
        const TiledInit ini{a->init, a->init ? a->init_slabs : 0, a->init_rows, a->init_map};
        return launch_tiled(a, ini, st);
    }
    GatherParams p{};
    p.rowptr = a->rowptr; p.col = a->col; p.w = a->x;
    p.hs = a->hs; p.ld_hs = a->ld_hs; p.n_dst = a->n_dst; p.dim = a->dim; p.mean = 0;
    p.dscale = a->dscale; p.selfcoef = a->selfcoef; p.hself = a->hself; p.ld_hself = a->ld_hself;
    p.out = a->out; p.ld_out = a->ld_out; p.raw = a->raw; p.ld_raw = a->ld_raw;
    p.q = a->q; p.ld_q = a->ld_q; p.dot = a->dot; p.row_perm = a->row_perm;
    return a->col_bits == WSAGE_COL_U16 ? launch_gather_fwd<uint16_t, false>(p, vec4, st)
                                        : launch_gather_fwd<int32_t, false>(p, vec4, st);
}

// ---- dense block of the popular genes on the tensor cores ---------------------------------------------------------
int wsage_amax(const float* x, int64_t ld, const int32_t* row_ids, const float* rowscale,
               int64_t rows, int32_t cols, float* amax, void* stream) {
    WSAGE_REQUIRE(rows >= 0 && cols > 0 && cols % 4 == 0, "cols must be a positive multiple of 4");
    WSAGE_REQUIRE(amax, "null amax");
    if (rows == 0) return WSAGE_OK;
    WSAGE_REQUIRE(x && ld >= cols && ld % 4 == 0 && aligned16(x), "x must be 16-byte aligned with ld % 4 == 0");
    const int64_t total = rows * (cols / 4);
    int64_t grid = (total + 255) / 256;
    if (grid > (int64_t)kNumSMs * 8) grid = (int64_t)kNumSMs * 8;
    amax_kernel<<<(int)grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, ld, row_ids, rowscale, rows, cols, reinterpret_cast<unsigned*>(amax));
    return check_launch("amax");
}

int wsage_split16(const float* x, int64_t ld, const int32_t* row_ids, const float* rowscale,
                  int64_t rows, int32_t cols, const float* amax, int32_t fmt, int32_t layout,
                  void* hi, void* lo, int64_t ld_out, void* stream) {
    return wsage_split16_masked(x, ld, nullptr, 0, row_ids, rowscale, rows, cols, amax, fmt, layout, hi, lo, ld_out, stream);
}

int wsage_split16_masked(const float* x, int64_t ld, const float* mask_src, int64_t ld_mask, const int32_t* row_ids, const float* rowscale,
                         int64_t rows, int32_t cols, const float* amax, int32_t fmt, int32_t layout,
                         void* hi, void* lo, int64_t ld_out, void* stream) {
    WSAGE_REQUIRE(rows >= 0 && cols > 0 && cols % 4 == 0, "cols must be a positive multiple of 4");
    WSAGE_REQUIRE(fmt == WSAGE_D16_F16X2 || fmt == WSAGE_D16_BF16, "unknown fmt");
    WSAGE_REQUIRE(layout >= WSAGE_SPLIT_ROWS && layout <= WSAGE_SPLIT_BLOCKED, "unknown layout");
    if (rows == 0) return WSAGE_OK;
    WSAGE_REQUIRE(x && hi && (lo || fmt == WSAGE_D16_BF16), "null pointer");
    WSAGE_REQUIRE(ld >= cols && ld % 4 == 0 && aligned16(x), "x must be 16-byte aligned with ld % 4 == 0");
    WSAGE_REQUIRE(aligned16(hi) && aligned16(lo), "planes must be 16-byte aligned");
    const bool transposed = layout == WSAGE_SPLIT_TRANSPOSED || layout == WSAGE_SPLIT_KBLOCKS;
    WSAGE_REQUIRE(!mask_src || (!transposed && ld_mask >= cols && ld_mask % 4 == 0 && aligned16(mask_src)), "mask_src needs a row layout and a 16-byte aligned mask");
    WSAGE_REQUIRE(layout == WSAGE_SPLIT_COLBLOCKS || layout == WSAGE_SPLIT_KBLOCKS || ld_out % 8 == 0, "ld_out % 8 != 0");
    WSAGE_REQUIRE(layout != WSAGE_SPLIT_BLOCKED || ld_out % 32 == 0, "blocked layout: ld_out (slot padding) must be a multiple of 32");
    WSAGE_REQUIRE(ld_out >= ((layout == WSAGE_SPLIT_ROWS || layout == WSAGE_SPLIT_KBLOCKS || layout == WSAGE_SPLIT_BLOCKED) ? (int64_t)cols : rows), "ld_out too small");
    WSAGE_REQUIRE(transposed || !row_ids, "row_ids needs a transposed layout");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned short* h = static_cast<unsigned short*>(hi);
    unsigned short* l = static_cast<unsigned short*>(lo);
    if (transposed) {
        WSAGE_REQUIRE(rows < ((int64_t)1 << 31) * 32, "too many rows");
        // few source rows (a weight matrix): spread the column tiles over blocks too; many: one block per 32 rows walks all of them
        const unsigned gx = (unsigned)((rows + 31) / 32);
        const unsigned gy = gx >= 4u * kNumSMs ? 1u : (unsigned)((cols + 31) / 32);
        if (layout == WSAGE_SPLIT_KBLOCKS && gy == 1u) {
            split16_kblocks_wide_kernel<<<gx, 128, 0, st>>>(x, ld, row_ids, rowscale, rows, cols, amax, fmt, h, l, ld_out);
            return check_launch("split16_kblocks_wide");
        }
        split16_transpose_kernel<<<dim3(gx, gy), 128, 0, st>>>(x, ld, row_ids, rowscale, rows, cols, amax, fmt, h, l, ld_out, layout == WSAGE_SPLIT_KBLOCKS ? 1 : 0);
        return check_launch("split16_transpose");
    }
    if (layout == WSAGE_SPLIT_BLOCKED) {
        const int nb_used = (cols + 31) / 32;
        int64_t groups = (int64_t)kNumSMs * 8 / nb_used;
        if (groups < 1) groups = 1;
        if (groups > (rows + 127) / 128) groups = (rows + 127) / 128;
        const int grid = (int)(groups * nb_used);
        split16_blocked_kernel<false><<<grid, 256, 0, st>>>(x, ld, mask_src, ld_mask, rowscale, rows, cols, amax, fmt, h, l, ld_out, nullptr);
        return check_launch("split16_blocked");
    }
    const int64_t cols_p = (cols + 31) / 32 * 32;
    const int64_t total = (rows + 127) / 128 * 128 * (cols_p / 4);
    int64_t grid = (total + 255) / 256;
    if (grid > (int64_t)kNumSMs * 16) grid = (int64_t)kNumSMs * 16;
    split16_kernel<<<(int)grid, 256, 0, st>>>(x, ld, mask_src, ld_mask, rowscale, rows, cols, amax, fmt, h, l, ld_out, layout);
    return check_launch("split16");
}

int wsage_split16_colsum(const float* x, int64_t ld, const float* mask_src, int64_t ld_mask, int64_t rows, int32_t cols,
                         const float* amax, int32_t fmt, void* hi, void* lo, int64_t ld_out,
                         float* partial, int32_t n_partial, float* colsum, void* stream) {
    WSAGE_REQUIRE(rows > 0 && cols > 0 && cols % 4 == 0, "rows must be positive, cols a positive multiple of 4");
    WSAGE_REQUIRE(fmt == WSAGE_D16_F16X2 || fmt == WSAGE_D16_BF16, "unknown fmt");
    WSAGE_REQUIRE(x && hi && (lo || fmt == WSAGE_D16_BF16) && partial && colsum && n_partial >= 1, "null pointer");
    WSAGE_REQUIRE(ld >= cols && ld % 4 == 0 && aligned16(x), "x must be 16-byte aligned with ld % 4 == 0");
    WSAGE_REQUIRE(aligned16(hi) && aligned16(lo) && aligned16(partial) && aligned16(colsum), "outputs must be 16-byte aligned");
    WSAGE_REQUIRE(!mask_src || (ld_mask >= cols && ld_mask % 4 == 0 && aligned16(mask_src)), "bad mask");
    WSAGE_REQUIRE(ld_out % 32 == 0 && ld_out >= cols, "ld_out (slot padding) must be a multiple of 32, at least cols");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int nb_used = (cols + 31) / 32;
    int64_t groups = (int64_t)kNumSMs * 8 / nb_used;
    if (groups < 1) groups = 1;
    if (groups > n_partial) groups = n_partial;
    if (groups > (rows + 127) / 128) groups = (rows + 127) / 128;
    split16_blocked_kernel<true><<<(int)(groups * nb_used), 256, 0, st>>>(x, ld, mask_src, ld_mask, nullptr, rows, cols, amax, fmt,
                                                                          static_cast<unsigned short*>(hi), static_cast<unsigned short*>(lo), ld_out, partial);
    int rc = check_launch("split16_blocked (colsum)");
    if (rc != WSAGE_OK) return rc;
    // the groups' sums, added in group order
    sum_slabs_kernel<<<(cols / 4 + 255) / 256, 256, 0, st>>>(partial, (int)groups, cols, 1, cols, colsum, cols);
    return check_launch("sum_slabs");
}

int wsage_sum_slabs(const float* slabs, int32_t n_slabs, int64_t slab_stride, int64_t rows, int32_t cols,
                    float* out, int64_t ld_out, void* stream) {
    WSAGE_REQUIRE(n_slabs > 0 && rows >= 0 && cols > 0 && cols % 4 == 0, "bad shape");
    if (rows == 0) return WSAGE_OK;
    WSAGE_REQUIRE(slabs && out && aligned16(slabs) && aligned16(out), "null or misaligned pointer");
    WSAGE_REQUIRE(slab_stride >= rows * cols && slab_stride % 4 == 0 && ld_out >= cols && ld_out % 4 == 0, "bad stride");
    const int64_t total = rows * (cols / 4);
    int64_t grid = (total + 255) / 256;
    if (grid > (int64_t)kNumSMs * 16) grid = (int64_t)kNumSMs * 16;
    sum_slabs_kernel<<<(int)grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(slabs, n_slabs, slab_stride, rows, cols, out, ld_out);
    return check_launch("sum_slabs");
}

int wsage_colsum_masked(const float* x, int64_t ld, const float* mask_src, int64_t ld_mask, int64_t rows, int32_t cols,
                        float* partial, int32_t n_partial, float* out, void* stream) {
    WSAGE_REQUIRE(rows >= 0 && cols > 0 && cols % 4 == 0 && cols <= 1024, "cols must be a multiple of 4, at most 1024");
    WSAGE_REQUIRE(n_partial >= 1 && partial && out && aligned16(partial) && aligned16(out), "null or misaligned output");
    WSAGE_REQUIRE(rows == 0 || (x && ld >= cols && ld % 4 == 0 && aligned16(x)), "x must be 16-byte aligned with ld % 4 == 0");
    WSAGE_REQUIRE(!mask_src || (ld_mask >= cols && ld_mask % 4 == 0 && aligned16(mask_src)), "bad mask");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    colsum_masked_kernel<<<n_partial, 256, 0, st>>>(x, ld, mask_src, ld_mask, rows, cols, partial);
    int rc = check_launch("colsum_masked");
    if (rc != WSAGE_OK) return rc;
    sum_slabs_kernel<<<(cols / 4 + 255) / 256, 256, 0, st>>>(partial, n_partial, cols, 1, cols, out, cols);
    return check_launch("sum_slabs");
}

int wsage_rowdot(const float* a, int64_t ld_a, const float* b, int64_t ld_b, int64_t rows, int32_t cols, float* out, void* stream) {
    WSAGE_REQUIRE(rows >= 0 && cols > 0 && cols % 4 == 0, "cols must be a positive multiple of 4");
    if (rows == 0) return WSAGE_OK;
    WSAGE_REQUIRE(a && b && out && aligned16(a) && aligned16(b), "null or misaligned pointer");
    WSAGE_REQUIRE(ld_a >= cols && ld_b >= cols && ld_a % 4 == 0 && ld_b % 4 == 0, "bad leading dimension");
    rowdot_kernel<<<gather_grid(rows), 256, 0, static_cast<cudaStream_t>(stream)>>>(a, ld_a, b, ld_b, rows, cols, out);
    return check_launch("rowdot");
}

// ---- exchange step over peer memory (peer_reduce.cuh) ----
static size_t peer_set_bytes(int64_t max_elems) { return ((size_t)max_elems * sizeof(float) + 255) / 256 * 256; }

size_t wsage_peer_bytes(int64_t max_elems) { return max_elems > 0 ? kPeerHeaderBytes + 4 * peer_set_bytes(max_elems) : 0; }

#define WSAGE_CUDA(call, what)                                                                     \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString(e_));                \
            cudaGetLastError();                                                                    \
            return WSAGE_ECUDA;                                                                    \
        }                                                                                          \
    } while (0)

int wsage_peer_alloc(int64_t max_elems, void** base, void* ipc_handle64) {
    WSAGE_REQUIRE(max_elems > 0 && base && ipc_handle64, "max_elems must be positive, base and handle non-null");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the ABI carries the handle as 64 bytes");
    void* p = nullptr;
    const size_t bytes = wsage_peer_bytes(max_elems);
    WSAGE_CUDA(cudaMalloc(&p, bytes), "cudaMalloc (peer buffer)");
    cudaError_t e = cudaMemset(p, 0, bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(static_cast<cudaIpcMemHandle_t*>(ipc_handle64), p);
    if (e != cudaSuccess) {
        snprintf(g_err, sizeof(g_err), "peer buffer set-up: %s", cudaGetErrorString(e));
        cudaFree(p);
        cudaGetLastError();
        return WSAGE_ECUDA;
    }
    *base = p;
    return WSAGE_OK;
}

int wsage_peer_open(const void* ipc_handle64, void** base) {
    WSAGE_REQUIRE(ipc_handle64 && base, "null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle64, sizeof(h));
    WSAGE_CUDA(cudaIpcOpenMemHandle(base, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
    return WSAGE_OK;
}

int wsage_peer_close(void* base) {
    if (base) WSAGE_CUDA(cudaIpcCloseMemHandle(base), "cudaIpcCloseMemHandle");
    return WSAGE_OK;
}

int wsage_peer_free(void* base) {
    if (base) WSAGE_CUDA(cudaFree(base), "cudaFree (peer buffer)");
    return WSAGE_OK;
}

int wsage_peer_status(const void* base, int32_t* status) {
    WSAGE_REQUIRE(base && status, "null argument");
    WSAGE_CUDA(cudaMemcpy(status, static_cast<const unsigned char*>(base) + kPeerStatusOff, sizeof(int32_t), cudaMemcpyDeviceToHost),
               "cudaMemcpy (peer status)");
    return WSAGE_OK;
}

int wsage_peer_reduce(const wsage_peer_reduce_args* a, void* stream) {
    WSAGE_REQUIRE(a != nullptr, "null args");
    WSAGE_REQUIRE(a->world >= 1 && a->world <= kPeerMax && a->rank >= 0 && a->rank < a->world, "rank / world out of range");
    WSAGE_REQUIRE(a->rows > 0 && a->dim > 0 && a->dim % 4 == 0 && a->rows * a->dim <= a->max_elems, "rows * dim must fit max_elems, dim % 4 == 0");
    WSAGE_REQUIRE(a->epoch >= 1 && a->epoch % 2 == 1 && a->epoch < 0xfffffff0u / (2 * kNumSMs), "epoch must be odd: 1, 3, 5, ...");
    WSAGE_REQUIRE(a->grid >= 0 && a->grid <= 2 * kNumSMs, "grid must be 0 (two CTAs per SM) or at most two CTAs per SM");
    WSAGE_REQUIRE(a->bases && a->slabs && aligned16(a->slabs) && a->n_slabs >= 1 && a->slab_rows >= 1, "null bases / slabs");
    WSAGE_REQUIRE(a->slot_of_row || a->slab_rows >= a->rows, "slab_rows < rows without a row map");
    WSAGE_REQUIRE(a->out || a->raw, "neither out nor raw");
    WSAGE_REQUIRE(!a->out || (aligned16(a->out) && a->ld_out >= a->dim && a->ld_out % 4 == 0), "bad out");
    WSAGE_REQUIRE(!a->raw || (aligned16(a->raw) && a->ld_raw >= a->dim && a->ld_raw % 4 == 0), "bad raw");
    WSAGE_REQUIRE(!a->selfcoef || (a->hself && aligned16(a->hself) && a->ld_hself >= a->dim && a->ld_hself % 4 == 0), "selfcoef needs hself");
    PeerReduceParams p{};
    p.rank = a->rank; p.world = a->world;
    const size_t set = peer_set_bytes(a->max_elems);
    const int parity = (int)((a->epoch >> 1) & 1u);
    for (int r = 0; r < a->world; ++r) {
        WSAGE_REQUIRE(a->bases[r] != nullptr, "null peer base");
        unsigned char* b = static_cast<unsigned char*>(a->bases[r]);
        p.header[r] = b;
        p.partial[r] = reinterpret_cast<float*>(b + kPeerHeaderBytes + (size_t)parity * set);
        p.result[r] = reinterpret_cast<float*>(b + kPeerHeaderBytes + (size_t)(2 + parity) * set);
    }
    p.slabs = a->slabs; p.n_slabs = a->n_slabs; p.slab_rows = a->slab_rows; p.slot_of_row = a->slot_of_row;
    p.rows = a->rows; p.dim = a->dim; p.epoch = a->epoch;
    p.dscale = a->dscale; p.selfcoef = a->selfcoef; p.hself = a->hself; p.ld_hself = a->ld_hself;
    p.out = a->out; p.ld_out = a->ld_out; p.raw = a->raw; p.ld_raw = a->ld_raw;
    p.timeout_ns = (unsigned long long)((a->timeout_s > 0.f ? a->timeout_s : 60.f) * 1e9);
    // two CTAs per SM, all resident: the barriers inside spin on a grid-wide counter
    peer_reduce_kernel<<<a->grid > 0 ? a->grid : 2 * kNumSMs, kPeerThreads, 0, static_cast<cudaStream_t>(stream)>>>(p);
    return check_launch("peer_reduce");
}

int wsage_dense16_slots_pad(int32_t gene_slots) { return gene_slots > 0 ? d16_slots_pad(gene_slots) : 0; }

static int dense16_validate(const wsage_dense16_args* a, bool need_out = true) {
    WSAGE_REQUIRE(a != nullptr, "null args");
    WSAGE_REQUIRE(a->fmt == WSAGE_D16_F16X2 || a->fmt == WSAGE_D16_BF16, "unknown fmt");
    WSAGE_REQUIRE(a->side == 0 || a->side == 1, "side must be 0 or 1");
    WSAGE_REQUIRE(a->cells > 0 && a->gene_slots > 0 && a->dim > 0, "cells, gene_slots and dim must be positive");
    WSAGE_REQUIRE(a->dim % 4 == 0 && a->dim <= kTcMaxN, "dim must be a multiple of 4, at most 512");
    WSAGE_REQUIRE(a->x_hi && (a->x_lo || a->fmt == WSAGE_D16_BF16) && a->h_hi && (a->h_lo || a->fmt == WSAGE_D16_BF16) && (a->out || !need_out), "null pointer");
    WSAGE_REQUIRE(a->x_scale > 0.f || a->x_amax, "x_scale must be positive (or x_amax given)");
    WSAGE_REQUIRE(!a->bias || aligned16(a->bias), "bias must be 16-byte aligned");
    WSAGE_REQUIRE(aligned16(a->h_hi) && aligned16(a->h_lo), "H planes must be 16-byte aligned");
    WSAGE_REQUIRE(a->ld_h >= ((a->dim + 15) & ~15) && a->ld_h < (1 << 20), "ld_h (rows per k-block of the H planes) must cover dim rounded up to 16");
    WSAGE_REQUIRE(aligned16(a->x_hi) && aligned16(a->x_lo) && aligned16(a->out), "X planes and out must be 16-byte aligned");
    WSAGE_REQUIRE(a->chunk_rows >= 0, "negative chunk_rows");
    if (a->side == 0) {
        WSAGE_REQUIRE(a->n_dst > 0 && a->n_dst <= a->cells, "side 0 needs 0 < n_dst <= cells");

        WSAGE_REQUIRE(a->ld_out >= a->dim && a->ld_out % 4 == 0, "ld_out must be >= dim and a multiple of 4");
        WSAGE_REQUIRE(!a->selfcoef || (a->hself && a->ld_hself >= a->dim && a->ld_hself % 4 == 0 && aligned16(a->hself)), "selfcoef needs a 16-byte aligned hself");
    } else {
        WSAGE_REQUIRE(a->n_src_cells > 0 && a->n_src_cells <= a->cells, "side 1 needs 0 < n_src_cells <= cells");

        WSAGE_REQUIRE(!a->dscale && !a->selfcoef && !a->bias && !a->relu, "side 1 writes raw partial sums (epilogue in wsage_spmm)");
    }
    const int64_t storage_rows = ((a->cells + kD16TileM - 1) / kD16TileM) * (d16_slots_pad(a->gene_slots) / kD16BlockK) * kD16TileM;
    WSAGE_REQUIRE(storage_rows < ((int64_t)1 << 31), "dense block too large for 32-bit TMA coordinates");
    return WSAGE_OK;
}

int wsage_dense16_splits(const wsage_dense16_args* a) {
    if (dense16_validate(a, false) != WSAGE_OK) return 0;
    D16Plan pl{};
    d16_plan(a, pl);
    return pl.n_splits;
}

int wsage_dense16(const wsage_dense16_args* a, void* stream) {
    int rc = dense16_validate(a);
    if (rc != WSAGE_OK) return rc;
    D16Plan pl{};
    d16_plan(a, pl);
    if (pl.stages < 2) return fail(WSAGE_EUNSUPPORTED, "%s: %s", "wsage_dense16", "dim too large for a two-stage shared-memory ring");
    const bool bf = a->fmt == WSAGE_D16_BF16;
    const CUtensorMapDataType dt = bf ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
    const int slots_pad = d16_slots_pad(a->gene_slots);
    const uint64_t storage_rows = (uint64_t)((a->cells + kD16TileM - 1) / kD16TileM) * (uint64_t)pl.nb * kD16TileM;
    const void* x_lo = bf ? a->x_hi : a->x_lo;
    const void* h_lo = bf ? a->h_hi : a->h_lo;
    CUtensorMap ma_hi, ma_lo, mb1_hi, mb1_lo, mb2_hi, mb2_lo, mo;
    const char* what = "wsage_dense16";
    D16Params p{};
    p.side = a->side; p.terms = bf ? 1 : 3; p.bf16 = bf ? 1 : 0;
    p.n = a->dim; p.n_pad = pl.n_pad; p.n1 = pl.n1; p.n2 = pl.n2;
    p.stages = pl.stages; p.stage_bytes = pl.stage_bytes; p.tx_bytes = pl.tx_bytes; p.b_bytes = pl.b_bytes; p.nbuf = pl.nbuf;
#ifdef WSAGE_TUNING
    { static const int dd = [] { const char* e = getenv("WSAGE_D16_DRAIN_DIAG"); return e ? atoi(e) : 0; }(); p.drain_diag = dd; }
#endif
    p.m_tiles = pl.m_tiles; p.nb = pl.nb; p.num_kb = pl.num_kb; p.chunk_kb = pl.chunk_kb;
    p.n_splits = pl.n_splits; p.kb_per_split = pl.kb_per_split;
    p.full_items = pl.full_items; p.tail_tiles = pl.tail_tiles; p.tail_splits = pl.tail_splits; p.tail_kb = pl.tail_kb;
    p.amax = bf ? nullptr : a->h_amax;
    p.x_amax = bf ? nullptr : a->x_amax;
    p.x_scale_inv = a->x_scale > 0.f ? 1.f / a->x_scale : 1.f;
    p.bias = a->bias; p.relu = a->relu;
    if (a->relu && pl.num_kb > pl.chunk_kb)
        return fail(WSAGE_EUNSUPPORTED, "%s: %s", "wsage_dense16", "relu needs the whole k range in one accumulation chain");
    // B = H^T in k-blocks [num_kb][ld_h rows][32] (WSAGE_SPLIT_KBLOCKS): a 2-D tensor of 64-byte rows
    const uint64_t b_rows = (uint64_t)pl.num_kb * (uint64_t)a->ld_h;
    const uint32_t h2_box = (uint32_t)(pl.h2 > 0 ? pl.h2 : pl.h1);
    p.ld_hb = (int)a->ld_h;
    if ((rc = make_map_2d(&mb1_hi, dt, 2, a->h_hi, 32, b_rows, 64, 32, (uint32_t)pl.h1, what)) != WSAGE_OK) return rc;
    if ((rc = make_map_2d(&mb1_lo, dt, 2, h_lo, 32, b_rows, 64, 32, (uint32_t)pl.h1, what)) != WSAGE_OK) return rc;
    if ((rc = make_map_2d(&mb2_hi, dt, 2, a->h_hi, 32, b_rows, 64, 32, h2_box, what)) != WSAGE_OK) return rc;
    if ((rc = make_map_2d(&mb2_lo, dt, 2, h_lo, 32, b_rows, 64, 32, h2_box, what)) != WSAGE_OK) return rc;
    if (a->side == 0) {
        if ((rc = make_map_2d(&ma_hi, dt, 2, a->x_hi, 32, storage_rows, 64, 32, kD16TileM, what)) != WSAGE_OK) return rc;
        if ((rc = make_map_2d(&ma_lo, dt, 2, x_lo, 32, storage_rows, 64, 32, kD16TileM, what)) != WSAGE_OK) return rc;
        if ((rc = make_map_2d(&mo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a->out, (uint64_t)a->dim, (uint64_t)a->n_dst, (uint64_t)a->ld_out * 4, kD16OutCols, 32, what)) != WSAGE_OK) return rc;
        p.rows_per_split = 0; p.m_total = a->n_dst;
        p.dscale = a->dscale; p.selfcoef = a->selfcoef; p.hself = a->hself; p.ld_hself = a->ld_hself;
    } else {
        // A: {32 slots, storage rows (64 B apart), 4 gene blocks (128 rows = 8 KB apart)}
        if ((rc = make_map_3d(&ma_hi, dt, a->x_hi, storage_rows, 4, 64, (uint64_t)kD16TileM * 64, 4, what)) != WSAGE_OK) return rc;
        if ((rc = make_map_3d(&ma_lo, dt, x_lo, storage_rows, 4, 64, (uint64_t)kD16TileM * 64, 4, what)) != WSAGE_OK) return rc;
        if ((rc = make_map_2d(&mo, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, a->out, (uint64_t)a->dim, (uint64_t)pl.n_splits * slots_pad, (uint64_t)a->dim * 4, kD16OutCols, 32, what)) != WSAGE_OK) return rc;
        p.rows_per_split = slots_pad; p.m_total = slots_pad;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int tiles_per_item = pl.pair ? 2 : 1;
    const int64_t items = (int64_t)pl.full_items + (int64_t)pl.tail_tiles * pl.tail_splits;
    if (pl.tail_tiles > 0) {        // the k-split tiles of the last round only ever reduce-add
        const int64_t row0 = (int64_t)pl.full_items * tiles_per_item * kD16TileM;
        cudaError_t e = cudaMemset2DAsync(a->out + row0 * a->ld_out, (size_t)a->ld_out * 4, 0, (size_t)a->dim * 4, (size_t)(a->n_dst - row0), st);
        if (e != cudaSuccess) return fail(WSAGE_ECUDA, "%s: %s", "cudaMemset2DAsync(dense16 tail rows)", cudaGetErrorString(e));
    }
    if (pl.pair) {
        auto kern = dense16_kernel<true>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bytes);
        if (e != cudaSuccess) return fail(WSAGE_ECUDA, "%s: %s", "cudaFuncSetAttribute(dense16 pair)", cudaGetErrorString(e));
        const int groups = (int)(items < kNumSMs / 2 ? items : kNumSMs / 2);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)(2 * groups));
        cfg.blockDim = dim3(kD16Threads);
        cfg.dynamicSmemBytes = pl.smem_bytes;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr; cfg.numAttrs = 1;
        e = cudaLaunchKernelEx(&cfg, kern, ma_hi, ma_lo, mb1_hi, mb1_lo, mb2_hi, mb2_lo, mo, p);
        if (e != cudaSuccess) return fail(WSAGE_ECUDA, "%s: %s", "cudaLaunchKernelEx(dense16 pair)", cudaGetErrorString(e));
        return check_launch("dense16 (CTA pairs)");
    }
    auto kern = dense16_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem_bytes);
    if (e != cudaSuccess) return fail(WSAGE_ECUDA, "%s: %s", "cudaFuncSetAttribute(dense16)", cudaGetErrorString(e));
    const int grid = (int)(items < kNumSMs ? items : kNumSMs);
    kern<<<grid, kD16Threads, pl.smem_bytes, st>>>(ma_hi, ma_lo, mb1_hi, mb1_lo, mb2_hi, mb2_lo, mo, p);
    return check_launch("dense16");
}

int wsage_split_tf32(const float* x, int64_t ld_x, const float* mask_src, int64_t ld_mask,
                     float* hi, float* lo, int64_t ld_out, float* masked, int64_t ld_masked,
                     int64_t rows, int32_t cols, void* stream) {
    WSAGE_REQUIRE(rows >= 0 && cols > 0 && cols % 4 == 0, "cols must be a positive multiple of 4");
    if (rows == 0) return WSAGE_OK;
    WSAGE_REQUIRE(x && hi, "null x/hi");
    WSAGE_REQUIRE(ld_x >= cols && ld_out >= cols && ld_x % 4 == 0 && ld_out % 4 == 0, "bad leading dimension");
    WSAGE_REQUIRE(!mask_src || (ld_mask >= cols && ld_mask % 4 == 0), "bad mask leading dimension");
    WSAGE_REQUIRE(!masked || (mask_src && ld_masked >= cols && ld_masked % 4 == 0), "masked output needs mask_src");
    WSAGE_REQUIRE(aligned16(x) && aligned16(mask_src) && aligned16(masked) && aligned16(hi) && aligned16(lo), "misaligned pointer");
    const int64_t total = rows * (cols / 4);
    int64_t grid = (total + 255) / 256;
    if (grid > (int64_t)kNumSMs * 16) grid = (int64_t)kNumSMs * 16;
    split_tf32_kernel<<<(int)grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
        x, ld_x, mask_src, ld_mask, hi, lo, ld_out, masked, ld_masked, rows, cols);
    return check_launch("split_tf32");
}

int wsage_linear_tc(const float* a_hi, const float* a_lo, int64_t ld_a,
                    const float* b_hi, const float* b_lo, int64_t ld_b,
                    const float* bias, int32_t relu, float* out, int64_t ld_out,
                    int64_t m, int32_t n, int32_t k, void* stream) {
    WSAGE_REQUIRE(m >= 0 && n > 0 && k > 0, "bad shape");
    if (m == 0) return WSAGE_OK;
    WSAGE_REQUIRE(a_hi && b_hi && out, "null pointer");
    WSAGE_REQUIRE((a_lo != nullptr) == (b_lo != nullptr), "a_lo and b_lo go together (both NULL: single tf32 product)");
    const int terms = a_lo ? 3 : 1;
    if (!a_lo) { a_lo = a_hi; b_lo = b_hi; }             // descriptors only: never loaded when terms == 1
    // the MMA computes n_pad = N rounded up to 16 columns; rows of B past N are zero-filled by TMA
    const int n_pad = (n + 15) & ~15;
    WSAGE_REQUIRE(n_pad <= kTcMaxN, "N must be <= 512");
    WSAGE_REQUIRE(ld_a >= k && ld_b >= k && ld_a % 4 == 0 && ld_b % 4 == 0 && ld_out >= n, "bad leading dimension");
    WSAGE_REQUIRE(aligned16(a_hi) && aligned16(a_lo) && aligned16(b_hi) && aligned16(b_lo), "operands must be 16-byte aligned");
    WSAGE_REQUIRE(m < ((int64_t)1 << 31) - kTcBlockM, "M too large");
    LinearTcParams p{};
    p.terms = terms;
    p.m = m; p.n = n; p.n_pad = n_pad; p.k = k;
    p.n1 = n_pad < 256 ? n_pad : 256;
    p.n2 = n_pad - p.n1;
    p.b_boxes = (n_pad + 255) / 256;
    p.b_box_rows = n_pad / p.b_boxes;          // n_pad % 16 == 0, so two boxes are multiples of 8 rows
    p.bias = bias; p.relu = relu; p.out = out; p.ld_out = ld_out;
    p.num_tiles = (int)((m + kTcBlockM - 1) / kTcBlockM);
    const size_t smem = LinearTcSmem::total(n_pad);
    if (smem > 227 * 1024)
        return fail(WSAGE_EUNSUPPORTED, "%s: %s", "wsage_linear_tc", "N too large for the 3-stage shared-memory ring");
    CUtensorMap ma_hi, ma_lo, mb_hi, mb_lo;
    int rc;
    if ((rc = make_tf32_map(&ma_hi, a_hi, m, k, ld_a, kTcBlockM)) != WSAGE_OK) return rc;
    if ((rc = make_tf32_map(&ma_lo, a_lo, m, k, ld_a, kTcBlockM)) != WSAGE_OK) return rc;
    if ((rc = make_tf32_map(&mb_hi, b_hi, n, k, ld_b, p.b_box_rows)) != WSAGE_OK) return rc;
    if ((rc = make_tf32_map(&mb_lo, b_lo, n, k, ld_b, p.b_box_rows)) != WSAGE_OK) return rc;
    cudaError_t e = cudaFuncSetAttribute(linear_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(WSAGE_ECUDA, "%s: %s", "cudaFuncSetAttribute(linear_tc)", cudaGetErrorString(e));
    const int grid = p.num_tiles < kNumSMs ? p.num_tiles : kNumSMs;
    linear_tc_kernel<<<grid, kTcThreads, smem, static_cast<cudaStream_t>(stream)>>>(ma_hi, ma_lo, mb_hi, mb_lo, p);
    return check_launch("linear_tc");
}

int wsage_grad_w_splits(int64_t rows, int32_t n_out) {
    if (rows <= 0 || n_out <= 0) return 0;
    const int s = grad_w_splits(rows, n_out);
    const int64_t num_kb = (rows + kGwBlockK - 1) / kGwBlockK;
    const int64_t per = (num_kb + s - 1) / s;
    return (int)((num_kb + per - 1) / per);             // every split owns at least one k-block
}

int wsage_grad_w_tc(const float* g_hi, const float* g_lo, int64_t ld_g,
                    const float* x_hi, const float* x_lo, int64_t ld_x,
                    int64_t rows, int32_t n_out, int32_t n_in,
                    float* partial, int32_t n_splits, float* out, int64_t ld_out, void* stream) {
    WSAGE_REQUIRE(rows > 0 && n_out > 0 && n_in > 0, "bad shape");
    WSAGE_REQUIRE(g_hi && x_hi && partial && out, "null pointer");
    WSAGE_REQUIRE((g_lo != nullptr) == (x_lo != nullptr), "g_lo and x_lo go together (both NULL: single tf32 product)");
    const int terms = g_lo ? 3 : 1;
    if (!g_lo) { g_lo = g_hi; x_lo = x_hi; }
    WSAGE_REQUIRE(n_out % 4 == 0 && n_in % 4 == 0, "n_out and n_in must be multiples of 4");
    const int n_pad = (n_in + 15) & ~15;
    WSAGE_REQUIRE(n_pad <= kTcMaxN, "n_in must be <= 512");
    WSAGE_REQUIRE(ld_g >= n_out && ld_x >= n_in && ld_g % 4 == 0 && ld_x % 4 == 0 && ld_out >= n_in, "bad leading dimension");
    WSAGE_REQUIRE(aligned16(g_hi) && aligned16(g_lo) && aligned16(x_hi) && aligned16(x_lo), "operands must be 16-byte aligned");
    WSAGE_REQUIRE(rows < ((int64_t)1 << 31) - kGwBlockK, "too many rows");
    WSAGE_REQUIRE(n_splits == wsage_grad_w_splits(rows, n_out), "n_splits must come from wsage_grad_w_splits");
    GradWParams p{};
    p.terms = terms;
    p.rows = rows; p.n_out = n_out; p.n_in = n_in; p.n_pad = n_pad;
    p.n1 = n_pad < 256 ? n_pad : 256;
    p.n2 = n_pad - p.n1;
    p.b_blocks = (n_pad + kGwMnBlock - 1) / kGwMnBlock;
    p.num_kb = (int)((rows + kGwBlockK - 1) / kGwBlockK);
    p.n_splits = n_splits;
    p.kb_per_split = (p.num_kb + n_splits - 1) / n_splits;
    p.partial = partial;
    const size_t smem = GradWSmem::total(p.b_blocks);
    if (smem > 227 * 1024)
        return fail(WSAGE_EUNSUPPORTED, "%s: %s", "wsage_grad_w_tc", "n_in too large for the 3-stage shared-memory ring");
    CUtensorMap mg_hi, mg_lo, mx_hi, mx_lo;
    int rc;
    if ((rc = make_tf32_mn_map(&mg_hi, g_hi, rows, n_out, ld_g)) != WSAGE_OK) return rc;
    if ((rc = make_tf32_mn_map(&mg_lo, g_lo, rows, n_out, ld_g)) != WSAGE_OK) return rc;
    if ((rc = make_tf32_mn_map(&mx_hi, x_hi, rows, n_in, ld_x)) != WSAGE_OK) return rc;
    if ((rc = make_tf32_mn_map(&mx_lo, x_lo, rows, n_in, ld_x)) != WSAGE_OK) return rc;
    cudaError_t e = cudaFuncSetAttribute(grad_w_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(WSAGE_ECUDA, "%s: %s", "cudaFuncSetAttribute(grad_w_tc)", cudaGetErrorString(e));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int m_tiles = (n_out + kTcBlockM - 1) / kTcBlockM;
    grad_w_tc_kernel<<<m_tiles * n_splits, kTcThreads, smem, st>>>(mg_hi, mg_lo, mx_hi, mx_lo, p);
    rc = check_launch("grad_w_tc");
    if (rc != WSAGE_OK) return rc;
    const int64_t n = (int64_t)n_out * n_in;
    grad_w_reduce_kernel<<<(int)((n + 255) / 256), 256, 0, st>>>(partial, n_splits, n, out, ld_out, n_in);
    return check_launch("grad_w_reduce");
}

int wsage_sample_neighbors(const int64_t* rowptr, const int64_t* nodes, int64_t n_nodes,
                           int32_t fanout, uint64_t seed, int64_t* out_eid, int32_t* out_deg,
                           void* stream) {
    WSAGE_REQUIRE(n_nodes >= 0, "negative n_nodes");
    WSAGE_REQUIRE(fanout >= 1 && fanout <= kSampleMaxFanout, "fanout must be in [1, 32]");
    if (n_nodes == 0) return WSAGE_OK;
    WSAGE_REQUIRE(rowptr && nodes && out_eid && out_deg, "null pointer");
    sample_neighbors_kernel<<<gather_grid(n_nodes), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        rowptr, nodes, n_nodes, fanout, seed, out_eid, out_deg);
    return check_launch("sample_neighbors");
}

int wsage_softmax_ce(const float* logits, int64_t ld, const int64_t* labels, int64_t m, int32_t k,
                     float* d_logits, int64_t ld_d, float* loss_partial, int32_t n_partial, void* stream) {
    WSAGE_REQUIRE(m >= 0 && k > 0 && n_partial >= 1, "bad shape");
    WSAGE_REQUIRE(loss_partial, "null loss_partial");
    WSAGE_REQUIRE(m == 0 || (logits && labels && ld >= k && (!d_logits || ld_d >= k)), "null pointer or bad leading dimension");
    softmax_ce_kernel<<<n_partial, 256, 0, static_cast<cudaStream_t>(stream)>>>(logits, ld, labels, m, k, d_logits, ld_d, loss_partial);
    return check_launch("softmax_ce");
}

int wsage_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                    double lr, double beta1, double beta2, double eps, double weight_decay, int32_t step,
                    void* stream) {
    WSAGE_REQUIRE(n >= 0 && step >= 1, "n < 0 or step < 1");
    if (n == 0) return WSAGE_OK;
    WSAGE_REQUIRE(param && grad && exp_avg && exp_avg_sq, "null pointer");
    const double bc1 = 1.0 - pow(beta1, (double)step);
    const double bc2 = 1.0 - pow(beta2, (double)step);
    int64_t grid = (n + 255) / 256;
    if (grid > (int64_t)kNumSMs * 8) grid = (int64_t)kNumSMs * 8;
    adam_kernel<<<(int)grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(param, grad, exp_avg, exp_avg_sq, n, (float)(lr / bc1), (float)beta1, (float)(1.0 - beta1),
                                                                      (float)beta2, (float)(1.0 - beta2), (float)eps, (float)weight_decay, (float)sqrt(bc2));
    return check_launch("adam_step");
}

}  // extern "C"
