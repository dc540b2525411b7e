// extern "C" entry points of libwsage.so (declared in include/wsage.h).
#include "agg_gather.cuh"
#include "agg_tiled.cuh"
#include "agg_dense.cuh"
#include "dense_tc.cuh"
#include "sampler.cuh"
#include "loss_adam.cuh"
#include <math.h>

using namespace wsage;

extern "C" {

int wsage_version(void) { return 1001; }

int wsage_dense_tile(void) { return kDenseT; }

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
    if (a->dense_x) {
        WSAGE_REQUIRE(a->dense_k > 0 && a->dense_t > 0, "dense block needs dense_k, dense_t > 0");
        WSAGE_REQUIRE(a->dense_src_ids || a->dense_k == a->n_src, "dense_src_ids == NULL needs dense_k == n_src");
        WSAGE_REQUIRE(a->dense_dst_map || a->dense_t == a->n_dst, "dense_dst_map == NULL needs dense_t == n_dst");
        WSAGE_REQUIRE(aligned16(a->dense_x), "dense_x must be 16-byte aligned");
        WSAGE_REQUIRE(a->algo != 1, "a dense block needs the tiled kernel (algo 0 or 2)");
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

// workspace = [tiled split partials, rounded up to 256 B][dense block sums]
static size_t spmm_tiled_ws(const wsage_spmm_args* a, bool vec4) {
    const bool tiled = dense_requested(a) ? tiled_supported(a, vec4) : (a->algo == 2 || (a->algo == 0 && tiled_profitable(a, vec4)));
    if (!tiled || !tiled_supported(a, vec4)) return 0;
    return (tiled_plan(a).workspace_bytes + 255) & ~(size_t)255;
}

size_t wsage_spmm_workspace_bytes(const wsage_spmm_args* a) {
    if (!a || spmm_validate(a) != WSAGE_OK || a->n_dst == 0) return 0;
    const bool vec4 = spmm_vec4(a);
    size_t n = spmm_tiled_ws(a, vec4);
    if (dense_requested(a) && tiled_supported(a, vec4)) n += dense_plan(a).out_bytes;
    return n;
}

int wsage_spmm_algo(const wsage_spmm_args* a) {
    if (!a || spmm_validate(a) != WSAGE_OK) return 0;
    if (a->algo != 0) return a->algo;
    if (dense_requested(a)) return 2;
    return tiled_profitable(a, spmm_vec4(a)) ? 2 : 1;
}

int wsage_spmm(const wsage_spmm_args* a, void* stream) {
    const int rc = spmm_validate(a);
    if (rc != WSAGE_OK) return rc;
    if (a->n_dst == 0) return WSAGE_OK;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool vec4 = spmm_vec4(a);
    int algo = a->algo;
    if (algo == 0) algo = (dense_requested(a) || tiled_profitable(a, vec4)) ? 2 : 1;
    if (algo == 2) {
        if (!tiled_supported(a, vec4))
            return fail(WSAGE_EUNSUPPORTED, "%s: %s", "wsage_spmm", "tiled kernel needs dim % 4 == 0, contiguous 16-byte aligned hs and dim <= 512");
        const size_t tiled_ws = spmm_tiled_ws(a, vec4);
        const size_t need = tiled_ws + (dense_requested(a) ? dense_plan(a).out_bytes : 0);
        if (need > a->workspace_bytes || (need && !a->workspace))
            return fail(WSAGE_EINVAL, "%s: %s", "wsage_spmm", "workspace too small (see wsage_spmm_workspace_bytes)");
        TiledInit ini{nullptr, 0, 0, nullptr};
        if (dense_requested(a)) {
            const DensePlan dp = dense_plan(a);
            float* dout = reinterpret_cast<float*>(static_cast<char*>(a->workspace) + tiled_ws);
            const int rc2 = launch_dense(a, dp, dout, st);
            if (rc2 != WSAGE_OK) return rc2;
            ini = TiledInit{dout, dp.n_splits, a->dense_t, a->dense_dst_map};
        }
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

int wsage_split_tf32(const float* x, int64_t ld_x, const float* mask_src, int64_t ld_mask,
                     float* hi, float* lo, int64_t ld_out, float* masked, int64_t ld_masked,
                     int64_t rows, int32_t cols, void* stream) {
    WSAGE_REQUIRE(rows >= 0 && cols > 0 && cols % 4 == 0, "cols must be a positive multiple of 4");
    if (rows == 0) return WSAGE_OK;
    WSAGE_REQUIRE(x && hi && lo, "null x/hi/lo");
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
    WSAGE_REQUIRE(a_hi && a_lo && b_hi && b_lo && out, "null pointer");
    // the MMA computes n_pad = N rounded up to 16 columns; rows of B past N are zero-filled by TMA
    const int n_pad = (n + 15) & ~15;
    WSAGE_REQUIRE(n_pad <= kTcMaxN, "N must be <= 512");
    WSAGE_REQUIRE(ld_a >= k && ld_b >= k && ld_a % 4 == 0 && ld_b % 4 == 0 && ld_out >= n, "bad leading dimension");
    WSAGE_REQUIRE(aligned16(a_hi) && aligned16(a_lo) && aligned16(b_hi) && aligned16(b_lo), "operands must be 16-byte aligned");
    WSAGE_REQUIRE(m < ((int64_t)1 << 31) - kTcBlockM, "M too large");
    LinearTcParams p{};
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
    WSAGE_REQUIRE(g_hi && g_lo && x_hi && x_lo && partial && out, "null pointer");
    WSAGE_REQUIRE(n_out % 4 == 0 && n_in % 4 == 0, "n_out and n_in must be multiples of 4");
    const int n_pad = (n_in + 15) & ~15;
    WSAGE_REQUIRE(n_pad <= kTcMaxN, "n_in must be <= 512");
    WSAGE_REQUIRE(ld_g >= n_out && ld_x >= n_in && ld_g % 4 == 0 && ld_x % 4 == 0 && ld_out >= n_in, "bad leading dimension");
    WSAGE_REQUIRE(aligned16(g_hi) && aligned16(g_lo) && aligned16(x_hi) && aligned16(x_lo), "operands must be 16-byte aligned");
    WSAGE_REQUIRE(rows < ((int64_t)1 << 31) - kGwBlockK, "too many rows");
    WSAGE_REQUIRE(n_splits == wsage_grad_w_splits(rows, n_out), "n_splits must come from wsage_grad_w_splits");
    GradWParams p{};
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
