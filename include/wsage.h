/*
 * wsage.h — C ABI of the B200-native weighted-GraphSAGE hot path (libwsage.so).
 *
 * Drop-in boundary for scDeepSort's edge-weighted message passing.  The reference has no
 * FFI of its own: the path is reached through Python (DGL 0.4.3 UDF protocol), so each
 * entry point cites the reference code it replaces.  All pointers are DEVICE pointers owned
 * by the caller (the library allocates nothing persistent); every call is asynchronous on
 * the caller's CUDA stream (`stream` is a cudaStream_t passed as void*); the library is
 * stateless and re-entrant.  Every function returns a status code (0 = ok) and never
 * throws; wsage_last_error() returns a thread-local message for the last non-zero status.
 *
 * Feature matrices are row-major [n_rows, dim] with a leading dimension `ld*` given in
 * ELEMENTS (ld >= dim), fp32.  Rows must be 4-byte aligned; 16-byte aligned rows with
 * dim % 4 == 0 take the vectorised path.
 */
#ifndef WSAGE_H
#define WSAGE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WSAGE_OK            0
#define WSAGE_EINVAL        1   /* bad argument (null pointer, negative size, ...)        */
#define WSAGE_EUNSUPPORTED  2   /* shape/dtype/alignment this build has no kernel for      */
#define WSAGE_ECUDA         3   /* a CUDA runtime call failed; see wsage_last_error()      */

#define WSAGE_COL_I32  32
#define WSAGE_COL_U16  16

/* Library / ABI version (major*1000 + minor). */
int wsage_version(void);
/* Thread-local, never NULL. */
const char* wsage_last_error(void);
/* Number of kernels this library has launched in this process since the last reset
 * (bench.py's `gpu_launches`). */
int64_t wsage_launch_count(int reset);

/* ---------------------------------------------------------------------------------------
 * Generic NodeFlow block (any sampled or full-neighbour block).
 *
 * Replaces, for one block: GNN.message_func (/root/reference/models/gnn.py:47-56, incl. the
 * host np.where α-index cascade and its two D2H syncs), fn.mean('m','neigh')
 * (models/gnn.py:65) and the DGL 0.4.3 src-gather / copy-reduce kernels behind
 * nf.block_compute.  The per-edge message tensor is never materialised.
 *
 *   out[v,:] = (1 / max(deg(v),1)) * SUM_{e in row v} w[e] * alpha[k(e)] * h_src[col[e],:]
 *   k(e) = src_id>=0 && dst_id<0 ? src_id : dst_id>=0 && src_id<0 ? dst_id
 *        : dst_id>=0 && src_id>=0 ? gene_num : gene_num+1
 *
 * CSR is destination-major: rowptr[n_dst+1] (int64), col[E] = local source index, w[E] =
 * edata['weight'].  src_id / dst_id are ndata['id'] of the source / destination layer
 * (int32; gene -> gene index, cell -> -1).  alpha is the [gene_num+2] parameter.
 * ------------------------------------------------------------------------------------- */
int wsage_block_agg_fwd(const int64_t* rowptr, const int32_t* col, const float* w,
                        const int32_t* src_id, const int32_t* dst_id,
                        const float* alpha, int32_t gene_num,
                        const float* h_src, int64_t ld_src, int64_t n_src,
                        float* out, int64_t ld_out, int64_t n_dst, int32_t dim,
                        void* stream);

/* Backward of wsage_block_agg_fwd (replaces torch autograd through the DGL UDF, i.e. the
 * scatter of models/gnn.py:54-56 and the index of self.alpha at :54).
 *   d_h_src[col[e],:] += s_v * w[e] * alpha[k(e)] * d_out[v,:]          (may be NULL)
 *   d_alpha[k(e)]     += s_v * w[e] * <h_src[col[e],:], d_out[v,:]>     (may be NULL)
 * Both outputs are ACCUMULATED with atomics; the caller zero-initialises them. */
int wsage_block_agg_bwd(const int64_t* rowptr, const int32_t* col, const float* w,
                        const int32_t* src_id, const int32_t* dst_id,
                        const float* alpha, int32_t gene_num,
                        const float* h_src, int64_t ld_src, int64_t n_src,
                        const float* d_out, int64_t ld_dout, int64_t n_dst, int32_t dim,
                        float* d_h_src, int64_t ld_dh, float* d_alpha,
                        void* stream);

/* ---------------------------------------------------------------------------------------
 * Full-graph bipartite pass (the throughput path; same reference lines as above, applied to
 * every destination of one kind at once instead of per 500-seed NodeFlow).
 *
 *   acc[v,:] = SUM_{e in row v} x[e] * hs[col[e],:]
 *   out[v,:] = dscale[v] * acc[v,:] + selfcoef[v] * hself[v,:]      (each term optional)
 *   raw[v,:] = acc[v,:]                                              (optional)
 *   dot[v]   = <acc[v,:], q[v,:]>                                    (optional)
 *
 * x holds RAW expression values; the reference's per-destination normalisation
 * (utils/preprocess_internal.py:15-23), the mean's 1/(deg+1), alpha and the self-loop
 * (preprocess_internal.py:213-214) enter through dscale / selfcoef and a pre-scaled hs, so one
 * CSR per direction serves forward and backward.  col is int32 or uint16 (col_bits), sorted
 * ascending inside each row.  row_perm (optional, int32[n_dst]) gives the order in which
 * destination rows are assigned to warps (load balancing); output rows are NOT permuted.
 * algo: 0 = auto, 1 = gather (L2 gather, warp per row), 2 = tiled (source windows staged in
 * shared memory by bulk-async copies, register-stationary destination tiles).
 * workspace: wsage_spmm_workspace_bytes() bytes of scratch (may be NULL when that is 0).
 * ------------------------------------------------------------------------------------- */
typedef struct wsage_spmm_args {
    const int64_t* rowptr;      /* [n_dst+1]                                            */
    const void*    col;         /* [nnz] int32 or uint16                                */
    int32_t        col_bits;    /* WSAGE_COL_I32 | WSAGE_COL_U16                        */
    const float*   x;           /* [nnz]                                                */
    int64_t        nnz;         /* = rowptr[n_dst]; lets algo 0 pick a kernel without a device read */
    const float*   hs;          /* [n_src, dim] source table                            */
    int64_t        ld_hs;
    int64_t        n_src;
    int64_t        n_dst;
    int32_t        dim;
    const float*   dscale;      /* [n_dst] or NULL (=1)                                 */
    const float*   selfcoef;    /* [n_dst] or NULL (no self term)                       */
    const float*   hself;       /* [n_dst, dim] (required iff selfcoef)                 */
    int64_t        ld_hself;
    float*         out;         /* [n_dst, dim] or NULL                                 */
    int64_t        ld_out;
    float*         raw;         /* [n_dst, dim] or NULL                                 */
    int64_t        ld_raw;
    const float*   q;           /* [n_dst, dim] (required iff dot)                      */
    int64_t        ld_q;
    float*         dot;         /* [n_dst] or NULL                                      */
    const int32_t* row_perm;    /* [n_dst] or NULL                                      */
    int32_t        algo;
    void*          workspace;
    size_t         workspace_bytes;
    /* Optional seed of the accumulators (ABI >= 2000): partial sums of entries that are not in the CSR — the
     * popular genes' dense block computed by wsage_dense16 — added in slab order before the CSR walk:
     *   acc[v,:] = SUM_k init[k][slot(v)][:] + SUM_{e in row v} ...     slot(v) = init_map[v] (negative: no seed for v. NULL: v)
     * With init the CSR may be empty (nnz == 0): the call then only sums the slabs and applies the epilogue.
     * Needs dim % 4 == 0, dim <= 512 and 16-byte aligned rows. */
    const float*   init;           /* [init_slabs][init_rows][dim] or NULL                   */
    int32_t        init_slabs;
    int64_t        init_rows;      /* rows per slab                                          */
    const int32_t* init_map;       /* [n_dst] or NULL                                        */
} wsage_spmm_args;

size_t wsage_spmm_workspace_bytes(const wsage_spmm_args* a);
/* The kernel wsage_spmm would run for these arguments: 1 (gather) or 2 (tiled); 0 on bad args. */
int wsage_spmm_algo(const wsage_spmm_args* a);
int wsage_spmm(const wsage_spmm_args* a, void* stream);

/* ---------------------------------------------------------------------------------------
 * Dense block of the popular genes on the tensor cores (tcgen05 / TMEM / TMA).
 *
 * Same reference lines as wsage_spmm (message_func + fn.mean, /root/reference/models/gnn.py:47-56,65), for the
 * entries X_dense of the expression matrix that the graph builder moved out of the CSRs (genes expressed in more
 * than a few percent of the cells).  X_dense is stored ONCE, zero-filled, as 16-bit tiles
 *   plane[cell / 128][slot / 32][cell % 128][slot % 32]      (slot = dense index of the gene, padded to 128)
 * and serves both directions:
 *   side 0  out[c,:]  = dscale[c] * SUM_s X[c,s] * H[gene(s),:] + selfcoef[c] * hself[c,:]       c < n_dst
 *   side 1  out[k][s,:] = SUM_{c in split k} X[c,s] * H[c,:]       partial sums, [n_splits][slots_pad][dim]
 * (side 1 feeds wsage_spmm's `init`).  fmt WSAGE_D16_F16X2: fp32-grade — X and H are given as fp16 hi + lo
 * planes of value * 2^k (x_scale for X, chosen at build time; H by wsage_split16 from its amax) and
 * hi*hi + lo*hi + hi*lo is accumulated in fp32, chains cut every chunk_rows rows.  fmt WSAGE_D16_BF16: bf16
 * planes, one product (lo pointers ignored).
 *
 * wsage_amax:    *amax = max(*amax, max |x[r,c] * rowscale[r]|) over rows r (or row_ids[r]) — caller zero-initialises.
 * wsage_split16: hi/lo = 16-bit split of x[r,c] * rowscale[r] * 2^k(amax), in one of three layouts:
 *                WSAGE_SPLIT_ROWS        planes [rows][ld_out]                     (ld_out % 8 == 0)
 *                WSAGE_SPLIT_TRANSPOSED  planes [cols][ld_out], the (gathered: row_ids) rows as columns
 *                WSAGE_SPLIT_COLBLOCKS   planes [ceil(cols / 32)][ld_out rows][32]: 32-column blocks of 64-byte rows
 *                WSAGE_SPLIT_BLOCKED     planes [ceil(rows / 128)][ld_out / 32][128][32]: the X-plane layout of wsage_dense16 for a
 *                                        dense activation / gradient matrix (ld_out = slot padding).  Zero-filled past the matrix
 *                                        inside the 32-column blocks that hold columns; blocks entirely past `cols` are left
 *                                        unwritten (no k-block of side 0 covers them, on side 1 they are output rows past the matrix)
 *                WSAGE_SPLIT_KBLOCKS     planes [ceil(rows / 32)][ld_out][32]: transposed, cut into k-blocks of 32 (gathered)
 *                                        rows — the H^T operand of wsage_dense16: one k-block of all columns is one contiguous
 *                                        piece.  Entries past `rows` in the last block are zeros; ld_out >= cols.
 * ------------------------------------------------------------------------------------- */
#define WSAGE_D16_F16X2 0
#define WSAGE_D16_BF16  1
#define WSAGE_SPLIT_ROWS        0
#define WSAGE_SPLIT_TRANSPOSED  1
#define WSAGE_SPLIT_COLBLOCKS   2
#define WSAGE_SPLIT_KBLOCKS     3
#define WSAGE_SPLIT_BLOCKED     4

int wsage_amax(const float* x, int64_t ld, const int32_t* row_ids, const float* rowscale,
               int64_t rows, int32_t cols, float* amax, void* stream);
int wsage_split16(const float* x, int64_t ld, const int32_t* row_ids, const float* rowscale,
                  int64_t rows, int32_t cols, const float* amax, int32_t fmt, int32_t layout,
                  void* hi, void* lo, int64_t ld_out, void* stream);
/* Same with v = x * (mask_src > 0): the ReLU backward of models/gnn.py:22 fused into the split (row layouts only). */
int wsage_split16_masked(const float* x, int64_t ld, const float* mask_src, int64_t ld_mask,
                         const int32_t* row_ids, const float* rowscale,
                         int64_t rows, int32_t cols, const float* amax, int32_t fmt, int32_t layout,
                         void* hi, void* lo, int64_t ld_out, void* stream);
/* wsage_split16_masked(layout = WSAGE_SPLIT_BLOCKED) and wsage_colsum_masked of the same x and mask in one pass over them
 * (ABI >= 2001): the split of a layer's output gradient for its dx / dW products and its bias gradient.  partial: scratch of
 * n_partial * cols floats (any n_partial >= 1; more rows = more CTAs, 90 are plenty), colsum: [cols]. */
int wsage_split16_colsum(const float* x, int64_t ld, const float* mask_src, int64_t ld_mask, int64_t rows, int32_t cols,
                         const float* amax, int32_t fmt, void* hi, void* lo, int64_t ld_out,
                         float* partial, int32_t n_partial, float* colsum, void* stream);
/* Bias gradient of a Linear (+ ReLU) layer (autograd through nn.Linear's bias, models/gnn.py:13,21-22):
 * out[c] = SUM_r x[r,c] * (mask_src[r,c] > 0) (mask_src NULL: plain column sums).  partial: [n_partial][cols] scratch,
 * added in index order (deterministic).  cols % 4 == 0, cols <= 1024. */
int wsage_colsum_masked(const float* x, int64_t ld, const float* mask_src, int64_t ld_mask, int64_t rows, int32_t cols,
                        float* partial, int32_t n_partial, float* out, void* stream);
/* out[r] = <a[r,:], b[r,:]> (the self-loop terms of the alpha gradient: autograd through models/gnn.py:54-56 for the
 * gene-gene / cell-cell loops).  cols % 4 == 0, 16-byte aligned rows. */
int wsage_rowdot(const float* a, int64_t ld_a, const float* b, int64_t ld_b, int64_t rows, int32_t cols, float* out, void* stream);
/* out[r,:] = SUM_k slabs[k * slab_stride + r * cols ...] in slab order (the split-K slabs of wsage_dense16 side 1). */
int wsage_sum_slabs(const float* slabs, int32_t n_slabs, int64_t slab_stride, int64_t rows, int32_t cols,
                    float* out, int64_t ld_out, void* stream);

typedef struct wsage_dense16_args {
    const void*    x_hi;         /* 16-bit planes of X_dense (layout above)                              */
    const void*    x_lo;         /* NULL for WSAGE_D16_BF16                                              */
    int32_t        fmt;
    int64_t        cells;        /* cells the planes cover (storage: ceil(cells / 128) tiles)            */
    int32_t        gene_slots;   /* dense genes (storage: wsage_dense16_slots_pad(gene_slots) slots)     */
    float          x_scale;      /* stored value = x * x_scale (a power of two)                          */
    int32_t        side;
    const void*    h_hi;         /* H^T in k-blocks [ceil(K/32)][ld_h][32] (WSAGE_SPLIT_KBLOCKS); K = dense genes (side 0) / cells (side 1) */
    const void*    h_lo;
    int64_t        ld_h;         /* rows per k-block of the H planes, >= dim rounded up to 16            */
    const float*   h_amax;       /* device scalar wsage_split16 scaled by (NULL: unscaled)               */
    int32_t        dim;
    int64_t        n_dst;        /* side 0: destination cells (<= cells)                                 */
    int64_t        n_src_cells;  /* side 1: cells that send (<= cells)                                   */
    const float*   dscale;       /* side 0, optional                                                     */
    const float*   selfcoef;     /* side 0, optional (then hself)                                        */
    const float*   hself;
    int64_t        ld_hself;
    float*         out;          /* side 0: [n_dst, dim] (row pitch below), side 1: contiguous slabs   */
    int64_t        ld_out;
    int32_t        chunk_rows;   /* accumulation chain length, 0 = default (2048)                        */
    /* The same kernel as the post-aggregate dense layer (NodeUpdate.forward, models/gnn.py:18-25, and its gradients):
     * X planes = an activation / gradient matrix in the WSAGE_SPLIT_BLOCKED layout, scaled by its own amax. */
    const float*   x_amax;       /* device scalar the X planes were scaled by (then x_scale is ignored), or NULL */
    const float*   bias;         /* side 0: [dim] added to every row, or NULL                            */
    int32_t        relu;         /* side 0: max(., 0); needs the k range to fit one chain                */
    /* ABI >= 2001.  Side 0 cuts the tiles of the last, partial round of destination tiles along k and lets the pieces
     * reduce-add in arrival order (fp32 sums of those rows may differ in the last bit from run to run); 1 keeps every
     * tile whole: bitwise reproducible, up to one tile time slower per call. */
    int32_t        deterministic;
} wsage_dense16_args;

int wsage_dense16_slots_pad(int32_t gene_slots);
/* side 1: number of partial slabs the call writes (0 on bad arguments). */
int wsage_dense16_splits(const wsage_dense16_args* a);
int wsage_dense16(const wsage_dense16_args* a, void* stream);

/* ---------------------------------------------------------------------------------------
 * Post-aggregate dense layer on the tensor cores (tcgen05 / TMEM / TMA).
 *
 * Replaces NodeUpdate.forward (/root/reference/models/gnn.py:18-25: fc_neigh Linear + activation),
 * the classifier (models/gnn.py:67) and the input-gradient GEMM of their backward.  fp32 accuracy
 * is kept by splitting every operand into tf32 hi + tf32 lo (wsage_split_tf32) and accumulating
 * hi*hi + lo*hi + hi*lo in fp32 (residual ~2^-21 per product).
 *
 * wsage_split_tf32:  hi = rn_tf32(v), lo = rn_tf32(v - hi), stored as fp32, with v = x, or
 *   v = x * (mask_src > 0) when mask_src is given (ReLU backward, models/gnn.py:22); `masked`
 *   (optional) receives v in fp32.  cols % 4 == 0; all row pitches in elements, multiples of 4.
 * Passing NULL for BOTH lo operands (split: lo == NULL writes hi only) selects a single tf32 product — the
 * precision of the bf16 configuration (BASELINE configs[2]), not of the fp32 parity path.
 * wsage_linear_tc:   out[M,N] = act(A[M,K] * B[N,K]^T + bias), A/B given as hi/lo pairs with row
 *   pitches ld_a / ld_b, N <= 512 (computed on N rounded up to 16; the extra rows of B are
 *   zero-filled by TMA); relu != 0 applies max(.,0); bias may be NULL.
 * ------------------------------------------------------------------------------------- */
int wsage_split_tf32(const float* x, int64_t ld_x, const float* mask_src, int64_t ld_mask,
                     float* hi, float* lo, int64_t ld_out, float* masked, int64_t ld_masked,
                     int64_t rows, int32_t cols, void* stream);
int wsage_linear_tc(const float* a_hi, const float* a_lo, int64_t ld_a,
                    const float* b_hi, const float* b_lo, int64_t ld_b,
                    const float* bias, int32_t relu, float* out, int64_t ld_out,
                    int64_t m, int32_t n, int32_t k, void* stream);
/* Weight gradient of the same layer (replaces autograd's g^T x through the reference's nn.Linear,
 * models/gnn.py:13,21):  out[n_out, n_in] = g[rows, n_out]^T * x[rows, n_in], both operands as tf32 hi/lo
 * pairs from wsage_split_tf32 (row pitches ld_g / ld_x, multiples of 4), n_in <= 512, n_out and n_in
 * multiples of 4.  The row range is cut into n_splits = wsage_grad_w_splits(rows, n_out) pieces whose
 * partial products go to `partial` ([n_splits][n_out][n_in] floats, caller-owned) and are added in
 * split order (deterministic). */
int wsage_grad_w_splits(int64_t rows, int32_t n_out);
int wsage_grad_w_tc(const float* g_hi, const float* g_lo, int64_t ld_g,
                    const float* x_hi, const float* x_lo, int64_t ld_x,
                    int64_t rows, int32_t n_out, int32_t n_in,
                    float* partial, int32_t n_splits, float* out, int64_t ld_out, void* stream);

/* ---------------------------------------------------------------------------------------
 * GPU neighbour sampler.  Replaces dgl.contrib.sampling.NeighborSampler's per-hop draw
 * (/root/reference/train.py:71-78: expand_factor in-edges per node, uniform, without
 * replacement; DGL 0.4.3 _CAPI_UniformSampling on the host).  rowptr is the parent graph's
 * in-edge CSR row pointer; for destination node nodes[i] the chosen edge positions (absolute
 * indices into the CSR arrays, ascending) are written to out_eid[i*fanout .. i*fanout+out_deg[i])
 * with out_deg[i] = min(in-degree, fanout).  1 <= fanout <= 32.  Same (seed, node) -> same draw.
 * ------------------------------------------------------------------------------------- */
int wsage_sample_neighbors(const int64_t* rowptr, const int64_t* nodes, int64_t n_nodes,
                           int32_t fanout, uint64_t seed, int64_t* out_eid, int32_t* out_deg,
                           void* stream);

/* ---------------------------------------------------------------------------------------
 * Loss and optimiser step of the training inner loop.
 *
 * wsage_softmax_ce: CrossEntropyLoss(reduction='sum') of /root/reference/train.py:36,82.  One launch
 *   computes per-block loss partials (loss_partial[0..n_partial), summed by the caller in index order:
 *   deterministic) and, if d_logits != NULL, the gradient softmax(logits) - onehot(labels).
 *   labels are int64 class indices in [0, k); a label outside that range makes the loss NaN (nothing is read out of bounds).
 * wsage_adam_step: one torch.optim.Adam(lr, betas, eps, weight_decay) update of n parameters
 *   (train.py:34-35,85; L2 decay added to the gradient, bias correction with `step` >= 1).
 * ------------------------------------------------------------------------------------- */
int wsage_softmax_ce(const float* logits, int64_t ld, const int64_t* labels, int64_t m, int32_t k,
                     float* d_logits, int64_t ld_d, float* loss_partial, int32_t n_partial, void* stream);
int wsage_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                    double lr, double beta1, double beta2, double eps, double weight_decay, int32_t step,
                    void* stream);   /* hyper-parameters in double: 1-beta and lr/(1-beta1^step) are formed in fp64 as torch does */

/* ---------------------------------------------------------------------------------------
 * Exchange step of the cell-sharded full-graph pass over NVLink peer memory (one process per GPU, one node).
 *
 * The gene destinations of /root/reference/models/gnn.py:65 (`fn.mean` over ALL in-edges) receive messages from every
 * rank's cell shard: the per-rank sums of wsage_dense16 (side 1) must be added over the ranks.  wsage_peer_reduce
 * does the local split-K reduction, that sum (reduce-scatter + all-gather by peer loads, rank order, every rank
 * gets identical bits) and the epilogue  out = dscale * raw + selfcoef * hself  in ONE kernel; the reference has no
 * counterpart (single process), the NCCL form is  sum_slabs -> ncclAllReduce -> two element-wise kernels.
 *
 * Set-up, once per process: wsage_peer_alloc (device memory of wsage_peer_bytes(max_elems), zeroed, plus its 64-byte
 * cudaIpc handle), handles exchanged by the host (torch.distributed / MPI / files), wsage_peer_open on the others'.
 * Every rank then calls wsage_peer_reduce in the same order with the same rows / dim and epoch = 1, 3, 5, ...
 * A barrier that does not complete within timeout_s marks the allocation failed (wsage_peer_status != 0) and the
 * kernel returns: nothing hangs, the host raises.
 * ------------------------------------------------------------------------------------- */
#define WSAGE_PEER_MAX 8
size_t wsage_peer_bytes(int64_t max_elems);
int wsage_peer_alloc(int64_t max_elems, void** base, void* ipc_handle64);
int wsage_peer_open(const void* ipc_handle64, void** base);
int wsage_peer_close(void* base);                /* a base from wsage_peer_open  */
int wsage_peer_free(void* base);                 /* a base from wsage_peer_alloc */
int wsage_peer_status(const void* base, int32_t* status);   /* synchronises the device */

typedef struct wsage_peer_reduce_args {
    int32_t        rank;
    int32_t        world;        /* <= WSAGE_PEER_MAX                                                    */
    void* const*   bases;        /* host array of `world` allocation bases in rank order (own: from alloc) */
    int64_t        max_elems;    /* what the allocations were sized for, rows * dim <= max_elems         */
    uint32_t       epoch;        /* 1, 3, 5, ... identical on every rank                                 */
    const float*   slabs;        /* [n_slabs][slab_rows][dim] partial sums of this rank                  */
    int32_t        n_slabs;
    int64_t        slab_rows;
    const int32_t* slot_of_row;  /* slab row of output row r, or NULL for r                              */
    int64_t        rows;
    int32_t        dim;          /* % 4 == 0                                                             */
    const float*   dscale;       /* optional, per row                                                    */
    const float*   selfcoef;     /* optional, per row (then hself)                                       */
    const float*   hself;
    int64_t        ld_hself;
    float*         out;          /* [rows][ld_out] or NULL                                               */
    int64_t        ld_out;
    float*         raw;          /* [rows][ld_raw] the sum itself, or NULL                               */
    int64_t        ld_raw;
    float          timeout_s;    /* 0 = default (60 s)                                                   */
    int32_t        grid;         /* CTAs, 0 = two per SM: all must be resident, and every call on an allocation uses the same value */
} wsage_peer_reduce_args;

int wsage_peer_reduce(const wsage_peer_reduce_args* a, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* WSAGE_H */
