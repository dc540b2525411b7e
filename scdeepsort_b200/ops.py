"""PyTorch-facing wrappers of the C ABI (``include/wsage.h``): tensors in, tensors out.

PyTorch is plumbing here (device memory, streams, autograd bookkeeping); all aggregation
arithmetic runs in ``libwsage.so``.  No function in this module has a non-CUDA code path.
"""
import ctypes
from dataclasses import dataclass
from typing import Optional

import torch

from . import _lib


TIMING = None     # set to a list by bench.py to collect per-launch CUDA-event timings of wsage_spmm


def _stream():
    # the raw handle of torch's current stream: ~0.3 us, against ~10 us for torch.cuda.current_stream() (95 calls per sampled step)
    return ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice()))


def _timed(record, fn):
    """Runs fn(); with bench.py's TIMING list set, brackets it with CUDA events on the launching stream."""
    if TIMING is None:
        return fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    r = fn()
    e1.record()
    record["events"] = (e0, e1)
    TIMING.append(record)
    return r


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _check_mat(t: torch.Tensor, name: str, rows=None):
    if not (t.is_cuda and t.dtype == torch.float32 and t.dim() == 2 and (t.shape[1] == 1 or t.stride(1) == 1)):
        raise ValueError(f"{name}: expected a CUDA fp32 row-major matrix, got {t.dtype} {tuple(t.shape)} "
                         f"strides {t.stride()} on {t.device}")
    if rows is not None and t.shape[0] != rows:
        raise ValueError(f"{name}: expected {rows} rows, got {t.shape[0]}")


def _check_vec(t: torch.Tensor, name: str, dtype, n=None):
    if not (t.is_cuda and t.dtype == dtype and t.dim() == 1 and t.is_contiguous()):
        raise ValueError(f"{name}: expected a contiguous CUDA {dtype} vector, got {t.dtype} {tuple(t.shape)} on {t.device}")
    if n is not None and t.shape[0] != n:
        raise ValueError(f"{name}: expected {n} elements, got {t.shape[0]}")


@dataclass
class Block:
    """One NodeFlow block as destination-major CSR (device tensors).

    Replaces DGL's per-block edge frame (``nf.blocks[i].data['weight']`` +
    ``nf.block_edges``), reference call site models/gnn.py:65."""
    rowptr: torch.Tensor   # int64 [n_dst+1]
    col: torch.Tensor      # int32 [E] local source index (layer i)
    weight: torch.Tensor   # fp32  [E] edata['weight']
    n_src: int
    n_dst: int

    def validate(self):
        _check_vec(self.rowptr, "rowptr", torch.int64, self.n_dst + 1)
        _check_vec(self.col, "col", torch.int32)
        _check_vec(self.weight, "weight", torch.float32, self.col.shape[0])


class _BlockAggregate(torch.autograd.Function):
    """message_func + fn.mean of one block (models/gnn.py:47-56,65), fused, with its backward."""

    @staticmethod
    def forward(ctx, h, alpha, block: Block, src_id, dst_id, gene_num):
        _check_mat(h, "h", block.n_src)
        alpha_flat = alpha.reshape(-1)
        _check_vec(alpha_flat, "alpha", torch.float32, gene_num + 2)
        _check_vec(src_id, "src_id", torch.int32, block.n_src)
        _check_vec(dst_id, "dst_id", torch.int32, block.n_dst)
        dim = h.shape[1]
        out = torch.empty(block.n_dst, dim, device=h.device, dtype=torch.float32)
        lib = _lib.load()
        rec = dict(kind="block_fwd", edges=int(block.col.shape[0]), n_src=block.n_src, n_dst=block.n_dst, dim=dim)
        _timed(rec, lambda: _lib.check(lib.wsage_block_agg_fwd(_ptr(block.rowptr), _ptr(block.col), _ptr(block.weight),
                                                               _ptr(src_id), _ptr(dst_id), _ptr(alpha_flat), gene_num,
                                                               _ptr(h), h.stride(0), block.n_src,
                                                               _ptr(out), out.stride(0), block.n_dst, dim, _stream()),
                                       "wsage_block_agg_fwd"))
        ctx.save_for_backward(h, alpha_flat, src_id, dst_id)
        ctx.block, ctx.gene_num, ctx.alpha_shape = block, gene_num, alpha.shape
        return out

    @staticmethod
    def backward(ctx, d_out):
        h, alpha_flat, src_id, dst_id = ctx.saved_tensors
        block = ctx.block
        need_h, need_a = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        d_out = d_out.contiguous()
        dh = torch.zeros_like(h, memory_format=torch.contiguous_format) if need_h else None
        da = torch.zeros_like(alpha_flat) if need_a else None
        if need_h or need_a:
            lib = _lib.load()
            rec = dict(kind="block_bwd", edges=int(block.col.shape[0]), n_src=block.n_src, n_dst=block.n_dst, dim=h.shape[1],
                       need_h=bool(need_h), need_a=bool(need_a))
            _timed(rec, lambda: _lib.check(lib.wsage_block_agg_bwd(_ptr(block.rowptr), _ptr(block.col), _ptr(block.weight),
                                                                   _ptr(src_id), _ptr(dst_id), _ptr(alpha_flat), ctx.gene_num,
                                                                   _ptr(h), h.stride(0), block.n_src,
                                                                   _ptr(d_out), d_out.stride(0), block.n_dst, h.shape[1],
                                                                   _ptr(dh), dh.stride(0) if need_h else 0, _ptr(da), _stream()),
                                           "wsage_block_agg_bwd"))
        return dh, (da.reshape(ctx.alpha_shape) if need_a else None), None, None, None, None


def block_aggregate(h, alpha, block: Block, src_id, dst_id, gene_num: int):
    """neigh[v] = mean_{e→v} w_e · α[k(e)] · h[src_e]  for one NodeFlow block."""
    return _BlockAggregate.apply(h, alpha, block, src_id, dst_id, gene_num)


@dataclass
class DenseBlock:
    """Entries of the popular genes, taken out of the CSRs and stored ONCE, zero-filled, as 16-bit planes
    ``plane[cell // 128][slot // 32][cell % 128][slot % 32]`` (include/wsage.h, csrc/dense16.cuh); both
    directions of the bipartite pass read the same planes through wsage_dense16."""
    hi: torch.Tensor                             # int16 storage: fp16 (hi part of x * x_scale) or bf16 bits
    lo: Optional[torch.Tensor]                   # fp16 residual; None for bf16
    fmt: int                                     # _lib.D16_F16X2 | _lib.D16_BF16
    cells: int                                   # cells the planes cover
    gene_slots: int                              # dense genes
    slots_pad: int
    gene_ids: torch.Tensor                       # int32 [gene_slots] slot -> gene id
    slot_of_gene: torch.Tensor                   # int32 [num_genes] gene id -> slot, -1 = not in the block
    x_scale: float
    nnz: int = 0                                 # expression entries the block stands for (over all its cells)


@dataclass
class Csr:
    """Destination-major CSR of RAW expression values for the full-graph path."""
    rowptr: torch.Tensor            # int64 [n_dst+1]
    col: torch.Tensor               # int32 or uint16-as-int16 storage [nnz], ascending inside each row
    x: torch.Tensor                 # fp32 [nnz]
    n_src: int
    n_dst: int
    col_bits: int = _lib.COL_I32
    row_perm: Optional[torch.Tensor] = None   # int32 [n_dst] warp-assignment order (load balance)
    dense: Optional[DenseBlock] = None        # entries removed from col/x and handled on the tensor cores
    dense_side: int = 0                       # 0: rows are cells (block sources = genes); 1: rows are genes

    @property
    def nnz(self):
        return int(self.x.shape[0])


def rowdot(a, b):
    """out[r] = <a[r], b[r]> on the device (fp32, warp per row)."""
    if not (a.is_cuda and a.shape == b.shape and a.shape[1] % 4 == 0 and a.stride(1) == 1 and b.stride(1) == 1):
        return (a * b).sum(dim=1)
    out = torch.empty(a.shape[0], device=a.device, dtype=torch.float32)
    _lib.check(_lib.load().wsage_rowdot(_ptr(a), a.stride(0), _ptr(b), b.stride(0), a.shape[0], a.shape[1], _ptr(out), _stream()), "wsage_rowdot")
    return out


def amax_split16(x, fmt, *, row_ids=None, rowscale=None, layout=_lib.SPLIT_ROWS):
    """(hi, lo, amax, ld): 16-bit planes of x * rowscale * 2^k for wsage_dense16 (k from the amax, device-side), in the
    layout wsage_split16 documents (rows / transposed / 32-column blocks)."""
    lib = _lib.load()
    dev = x.device
    rows = int(row_ids.shape[0]) if row_ids is not None else x.shape[0]
    cols = x.shape[1]
    amax = None
    if fmt == _lib.D16_F16X2:
        amax = torch.zeros(1, device=dev, dtype=torch.float32)
        _lib.check(lib.wsage_amax(_ptr(x), x.stride(0), _ptr(row_ids), _ptr(rowscale), rows, cols, _ptr(amax), _stream()), "wsage_amax")
    if layout == _lib.SPLIT_TRANSPOSED:
        ld = (rows + 7) // 8 * 8
        shape = (cols, ld)
    elif layout == _lib.SPLIT_COLBLOCKS:
        ld = rows
        shape = ((cols + 31) // 32, rows, 32)
    elif layout == _lib.SPLIT_KBLOCKS:
        ld = (cols + 15) // 16 * 16
        shape = ((rows + 31) // 32, ld, 32)
    else:
        ld = (cols + 7) // 8 * 8
        shape = (rows, ld)
    hi = torch.empty(shape, device=dev, dtype=torch.int16)
    lo = torch.empty(shape, device=dev, dtype=torch.int16) if fmt == _lib.D16_F16X2 else None
    _lib.check(lib.wsage_split16(_ptr(x), x.stride(0), _ptr(row_ids), _ptr(rowscale), rows, cols, _ptr(amax), fmt,
                                 layout, _ptr(hi), _ptr(lo), ld, _stream()), "wsage_split16")
    return hi, lo, amax, ld


def dense16(block: DenseBlock, side: int, hs, *, n_dst=None, n_src_cells=None, dscale=None, selfcoef=None, hself=None,
            out=None, chunk_rows=0, src_scale=None, deterministic=None):
    """The dense block's share of one pass.  side 0: returns out[n_dst, dim] (= dscale·acc + selfcoef·hself);
    side 1: returns the partial slabs [n_splits, slots_pad, dim] for ``spmm(init=...)``."""
    lib = _lib.load()
    dim = hs.shape[1]
    dev = hs.device
    a = _lib.Dense16Args()
    a.x_hi, a.x_lo, a.fmt, a.cells, a.gene_slots, a.x_scale = _ptr(block.hi), _ptr(block.lo), block.fmt, block.cells, block.gene_slots, block.x_scale
    a.side, a.dim, a.chunk_rows = side, dim, chunk_rows
    # bitwise run-to-run reproducibility of the last partial round of side-0 tiles (see wsage_dense16_args.deterministic)
    a.deterministic = int(torch.are_deterministic_algorithms_enabled() if deterministic is None else deterministic)
    if side == 0:
        h_hi, h_lo, amax, ld = amax_split16(hs, block.fmt, row_ids=block.gene_ids, rowscale=src_scale, layout=_lib.SPLIT_KBLOCKS)
        a.n_dst = n_dst
        if out is None:
            out = torch.empty(n_dst, dim, device=dev, dtype=torch.float32)
        a.dscale, a.selfcoef = _ptr(dscale), _ptr(selfcoef)
        if selfcoef is not None:
            a.hself, a.ld_hself = _ptr(hself), hself.stride(0)
        a.out, a.ld_out = _ptr(out), out.stride(0)
    else:
        h_hi, h_lo, amax, ld = amax_split16(hs[:n_src_cells], block.fmt, rowscale=src_scale, layout=_lib.SPLIT_KBLOCKS)
        a.n_src_cells = n_src_cells
    a.h_hi, a.h_lo, a.ld_h, a.h_amax = _ptr(h_hi), _ptr(h_lo), ld, _ptr(amax)
    if side == 1:
        n_splits = int(lib.wsage_dense16_splits(ctypes.byref(a)))
        if n_splits <= 0:
            _lib.check(_lib.EINVAL, "wsage_dense16_splits")
        out = torch.empty(n_splits, block.slots_pad, dim, device=dev, dtype=torch.float32)
        a.out, a.ld_out = _ptr(out), dim
    rec = dict(kind="dense16", side=side, cells=int(n_dst if side == 0 else n_src_cells), gene_slots=block.gene_slots,
               slots_pad=block.slots_pad, dim=dim, fmt=block.fmt, dense_nnz=block.nnz)
    _timed(rec, lambda: _lib.check(lib.wsage_dense16(ctypes.byref(a), _stream()), "wsage_dense16"))
    return out


def spmm(csr: Csr, hs, *, dscale=None, selfcoef=None, hself=None, out=None, want_out=True,
         raw=None, want_raw=False, q=None, want_dot=False, algo=_lib.ALGO_AUTO, src_scale=None):
    """acc = Σ_e x_e·src_scale[col_e]·hs[col_e];  out = dscale·acc + selfcoef·hself;  raw = acc;  dot = <acc, q>.

    Returns (out, raw, dot) with None for the ones not requested.  Entries held by ``csr.dense`` run on the
    tensor-core kernel first; the CSR walk (or, when the CSR is empty, a plain reduction) adds the rest.
    ``src_scale`` (optional, one factor per source row) is folded into the dense block's operand split; only a CSR
    remainder needs the scaled table materialised."""
    _check_mat(hs, "hs", csr.n_src)
    dim = hs.shape[1]
    dev = hs.device
    if src_scale is not None:
        _check_vec(src_scale, "src_scale", torch.float32, csr.n_src)
        if csr.dense is None or csr.nnz > 0:
            hs = hs * src_scale[:, None]
            src_scale = None
    if dscale is not None:
        _check_vec(dscale, "dscale", torch.float32, csr.n_dst)
    if selfcoef is not None:
        _check_vec(selfcoef, "selfcoef", torch.float32, csr.n_dst)
        _check_mat(hself, "hself", csr.n_dst)
    if want_out and out is None:
        out = torch.empty(csr.n_dst, dim, device=dev, dtype=torch.float32)
    if out is not None:
        _check_mat(out, "out", csr.n_dst)
    init = init_map = None
    if csr.dense is not None:
        d = csr.dense
        if csr.dense_side == 0:
            direct = csr.nnz == 0 and out is not None and raw is None and not want_raw and not want_dot
            if direct:      # the whole pass is the dense block: its last drain applies the epilogue
                dense16(d, 0, hs, n_dst=csr.n_dst, dscale=dscale, selfcoef=selfcoef, hself=hself, out=out, src_scale=src_scale)
                return out, None, None
            init = dense16(d, 0, hs, n_dst=csr.n_dst, src_scale=src_scale).unsqueeze(0)
        else:
            init = dense16(d, 1, hs, n_src_cells=csr.n_src, src_scale=src_scale)
            init_map = d.slot_of_gene
    if want_raw and raw is None:
        raw = torch.empty(csr.n_dst, dim, device=dev, dtype=torch.float32)
    dot = torch.empty(csr.n_dst, device=dev, dtype=torch.float32) if want_dot else None
    a = _lib.SpmmArgs()
    a.rowptr, a.col, a.col_bits, a.x, a.nnz = _ptr(csr.rowptr), _ptr(csr.col), csr.col_bits, _ptr(csr.x), csr.nnz
    a.hs, a.ld_hs, a.n_src, a.n_dst, a.dim = _ptr(hs), hs.stride(0), csr.n_src, csr.n_dst, dim
    if dscale is not None:
        a.dscale = _ptr(dscale)
    if selfcoef is not None:
        a.selfcoef, a.hself, a.ld_hself = _ptr(selfcoef), _ptr(hself), hself.stride(0)
    if out is not None:
        a.out, a.ld_out = _ptr(out), out.stride(0)
    if raw is not None:
        _check_mat(raw, "raw", csr.n_dst)
        a.raw, a.ld_raw = _ptr(raw), raw.stride(0)
    if want_dot:
        _check_mat(q, "q", csr.n_dst)
        a.q, a.ld_q, a.dot = _ptr(q), q.stride(0), _ptr(dot)
    if csr.row_perm is not None:
        a.row_perm = _ptr(csr.row_perm)
    a.algo = algo
    if init is not None:
        a.init, a.init_slabs, a.init_rows, a.init_map = _ptr(init), init.shape[0], init.shape[1], _ptr(init_map)
        if algo == _lib.ALGO_GATHER:
            a.algo = _lib.ALGO_AUTO
    lib = _lib.load()
    nbytes = lib.wsage_spmm_workspace_bytes(ctypes.byref(a))
    ws = torch.empty(nbytes, device=dev, dtype=torch.uint8) if nbytes else None
    a.workspace, a.workspace_bytes = _ptr(ws), nbytes
    rec = dict(kind="spmm", n_dst=csr.n_dst, n_src=csr.n_src, nnz=csr.nnz, dim=dim, col_bits=csr.col_bits,
               self=selfcoef is not None, init=init is not None, n_out=int(out is not None) + int(raw is not None), dot=want_dot)
    if TIMING is not None:
        rec["algo"] = int(lib.wsage_spmm_algo(ctypes.byref(a))) if csr.nnz > 0 or init is None else 0
    _timed(rec, lambda: _lib.check(lib.wsage_spmm(ctypes.byref(a), _stream()), "wsage_spmm"))
    return out, raw, dot
