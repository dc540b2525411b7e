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
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


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
        _lib.check(lib.wsage_block_agg_fwd(_ptr(block.rowptr), _ptr(block.col), _ptr(block.weight),
                                           _ptr(src_id), _ptr(dst_id), _ptr(alpha_flat), gene_num,
                                           _ptr(h), h.stride(0), block.n_src,
                                           _ptr(out), out.stride(0), block.n_dst, dim, _stream()),
                   "wsage_block_agg_fwd")
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
            _lib.check(lib.wsage_block_agg_bwd(_ptr(block.rowptr), _ptr(block.col), _ptr(block.weight),
                                               _ptr(src_id), _ptr(dst_id), _ptr(alpha_flat), ctx.gene_num,
                                               _ptr(h), h.stride(0), block.n_src,
                                               _ptr(d_out), d_out.stride(0), block.n_dst, h.shape[1],
                                               _ptr(dh), dh.stride(0) if need_h else 0, _ptr(da), _stream()),
                       "wsage_block_agg_bwd")
        return dh, (da.reshape(ctx.alpha_shape) if need_a else None), None, None, None, None


def block_aggregate(h, alpha, block: Block, src_id, dst_id, gene_num: int):
    """neigh[v] = mean_{e→v} w_e · α[k(e)] · h[src_e]  for one NodeFlow block."""
    return _BlockAggregate.apply(h, alpha, block, src_id, dst_id, gene_num)


@dataclass
class DenseBlock:
    """Entries of the popular genes, taken out of a ``Csr`` and stored zero-filled and tile-blocked
    (``x[tile][k][T]``, T = ``wsage_dense_tile()``; see include/wsage.h and csrc/agg_dense.cuh)."""
    x: torch.Tensor                              # fp32 [n_tiles * k * T]
    k: int                                       # sources of the block
    t: int                                       # destination slots of the block
    src_ids: Optional[torch.Tensor] = None       # int32 [k] rows of hs (None: source k = row k)
    dst_map: Optional[torch.Tensor] = None       # int32 [n_dst] destination row -> slot, -1 = not in the block
    nnz: int = 0                                 # expression entries the block stands for


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
    dense: Optional[DenseBlock] = None        # entries removed from col/x and handled by the dense-block kernel

    @property
    def nnz(self):
        return int(self.x.shape[0])


def spmm(csr: Csr, hs, *, dscale=None, selfcoef=None, hself=None, out=None, want_out=True,
         raw=None, want_raw=False, q=None, want_dot=False, algo=_lib.ALGO_AUTO):
    """acc = Σ_e x_e·hs[col_e];  out = dscale·acc + selfcoef·hself;  raw = acc;  dot = <acc, q>.

    Returns (out, raw, dot) with None for the ones not requested."""
    _check_mat(hs, "hs", csr.n_src)
    dim = hs.shape[1]
    dev = hs.device
    if want_out and out is None:
        out = torch.empty(csr.n_dst, dim, device=dev, dtype=torch.float32)
    if want_raw and raw is None:
        raw = torch.empty(csr.n_dst, dim, device=dev, dtype=torch.float32)
    dot = torch.empty(csr.n_dst, device=dev, dtype=torch.float32) if want_dot else None
    a = _lib.SpmmArgs()
    a.rowptr, a.col, a.col_bits, a.x, a.nnz = _ptr(csr.rowptr), _ptr(csr.col), csr.col_bits, _ptr(csr.x), csr.nnz
    a.hs, a.ld_hs, a.n_src, a.n_dst, a.dim = _ptr(hs), hs.stride(0), csr.n_src, csr.n_dst, dim
    if dscale is not None:
        _check_vec(dscale, "dscale", torch.float32, csr.n_dst)
        a.dscale = _ptr(dscale)
    if selfcoef is not None:
        _check_vec(selfcoef, "selfcoef", torch.float32, csr.n_dst)
        _check_mat(hself, "hself", csr.n_dst)
        a.selfcoef, a.hself, a.ld_hself = _ptr(selfcoef), _ptr(hself), hself.stride(0)
    if out is not None:
        _check_mat(out, "out", csr.n_dst)
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
    if csr.dense is not None:
        d = csr.dense
        a.dense_x, a.dense_k, a.dense_t = _ptr(d.x), d.k, d.t
        a.dense_src_ids, a.dense_dst_map = _ptr(d.src_ids), _ptr(d.dst_map)
    lib = _lib.load()
    nbytes = lib.wsage_spmm_workspace_bytes(ctypes.byref(a))
    ws = torch.empty(nbytes, device=dev, dtype=torch.uint8) if nbytes else None
    a.workspace, a.workspace_bytes = _ptr(ws), nbytes
    if TIMING is None:
        _lib.check(lib.wsage_spmm(ctypes.byref(a), _stream()), "wsage_spmm")
    else:       # bench.py: per-launch CUDA events on the launching stream
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.wsage_spmm(ctypes.byref(a), _stream()), "wsage_spmm")
        e1.record()
        TIMING.append(dict(algo=int(lib.wsage_spmm_algo(ctypes.byref(a))), n_dst=csr.n_dst, n_src=csr.n_src,
                           nnz=csr.nnz, dim=dim, col_bits=csr.col_bits, self=selfcoef is not None,
                           dense_nnz=csr.dense.nnz if csr.dense is not None else 0,
                           dense_pairs=csr.dense.k * csr.dense.t if csr.dense is not None else 0,
                           n_out=int(out is not None) + int(raw is not None), dot=want_dot, events=(e0, e1)))
    return out, raw, dot
