"""Tensor-core dense layer: PyTorch-facing wrappers of ``wsage_split_tf32`` / ``wsage_linear_tc`` / ``wsage_grad_w_tc``.

``linear_relu(x, weight, bias, relu)`` computes ``act(x @ weight.T + bias)`` (NodeUpdate,
/root/reference/models/gnn.py:18-25) with the tcgen05 kernel: operands split into tf32 hi+lo,
three MMAs per k-step accumulated in fp32 (error ~3e-6 relative: fp32-grade, far inside the 1e-4 parity bar).
Backward: the input gradient goes through the same kernel (B = weight transposed), the ReLU mask is
fused into the split kernel, the weight gradient (a [N, K] reduction over all rows) runs on ``wsage_grad_w_tc``;
only the bias gradient (a column sum) is left to torch.  ``single_product = True`` (set by the bf16 configuration)
drops the lo operands: one tf32 product per k-step.
"""
import ctypes

import torch

from . import _lib
from .ops import _ptr, _stream


single_product = False     # True: plain tf32 (hi operands only) — BASELINE configs[2], not the fp32 parity path


def split_tf32(x: torch.Tensor, mask_src: torch.Tensor = None, want_masked=False):
    """(hi, lo, masked): hi = rn_tf32(v), lo = rn_tf32(v - hi) as fp32 [rows, cols]; v = x or x * (mask_src > 0).
    lo is None under ``single_product``."""
    if not (x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1 and x.shape[1] % 4 == 0):
        raise ValueError(f"split_tf32: expected a CUDA fp32 row-major matrix with cols % 4 == 0, got {tuple(x.shape)} {x.dtype}")
    rows, cols = x.shape
    hi = torch.empty(rows, cols, device=x.device, dtype=torch.float32)
    lo = None if single_product else torch.empty(rows, cols, device=x.device, dtype=torch.float32)
    masked = torch.empty(rows, cols, device=x.device, dtype=torch.float32) if (want_masked and mask_src is not None) else None
    lib = _lib.load()
    _lib.check(lib.wsage_split_tf32(_ptr(x), x.stride(0), _ptr(mask_src), mask_src.stride(0) if mask_src is not None else 0,
                                    _ptr(hi), _ptr(lo), cols, _ptr(masked), cols if masked is not None else 0,
                                    rows, cols, _stream()), "wsage_split_tf32")
    return hi, lo, masked


_MAX_N = 512      # TMEM columns of one CTA: wider outputs are produced in column blocks
use_tc_grad_w = True     # weight gradient on the tensor cores (wsage_grad_w_tc); False: torch (cuBLAS fp32)


def linear_tc(a_hi, a_lo, b_hi, b_lo, m, n, k, bias=None, relu=False):
    out = torch.empty(m, n, device=a_hi.device, dtype=torch.float32)
    lib = _lib.load()
    blocks = (n + _MAX_N - 1) // _MAX_N
    step = -(-n // blocks)
    step += -step % 16                       # equal-ish column blocks, multiples of 16
    for n0 in range(0, n, step):
        nc = min(step, n - n0)
        o = out[:, n0:n0 + nc]
        _lib.check(lib.wsage_linear_tc(_ptr(a_hi), _ptr(a_lo), a_hi.stride(0), _ptr(b_hi[n0:]), _ptr(b_lo[n0:]) if b_lo is not None else None, b_hi.stride(0),
                                       _ptr(bias[n0:]) if bias is not None else None, 1 if relu else 0,
                                       _ptr(o), out.stride(0), m, nc, k, _stream()), "wsage_linear_tc")
    return out


_GW_MAX_IN = 400       # input columns per wsage_grad_w_tc call (thirteen 32-column blocks fill its 3-stage shared-memory ring)


def grad_w_tc(g_hi, g_lo, x_hi, x_lo):
    """dW[n_out, n_in] = g^T x on the tensor cores (MN-major tf32 hi/lo operands, reduction over the rows).  Inputs wider
    than 400 columns (hidden 800 of BASELINE configs[4]) are produced in column blocks of x — strided views, no copies."""
    rows, n_out = g_hi.shape
    n_in = x_hi.shape[1]
    lib = _lib.load()
    splits = int(lib.wsage_grad_w_splits(rows, n_out))
    out = torch.empty(n_out, n_in, device=g_hi.device, dtype=torch.float32)
    blocks = -(-n_in // _GW_MAX_IN)
    step = -(-n_in // blocks)
    step += -step % 4
    partial = torch.empty(splits, n_out, min(step, n_in), device=g_hi.device, dtype=torch.float32)
    for c0 in range(0, n_in, step):
        nc = min(step, n_in - c0)
        _lib.check(lib.wsage_grad_w_tc(_ptr(g_hi), _ptr(g_lo), g_hi.stride(0), _ptr(x_hi[:, c0:]), _ptr(x_lo[:, c0:]) if x_lo is not None else None,
                                       x_hi.stride(0), rows, n_out, nc, _ptr(partial), splits, _ptr(out[:, c0:]), out.stride(0), _stream()),
                   "wsage_grad_w_tc")
    return out


def grad_w_supported(rows: int, n_out: int, n_in: int) -> bool:
    return rows > 0 and n_out % 4 == 0 and n_in % 4 == 0 and n_in > 0 and n_out > 0


def tc_supported(in_features: int, out_features: int) -> bool:
    """Shapes the tensor-core kernel takes in both directions (forward N = out, input-gradient N = in):
    widths that are multiples of 4 (16-byte fp32 rows for the split kernel); outputs wider than 512 (TMEM
    columns) are produced in column blocks."""
    return in_features % 4 == 0 and out_features % 4 == 0 and in_features > 0 and out_features > 0


class _LinearReluTC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, relu):
        x = x.contiguous()
        m, k = x.shape
        n = weight.shape[0]
        x_hi, x_lo, _ = split_tf32(x)
        w_hi, w_lo, _ = split_tf32(weight.contiguous())
        y = linear_tc(x_hi, x_lo, w_hi, w_lo, m, n, k, bias=bias, relu=relu)
        # the weight gradient consumes the split input again (as the MN-major operand of g^T x)
        ctx.tc_dw = use_tc_grad_w and grad_w_supported(m, n, k)
        if ctx.tc_dw:
            ctx.save_for_backward(x_hi, x_lo, weight, y if relu else None)
        else:
            ctx.save_for_backward(x, None, weight, y if relu else None)
        ctx.relu = relu
        return y

    @staticmethod
    def backward(ctx, dy):
        x, x_lo, weight, y = ctx.saved_tensors          # x is x_hi when ctx.tc_dw
        need_x, need_w, need_b = ctx.needs_input_grad[:3]
        dy = dy.contiguous()
        m, n = dy.shape
        k = x.shape[1]
        # g = dy * (y > 0): hi/lo for the tensor-core input-gradient GEMM, fp32 copy for dW / db
        g_hi, g_lo, g = split_tf32(dy, mask_src=y if ctx.relu else None, want_masked=ctx.relu)
        if g is None:
            g = dy
        dx = dw = db = None
        if need_x:
            wt_hi, wt_lo, _ = split_tf32(weight.t().contiguous())          # B = W^T : [K, N], reduction over N
            dx = linear_tc(g_hi, g_lo, wt_hi, wt_lo, m, k, n)
        if need_w:
            dw = grad_w_tc(g_hi, g_lo, x, x_lo) if ctx.tc_dw else g.t() @ x
        if need_b:
            db = colsum_masked(g)
        return dx, dw, db, None


# ---------------------------------------------------------------------------------------------------------------------
# The same layer on the dense16 GEMM (csrc/dense16.cuh): fp16 hi+lo operands, three products — the precision grade of the
# tf32x3 kernels above at twice the MMA rate and half the operand bytes; bf16 (one product) under ``single_product``.
#   forward   out = act(x W^T + b)   side 0: A = BLOCKED(x),      B = COLBLOCKS(W)  (k = in-feature  = column of W)
#   dx        = g W                   side 0: A = BLOCKED(g),      B = KBLOCKS(W)    (k = out-feature = row of W)
#   dW        = g^T x                 side 1: A = BLOCKED(g) (its columns are the destination slots), B = KBLOCKS(x)
# ---------------------------------------------------------------------------------------------------------------------
use_dense16 = True          # False: the tf32x3 kernels (wsage_linear_tc / wsage_grad_w_tc)
_D16_MAX_N = 512


def _fmt():
    return _lib.D16_BF16 if single_product else _lib.D16_F16X2


def _amax(x, fmt):
    if fmt != _lib.D16_F16X2:
        return None
    amax = torch.zeros(1, device=x.device, dtype=torch.float32)
    _lib.check(_lib.load().wsage_amax(_ptr(x), x.stride(0), None, None, x.shape[0], x.shape[1], _ptr(amax), _stream()), "wsage_amax")
    return amax


def _split(x, fmt, layout, ld, shape, amax, mask_src=None):
    hi = torch.empty(shape, device=x.device, dtype=torch.int16)
    lo = torch.empty(shape, device=x.device, dtype=torch.int16) if fmt == _lib.D16_F16X2 else None
    _lib.check(_lib.load().wsage_split16_masked(_ptr(x), x.stride(0), _ptr(mask_src), mask_src.stride(0) if mask_src is not None else 0,
                                                None, None, x.shape[0], x.shape[1], _ptr(amax), fmt, layout, _ptr(hi), _ptr(lo), ld,
                                                _stream()), "wsage_split16")
    return hi, lo


def _planes_a(x, fmt, mask_src=None, colsum=None):
    """BLOCKED planes of a [rows, k] matrix: (hi, lo, amax, rows, k).  ``colsum`` ([k] floats): also receives the column sums
    of x * (mask_src > 0) — the bias gradient — from the same pass."""
    rows, k = x.shape
    lib = _lib.load()
    pad = int(lib.wsage_dense16_slots_pad(k))
    amax = _amax(x, fmt)
    shape = ((rows + 127) // 128 * pad * 128,)
    if colsum is None:
        hi, lo = _split(x, fmt, _lib.SPLIT_BLOCKED, pad, shape, amax, mask_src)
    else:
        hi = torch.empty(shape, device=x.device, dtype=torch.int16)
        lo = torch.empty(shape, device=x.device, dtype=torch.int16) if fmt == _lib.D16_F16X2 else None
        n_partial = max(1, min(148 * 8 // ((k + 31) // 32), (rows + 127) // 128))
        partial = torch.empty(n_partial, k, device=x.device, dtype=torch.float32)
        _lib.check(lib.wsage_split16_colsum(_ptr(x), x.stride(0), _ptr(mask_src), mask_src.stride(0) if mask_src is not None else 0,
                                            rows, k, _ptr(amax), fmt, _ptr(hi), _ptr(lo), pad, _ptr(partial), n_partial, _ptr(colsum),
                                            _stream()), "wsage_split16_colsum")
    return hi, lo, amax, rows, k


def _planes_b(x, fmt, k_is_row, amax=None):
    """B planes [ceil(K / 32)][ceil16(N)][32] of a matrix whose k index is its row (KBLOCKS) or its column (COLBLOCKS).
    ``amax``: a device scalar already known to bound |x| (any valid bound gives a valid scale)."""
    rows, cols = x.shape
    if amax is None:
        amax = _amax(x, fmt)
    if k_is_row:
        ld = (cols + 15) // 16 * 16
        hi, lo = _split(x, fmt, _lib.SPLIT_KBLOCKS, ld, ((rows + 31) // 32, ld, 32), amax)
    else:
        ld = (rows + 15) // 16 * 16
        hi, lo = _split(x, fmt, _lib.SPLIT_COLBLOCKS, ld, ((cols + 31) // 32, ld, 32), amax)
    return hi, lo, amax, ld


def _gemm16(a, b, fmt, side, dim, out, *, bias=None, relu=False, n_splits_out=None):
    """One wsage_dense16 call with activation planes as the X operand."""
    a_hi, a_lo, a_amax, rows, k = a
    b_hi, b_lo, b_amax, ld = b
    lib = _lib.load()
    args = _lib.Dense16Args()
    args.x_hi, args.x_lo, args.fmt, args.cells, args.gene_slots, args.x_scale = _ptr(a_hi), _ptr(a_lo), fmt, rows, k, 1.0
    args.x_amax, args.side, args.dim = _ptr(a_amax), side, dim
    args.h_hi, args.h_lo, args.ld_h, args.h_amax = _ptr(b_hi), _ptr(b_lo), ld, _ptr(b_amax)
    if side == 0:
        args.n_dst, args.bias, args.relu = rows, _ptr(bias), 1 if relu else 0
        args.out, args.ld_out = _ptr(out), out.stride(0)
    else:
        args.n_src_cells = rows
        n_splits = int(lib.wsage_dense16_splits(ctypes.byref(args)))
        if n_splits <= 0:
            _lib.check(_lib.EINVAL, "wsage_dense16_splits")
        pad = int(lib.wsage_dense16_slots_pad(k))
        out = torch.empty(n_splits, pad, dim, device=a_hi.device, dtype=torch.float32)
        args.out, args.ld_out = _ptr(out), dim
    _lib.check(lib.wsage_dense16(ctypes.byref(args), _stream()), "wsage_dense16")
    return out


def colsum_masked(x, mask_src=None):
    """Column sums of x * (mask_src > 0): the bias gradient (deterministic two-stage sum on the device)."""
    rows, cols = x.shape
    if cols % 4 or cols > 1024:
        return (x if mask_src is None else x * (mask_src > 0)).sum(dim=0)
    n_partial = max(1, min(148 * 4, (rows + 63) // 64))
    partial = torch.empty(n_partial, cols, device=x.device, dtype=torch.float32)
    out = torch.empty(cols, device=x.device, dtype=torch.float32)
    _lib.check(_lib.load().wsage_colsum_masked(_ptr(x), x.stride(0), _ptr(mask_src), mask_src.stride(0) if mask_src is not None else 0,
                                               rows, cols, _ptr(partial), n_partial, _ptr(out), _stream()), "wsage_colsum_masked")
    return out


class _LinearReluD16(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, relu):
        x = x.contiguous()
        w = weight.contiguous()
        m, k = x.shape
        n = w.shape[0]
        fmt = _fmt()
        a = _planes_a(x, fmt)
        y = torch.empty(m, n, device=x.device, dtype=torch.float32)
        for n0 in range(0, n, _D16_MAX_N):
            n1 = min(n, n0 + _D16_MAX_N)
            _gemm16(a, _planes_b(w[n0:n1], fmt, k_is_row=False), fmt, 0, n1 - n0, y[:, n0:n1],
                    bias=bias[n0:n1].contiguous() if bias is not None else None, relu=relu)
        ctx.save_for_backward(x, w, y if relu else None, a[2])
        ctx.relu, ctx.fmt = relu, fmt
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, y, x_amax = ctx.saved_tensors
        need_x, need_w, need_b = ctx.needs_input_grad[:3]
        dy = dy.contiguous()
        m, n = dy.shape
        k = x.shape[1]
        fmt = ctx.fmt
        dx = dw = db = None
        if need_x or need_w:
            if need_b and m > 0:        # the bias gradient rides on the split's pass over dy and the mask
                db = torch.empty(n, device=dy.device, dtype=torch.float32)
            g = _planes_a(dy, fmt, mask_src=y if ctx.relu else None, colsum=db)  # g = dy * (y > 0), split into A planes
        if need_x:
            dx = torch.empty(m, k, device=dy.device, dtype=torch.float32)
            for k0 in range(0, k, _D16_MAX_N):
                k1 = min(k, k0 + _D16_MAX_N)
                _gemm16(g, _planes_b(w[:, k0:k1].contiguous(), fmt, k_is_row=True), fmt, 0, k1 - k0, dx[:, k0:k1])
        if need_w:
            dw = torch.empty(n, k, device=dy.device, dtype=torch.float32)
            lib = _lib.load()
            for k0 in range(0, k, _D16_MAX_N):
                k1 = min(k, k0 + _D16_MAX_N)
                xs = x if (k0 == 0 and k1 == k) else x[:, k0:k1].contiguous()
                slabs = _gemm16(g, _planes_b(xs, fmt, k_is_row=True, amax=x_amax), fmt, 1, k1 - k0, None)
                _lib.check(lib.wsage_sum_slabs(_ptr(slabs), slabs.shape[0], slabs.shape[1] * slabs.shape[2], n, k1 - k0,
                                               _ptr(dw[:, k0:k1]), dw.stride(0), _stream()), "wsage_sum_slabs")
        if need_b and db is None:
            db = colsum_masked(dy, y if ctx.relu else None)
        return dx, dw, db, None


def d16_supported(in_features: int, out_features: int) -> bool:
    return in_features % 4 == 0 and out_features % 4 == 0 and in_features > 0 and out_features > 0


def linear_relu(x, weight, bias=None, relu=True):
    if use_dense16 and d16_supported(weight.shape[1], weight.shape[0]):
        return _LinearReluD16.apply(x, weight, bias, relu)
    return _LinearReluTC.apply(x, weight, bias, relu)
