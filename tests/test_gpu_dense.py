"""GPU parity of the tensor-core dense layer against an fp64 torch reference of NodeUpdate (models/gnn.py:18-25), on both
implementations: the dense16 GEMM (fp16 hi+lo operands, three products — the product's default) and the tf32x3 kernels
(wsage_split_tf32 + wsage_linear_tc + wsage_grad_w_tc).  Tolerance: 1e-4 relative (north_star); observed 1e-6 .. 3.4e-6."""
import numpy as np
import pytest
import torch

import scdeepsort_b200 as sd
from scdeepsort_b200 import dense
from scds_helpers import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_split_tf32_reconstructs_fp32():
    x = torch.randn(333, 400, device=DEV) * 3
    hi, lo, _ = dense.split_tf32(x)
    assert int((hi.view(torch.int32) & 0x1FFF).abs().max()) == 0 and int((lo.view(torch.int32) & 0x1FFF).abs().max()) == 0
    assert float((hi + lo - x).abs().max() / x.abs().max()) < 2 ** -20
    y = torch.randn(333, 400, device=DEV)
    hi, lo, masked = dense.split_tf32(x, mask_src=y, want_masked=True)
    assert torch.equal(masked, x * (y > 0))
    assert float((hi + lo - masked).abs().max()) < 2 ** -18


@pytest.mark.parametrize("m,k,n,relu,bias", [(1000, 400, 400, True, True), (128, 400, 400, False, True),
                                             (129, 32, 16, True, False), (5003, 200, 400, True, True),
                                             (777, 48, 256, False, False), (20000, 400, 208, True, True),
                                             (300, 400, 16, False, True), (1, 64, 64, True, True),
                                             (1102, 400, 200, True, True), (259, 48, 40, True, True), (700, 40, 12, False, True),
                                             (3000, 400, 800, True, True), (1500, 800, 800, True, True)])
@pytest.mark.parametrize("impl", ["dense16", "tf32x3"])
def test_linear_tc_forward_backward(m, k, n, relu, bias, impl, monkeypatch):
    monkeypatch.setattr(dense, "use_dense16", impl == "dense16")
    assert dense.tc_supported(k, n) and dense.d16_supported(k, n)
    g = torch.Generator(device="cpu").manual_seed(m + k + n)
    x = torch.randn(m, k, generator=g).to(DEV).requires_grad_(True)
    w = (torch.randn(n, k, generator=g) * 0.1).to(DEV).requires_grad_(True)
    b = torch.randn(n, generator=g).to(DEV).requires_grad_(True) if bias else None
    y = dense.linear_relu(x, w, b, relu)
    dy = torch.randn(m, n, generator=g).to(DEV)
    y.backward(dy)
    x64, w64 = x.detach().double(), w.detach().double()
    ref = x64 @ w64.t()
    if bias:
        ref = ref + b.detach().double()
    if relu:
        ref = torch.relu(ref)
    assert rel_err(y.detach().cpu(), ref.cpu()) < 1e-5          # observed 1e-6 .. 3.4e-6 (fp32 accumulation over K)
    # backward reference with OUR activation pattern: a pre-activation within 1e-6 of zero may land on
    # either side in fp32 vs fp64, which flips a whole gradient row and says nothing about the GEMM
    g64 = dy.double() * (y.detach() > 0) if relu else dy.double()
    assert rel_err(x.grad.cpu(), (g64 @ w64).cpu()) < 1e-5
    assert rel_err(w.grad.cpu(), (g64.t() @ x64).cpu()) < 1e-4
    if bias:
        assert rel_err(b.grad.cpu(), g64.sum(0).cpu()) < 1e-4


def test_unsupported_shapes_are_rejected():
    assert dense.tc_supported(400, 40) and dense.tc_supported(400, 200)    # N is padded to 16 inside the kernel
    assert not dense.tc_supported(18, 64)        # K not a multiple of 4
    assert dense.tc_supported(400, 800)          # > 512 TMEM columns: done in column blocks by the wrapper
    import ctypes
    lib = sd._lib.load()
    a = torch.zeros(8, 48, device=DEV)
    b = torch.zeros(520, 48, device=DEV)
    o = torch.zeros(8, 520, device=DEV)
    p = lambda t: ctypes.c_void_p(t.data_ptr())     # noqa: E731
    rc = lib.wsage_linear_tc(p(a), p(a), 48, p(b), p(b), 48, None, 0, p(o), 520, 8, 520, 48, None)
    assert rc == sd._lib.EINVAL and b"512" in lib.wsage_last_error()    # the raw ABI call takes N <= 512


@pytest.mark.parametrize("rows,n_out,n_in", [(5000, 400, 400), (777, 16, 400), (3001, 200, 400), (3000, 400, 200),
                                             (150000, 400, 400), (40, 8, 32), (20000, 132, 260), (15, 512, 416), (9000, 800, 800), (2000, 16, 804)])
def test_grad_w_tc_matches_fp64(rows, n_out, n_in):
    """wsage_grad_w_tc: dW = g^T x with MN-major tf32 hi/lo operands, split over the rows, vs fp64."""
    gen = torch.Generator(device="cpu").manual_seed(rows + n_out + n_in)
    g = torch.randn(rows, n_out, generator=gen).to(DEV)
    x = torch.randn(rows, n_in, generator=gen).to(DEV)
    g_hi, g_lo, _ = dense.split_tf32(g)
    x_hi, x_lo, _ = dense.split_tf32(x)
    assert dense.grad_w_supported(rows, n_out, n_in)
    dw = dense.grad_w_tc(g_hi, g_lo, x_hi, x_lo)
    ref = g.double().t() @ x.double()
    assert rel_err(dw.cpu(), ref.cpu()) < 1.5e-5      # chains of <= 1024 rows: ~7e-6 from the truncating accumulator
    assert torch.equal(dw, dense.grad_w_tc(g_hi, g_lo, x_hi, x_lo))          # split partials added in fixed order


def test_single_product_mode_is_tf32_grade():
    """dense.single_product (the bf16 configuration of bench.py --config c3): lo operands dropped, one tf32 product per
    k-step — stated tolerance 2e-3 (tf32 keeps 11 significand bits; observed ~3e-4), forward and both gradients."""
    gen = torch.Generator(device="cpu").manual_seed(0)
    x = torch.randn(3000, 400, generator=gen).to(DEV).requires_grad_(True)
    w = (torch.randn(200, 400, generator=gen) * 0.05).to(DEV).requires_grad_(True)
    b = torch.randn(200, generator=gen).to(DEV).requires_grad_(True)
    try:
        dense.single_product = True
        y = dense.linear_relu(x, w, b, relu=True)
        y.square().sum().backward()
    finally:
        dense.single_product = False
    x64, w64, b64 = (t.detach().double().requires_grad_(True) for t in (x, w, b))
    y64 = torch.relu(x64 @ w64.t() + b64)
    y64.square().sum().backward()
    assert 1e-6 < rel_err(y.detach().cpu(), y64.detach().cpu()) < 2e-3
    assert rel_err(x.grad.cpu(), x64.grad.cpu()) < 5e-3 and rel_err(w.grad.cpu(), w64.grad.cpu()) < 5e-3


def test_colsum_masked_matches_torch_and_is_deterministic():
    g = torch.Generator(device="cpu").manual_seed(1)
    for rows, cols in ((70000, 400), (5, 16), (1234, 800), (129, 1024)):
        x = torch.randn(rows, cols, generator=g).to(DEV)
        y = torch.randn(rows, cols, generator=g).to(DEV)
        got = dense.colsum_masked(x, y)
        assert rel_err(got.cpu(), (x.double() * (y > 0)).sum(0).cpu()) < 1e-5
        assert torch.equal(got, dense.colsum_masked(x, y))
        assert rel_err(dense.colsum_masked(x).cpu(), x.double().sum(0).cpu()) < 1e-5


@pytest.mark.parametrize("rows,cols", [(70000, 400), (5, 16), (1234, 132), (129, 512), (300, 36)])
def test_blocked_split_with_fused_column_sums(rows, cols):
    """wsage_split16_colsum: the planes are those of wsage_split16_masked (bit for bit, in every block that holds columns), decode
    to the masked matrix within the fp16 hi+lo resolution, and the fused bias gradient equals the two-pass one's value."""
    g = torch.Generator(device="cpu").manual_seed(rows + cols)
    x = (torch.randn(rows, cols, generator=g) * 3).to(DEV)
    y = torch.randn(rows, cols, generator=g).to(DEV)
    fmt = sd._lib.D16_F16X2
    db = torch.full((cols,), float("nan"), device=DEV)
    hi, lo, amax, _, _ = dense._planes_a(x, fmt, mask_src=y, colsum=db)
    hi2, lo2, amax2, _, _ = dense._planes_a(x, fmt, mask_src=y)
    pad = int(sd._lib.load().wsage_dense16_slots_pad(cols))
    nb_used = (cols + 31) // 32
    view = lambda p: p.view(-1, pad // 32, 128, 32)[:, :nb_used]
    assert torch.equal(view(hi), view(hi2)) and torch.equal(view(lo), view(lo2)) and torch.equal(amax, amax2)
    want = x.double() * (y > 0)
    assert rel_err(db.cpu(), want.sum(0).cpu()) < 1e-5
    db2 = torch.empty(cols, device=DEV)
    dense._planes_a(x, fmt, mask_src=y, colsum=db2)
    assert torch.equal(db, db2)                                            # fixed assignment and order
    # decode: [tile][block][row][col] -> [rows, cols]
    k = 14 - int(np.frexp(float(amax))[1])
    dec = (view(hi).view(torch.float16).double() + view(lo).view(torch.float16).double()) * 2.0 ** -k
    dec = dec.permute(0, 2, 1, 3).reshape(-1, nb_used * 32)
    assert float((dec[:rows, :cols] - want).abs().max()) <= 2.0 ** -20 * float(amax)
    assert float(dec[rows:].abs().max() if dec.shape[0] > rows else 0.0) == 0.0 and float(dec[:, cols:].abs().sum()) == 0.0
