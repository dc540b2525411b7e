"""Dense-block split  X = X_sparse + X_dense  (csrc/dense16.cuh, BipartiteGraph.densify): the popular genes'
entries leave the CSRs and run on the tensor cores (tcgen05, fp16 hi + lo planes, three products, chained
accumulation with TMA reduce-adds); results must match the plain CSR path (fp32 summation order aside), the fp64
reference and the golden logits / oracle gradients.  Tolerances are written where they are used."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

import scdeepsort_b200 as sd
from oracle import gnn_oracle, graph_oracle
from scdeepsort_b200 import _lib, ops
from scdeepsort_b200.synthetic import synthetic_bipartite
from scds_helpers import csr_dense_matrix, dense_block_matrix, golden_csr, golden_graph, golden_state, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-4


def _skewed_expression(n_cells, n_genes, seed, avg=0.08):
    rng = np.random.RandomState(seed)
    pop = np.minimum(1.0, avg * n_genes ** 0.8 / 5 * (np.arange(1, n_genes + 1)) ** -0.8)[rng.permutation(n_genes)]
    m = rng.rand(n_cells, n_genes) < pop[None, :]
    return sp.csr_matrix(np.where(m, rng.uniform(0.05, 9, (n_cells, n_genes)), 0).astype(np.float32))


@pytest.mark.parametrize("n_cells,n_genes,dim,chunk_rows,fmt", [
    (1000, 700, 400, 0, "f16x2"), (333, 257, 400, 64, "f16x2"), (5000, 300, 200, 512, "f16x2"), (700, 900, 128, 0, "f16x2"),
    (640, 500, 132, 96, "f16x2"), (150, 90, 512, 32, "f16x2"), (129, 33, 64, 0, "f16x2"), (40000, 160, 400, 0, "f16x2"),
    (1000, 700, 400, 0, "bf16"), (333, 257, 200, 64, "bf16"),
    # 76 / 90 pairs of tiles over 74 SM pairs: the last round (2 / 16 pairs) is cut along k into 4 / 3 pieces that reduce-add
    (19300, 600, 64, 128, "f16x2"), (22900, 300, 48, 64, "f16x2")])
def test_dense16_kernel_both_sides_vs_fp64(n_cells, n_genes, dim, chunk_rows, fmt):
    """wsage_dense16 alone: every gene in the block, both sides, ragged tiles, chains of ``chunk_rows`` rows (several
    TMA reduce-adds per tile), against the fp64 product of the DECODED planes.  fp16x2: ≤ 1e-5 of the output's largest
    entry (observed ~1e-6); bf16: the planes are exact inputs, only H is rounded to 8 bits: ≤ 1e-2."""
    x = _skewed_expression(n_cells, n_genes, n_cells + dim, avg=0.3)
    bg = sd.BipartiteGraph.from_expression(x, device=DEV).densify(0.0, fmt=fmt)
    d = bg.cell_csr.dense
    assert bg.cell_csr.nnz == 0 and d.gene_slots == int((np.diff(x.tocsc().indptr) > 0).sum())
    xd = dense_block_matrix(d).double()                                    # [cells, slots]
    tol = 1e-5 if fmt == "f16x2" else 1e-2
    g = torch.Generator(device=DEV).manual_seed(3)
    ids = d.gene_ids.cpu().long()
    # side 0: destinations = cells, with and without the fused epilogue
    hg = torch.randn(n_genes, dim, device=DEV, generator=g) * 3.0
    hself = torch.randn(n_cells, dim, device=DEV, generator=g)
    dscale = torch.rand(n_cells, device=DEV, generator=g) + 0.5
    selfcoef = torch.rand(n_cells, device=DEV, generator=g)
    acc = xd @ hg.double().cpu()[ids]
    out = ops.dense16(d, 0, hg, n_dst=n_cells, chunk_rows=chunk_rows)
    assert rel_err(out.cpu(), acc) < tol
    out2 = ops.dense16(d, 0, hg, n_dst=n_cells, dscale=dscale, selfcoef=selfcoef, hself=hself, chunk_rows=chunk_rows)
    ref2 = dscale.double().cpu()[:, None] * acc + selfcoef.double().cpu()[:, None] * hself.double().cpu()
    assert rel_err(out2.cpu(), ref2) < tol
    n_sub = max(1, n_cells - 70)                                           # fewer destinations than the block covers
    out3 = ops.dense16(d, 0, hg, n_dst=n_sub, chunk_rows=chunk_rows)
    assert out3.shape[0] == n_sub and rel_err(out3.cpu(), acc[:n_sub]) < tol
    det = ops.dense16(d, 0, hg, n_dst=n_cells, chunk_rows=chunk_rows, deterministic=True)      # every tile whole
    assert torch.equal(det, ops.dense16(d, 0, hg, n_dst=n_cells, chunk_rows=chunk_rows, deterministic=True))       # bitwise reproducible
    assert rel_err(det.cpu(), acc) < tol and float((det - out).abs().max()) <= 2e-6 * float(det.abs().max())
    det2 = ops.dense16(d, 0, hg, n_dst=n_cells, dscale=dscale, selfcoef=selfcoef, hself=hself, chunk_rows=chunk_rows, deterministic=True)
    assert float((det2 - out2).abs().max()) <= 2e-6 * float(det2.abs().max())
    # side 1: destinations = gene slots, partial slabs summed in order
    hc = torch.randn(n_cells, dim, device=DEV, generator=g) * 1e-3         # small values: exercises the dynamic scale
    part = ops.dense16(d, 1, hc, n_src_cells=n_cells, chunk_rows=chunk_rows)
    assert part.shape[1] == d.slots_pad and part.shape[2] == dim
    got = part.sum(0)[:d.gene_slots]
    ref = xd.t() @ hc.double().cpu()
    assert rel_err(got.cpu(), ref) < tol
    assert float(part[:, d.gene_slots:].abs().max()) == 0.0 if d.slots_pad > d.gene_slots else True
    n_sup = max(1, n_cells - 45)                                           # only the first cells send (support cells)
    part2 = ops.dense16(d, 1, hc, n_src_cells=n_sup, chunk_rows=chunk_rows)
    assert rel_err(part2.sum(0)[:d.gene_slots].cpu(), xd[:n_sup].t() @ hc.double().cpu()[:n_sup]) < tol


@pytest.mark.parametrize("n_cells,n_genes,dim,thr", [(1000, 700, 400, 0.2), (333, 257, 400, 0.05), (5000, 300, 200, 0.3),
                                                      (700, 900, 128, 0.2), (640, 500, 132, 0.2), (150, 90, 512, 0.5),
                                                      (70000, 120, 400, 0.2), (900, 400, 400, 0.0)])
def test_spmm_dense_block_matches_plain_csr_and_fp64(n_cells, n_genes, dim, thr):
    x = _skewed_expression(n_cells, n_genes, n_cells + dim)
    plain = sd.BipartiteGraph.from_expression(x, device=DEV)
    split = sd.BipartiteGraph.from_expression(x, device=DEV).densify(thr)
    assert split.densified and split.cell_csr.dense is split.gene_csr.dense is not None
    assert split.cell_csr.nnz + split.cell_csr.dense.nnz == plain.cell_csr.nnz
    assert (split.cell_csr.nnz == 0) == (thr == 0.0)
    g = torch.Generator(device=DEV).manual_seed(7)
    # the block's planes carry 22 bits of every entry: the reference uses what the planes hold
    xd = {"cell_csr": None, "gene_csr": None}
    for which in xd:
        csr = getattr(split, which)
        col = csr.col.to(torch.int64).cpu()
        if csr.col_bits == 16:
            col = col & 0xFFFF
        rest = sp.csr_matrix((csr.x.cpu().numpy(), col.numpy(), csr.rowptr.cpu().numpy()), shape=(csr.n_dst, csr.n_src)).toarray()
        xd[which] = torch.from_numpy(rest).double() + csr_dense_matrix(csr).double()
    assert float((xd["cell_csr"] - torch.from_numpy(x.toarray()).double()).abs().max()) < 9 * 2.0 ** -21
    for which in ("cell_csr", "gene_csr"):
        a, b = getattr(plain, which), getattr(split, which)
        hs = torch.randn(a.n_src, dim, device=DEV, generator=g)
        hself = torch.randn(a.n_dst, dim, device=DEV, generator=g)
        q = torch.randn(a.n_dst, dim, device=DEV, generator=g)
        dscale = torch.rand(a.n_dst, device=DEV, generator=g) + 0.5
        selfcoef = torch.rand(a.n_dst, device=DEV, generator=g)
        kw = dict(dscale=dscale, selfcoef=selfcoef, hself=hself, want_raw=True, q=q, want_dot=True)
        o1, r1, d1 = sd.spmm(a, hs, **kw)
        o2, r2, d2 = sd.spmm(b, hs, **kw)
        acc = xd[which] @ hs.double().cpu()
        ref_out = dscale.double().cpu()[:, None] * acc + selfcoef.double().cpu()[:, None] * hself.double().cpu()
        assert rel_err(r2.cpu(), acc) < 1e-5
        assert rel_err(o2.cpu(), ref_out) < 1e-5
        assert rel_err(d2.cpu(), (acc * q.double().cpu()).sum(1)) < 1e-5
        assert rel_err(o2.cpu(), o1.cpu()) < 2e-5 and rel_err(r2.cpu(), r1.cpu()) < 2e-5       # each is within 1e-5 of fp64
        assert torch.equal(o2, sd.spmm(b, hs, **kw)[0])                  # deterministic
        o3 = sd.spmm(b, hs, dscale=dscale, selfcoef=selfcoef, hself=hself)[0]       # side 0 + empty CSR: fused epilogue
        assert rel_err(o3.cpu(), ref_out) < 1e-5


@pytest.mark.parametrize("n_layers,thr", [(1, 0.2), (2, 0.2), (3, 0.2), (2, 0.0)])
def test_full_graph_fwd_bwd_with_dense_block_vs_oracle(golden_train, n_layers, thr):
    """Golden Muscle sub-sample: reference logits (L <= 2) and fp64 oracle gradients, popular genes (thr 0: every
    gene) on the tensor cores."""
    z = golden_train
    gg = golden_graph(z)
    if n_layers <= 2:
        params = golden_state(z, f"L{n_layers}")
    else:
        params = gnn_oracle.init_params(int(z["dense_dim"]), int(z["hidden"]), int(z["num_labels"]), n_layers,
                                        gg.num_genes, perturb_alpha=True)
    bg = sd.BipartiteGraph.from_expression(golden_csr(z), device=DEV).densify(thr)
    assert bg.densified and len(bg.dense_genes) > 0
    d_in, hidden, k = int(z["dense_dim"]), int(z["hidden"]), int(z["num_labels"])
    model = sd.GNN(d_in, hidden, k, n_layers, gg.num_genes, activation=torch.relu).to(DEV)
    model.load_state_dict(params)
    model.train()
    seeds = torch.arange(gg.num_genes, gg.num_nodes)
    labels = torch.from_numpy(z["labels"])[seeds]
    logits = model(sd.FullGraphFlow(bg, gg.features.to(DEV)))
    loss = torch.nn.functional.cross_entropy(logits, labels.to(DEV), reduction="sum")
    loss.backward()
    if n_layers <= 2:
        assert rel_err(logits.detach().cpu(), z[f"L{n_layers}/logits"]) < 1e-5
    flow = graph_oracle.full_neighbor_flow(gg, seeds, n_layers)
    loss_ref, logits_ref, grads_ref = gnn_oracle.loss_and_grads(params, flow, labels, gg.num_genes, dtype=torch.float64)
    assert rel_err(logits.detach().cpu(), logits_ref) < 1e-5
    assert abs(float(loss) - float(loss_ref)) < 1e-5 * float(loss_ref)
    for name, v in grads_ref.items():
        assert rel_err(dict(model.named_parameters())[name].grad.cpu(), v) < TOL, name


def test_dense_block_c3_agrees_with_plain_and_checksum():
    """BASELINE configs[2] shape (100k x 20k, 2e8 edges): densified vs plain kernels, checksum, adjointness; half the
    edges in the block (thr 0.15) and every edge in the block (thr 0)."""
    plain = synthetic_bipartite(100_000, 20_000, 2000, device=DEV)
    g = torch.Generator(device=DEV).manual_seed(5)
    hs = {w: torch.randn(getattr(plain, w).n_src, 400, device=DEV, generator=g) for w in ("cell_csr", "gene_csr")}
    o_plain = {w: sd.spmm(getattr(plain, w), hs[w])[0] for w in hs}
    rs = {"cell_csr": plain.rowsum_c, "gene_csr": plain.local_colsum_g}
    nnz = plain.cell_csr.nnz
    del plain
    for thr, lo, hi in ((0.15, 0.4, 0.6), (0.0, 1.0, 1.0)):
        split = synthetic_bipartite(100_000, 20_000, 2000, device=DEV).densify(thr)
        assert lo <= split.cell_csr.dense.nnz / nnz <= hi
        for which in ("cell_csr", "gene_csr"):
            b = getattr(split, which)
            o2 = sd.spmm(b, hs[which])[0]
            assert float((o_plain[which] - o2).abs().max() / o_plain[which].abs().max()) < 2e-5
            ones = sd.spmm(b, torch.ones(b.n_src, 400, device=DEV))[0]
            # all-positive sums are the worst case of the tensor core's truncating accumulation (2048-row chains: ≤ ~1.3e-5)
            assert float((ones[:, ::57] - rs[which][:, None]).abs().max() / rs[which].max()) < 2e-5
        hg = torch.randn(split.num_genes, 400, device=DEV, generator=g)
        yc = torch.randn(split.num_cells, 400, device=DEV, generator=g)
        ah = sd.spmm(split.cell_csr, hg)[0]
        lhs = (ah.double() * yc.double()).sum()
        rhs = (hg.double() * sd.spmm(split.gene_csr, yc)[0].double()).sum()
        assert abs(float(lhs - rhs)) < 1e-6 * float(ah.double().norm() * yc.double().norm())
        del split


@pytest.mark.parametrize("dense", [None, 0.2, 0.0])
def test_full_graph_trainer_host_inputs_match_device_inputs(dense):
    """FullGraphTrainer.step with pinned HOST tensors (cell rows copied on a side stream under the first
    cell<-gene pass, self-loop term added after the copy event) == the same steps with device tensors."""
    from scdeepsort_b200.synthetic import synthetic_features
    from scdeepsort_b200.trainer import FullGraphTrainer
    losses = {}
    for mode in ("device", "host"):
        bg = synthetic_bipartite(3000, 800, 60, device=DEV)
        feats = synthetic_features(bg, 128)
        if dense is not None:
            bg.densify(dense)
        labels = torch.randint(0, 5, (3000,), generator=torch.Generator().manual_seed(1)).to(DEV)
        tr = FullGraphTrainer(bg, 5, dense_dim=128, hidden_dim=128, n_layers=2, seed=3)
        f, l = (feats, labels) if mode == "device" else (feats.cpu().pin_memory(), labels.cpu().pin_memory())
        losses[mode] = [tr.step(f, l) for _ in range(4)]
    assert losses["device"][0] > 0
    for a, b in zip(losses["device"], losses["host"]):
        assert abs(a - b) < 2e-5 * abs(a)


def test_bf16_block_full_step_within_stated_tolerance(golden_train):
    """BASELINE configs[2] ("bf16 training"): X and the per-pass operand stored as bf16, fp32 accumulation.  The
    reference is fp32 (models/gnn.py:42), so this is an opt-in mode with its own bar: logits within 2e-2 of the
    fp32 golden logits (max|a-b|/max|b|; observed ~3e-3), gradients within 5e-2."""
    z = golden_train
    gg = golden_graph(z)
    params = golden_state(z, "L2")
    bg = sd.BipartiteGraph.from_expression(golden_csr(z), device=DEV).densify(0.0, fmt="bf16")
    assert bg.cell_csr.dense.fmt == _lib.D16_BF16
    model = sd.GNN(int(z["dense_dim"]), int(z["hidden"]), int(z["num_labels"]), 2, gg.num_genes, activation=torch.relu).to(DEV)
    model.load_state_dict(params)
    seeds = torch.arange(gg.num_genes, gg.num_nodes)
    labels = torch.from_numpy(z["labels"])[seeds]
    logits = model(sd.FullGraphFlow(bg, gg.features.to(DEV)))
    torch.nn.functional.cross_entropy(logits, labels.to(DEV), reduction="sum").backward()
    assert rel_err(logits.detach().cpu(), z["L2/logits"]) < 2e-2
    flow = graph_oracle.full_neighbor_flow(gg, seeds, 2)
    _, _, grads_ref = gnn_oracle.loss_and_grads(params, flow, labels, gg.num_genes, dtype=torch.float64)
    for name, v in grads_ref.items():
        assert rel_err(dict(model.named_parameters())[name].grad.cpu(), v) < 5e-2, name
