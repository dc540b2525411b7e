"""Dense-block split  X = X_sparse + X_dense  (csrc/agg_dense.cuh, BipartiteGraph.densify): the popular genes'
entries leave the CSRs and run on the FMA-bound dense kernel; results must match the plain CSR path (fp32
summation order aside), the fp64 reference and the golden logits / oracle gradients."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

import scdeepsort_b200 as sd
from oracle import gnn_oracle, graph_oracle
from scdeepsort_b200.synthetic import synthetic_bipartite
from scds_helpers import golden_csr, golden_graph, golden_state, rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-4


def _skewed_expression(n_cells, n_genes, seed, avg=0.08):
    rng = np.random.RandomState(seed)
    pop = np.minimum(1.0, avg * n_genes ** 0.8 / 5 * (np.arange(1, n_genes + 1)) ** -0.8)[rng.permutation(n_genes)]
    m = rng.rand(n_cells, n_genes) < pop[None, :]
    return sp.csr_matrix(np.where(m, rng.uniform(0.05, 9, (n_cells, n_genes)), 0).astype(np.float32))


@pytest.mark.parametrize("n_cells,n_genes,dim,thr", [(1000, 700, 400, 0.2), (333, 257, 400, 0.05), (5000, 300, 200, 0.3),
                                                      (700, 900, 128, 0.2), (640, 500, 132, 0.2), (150, 90, 512, 0.5),
                                                      (70000, 120, 400, 0.2)])
def test_spmm_dense_block_matches_plain_csr_and_fp64(n_cells, n_genes, dim, thr):
    x = _skewed_expression(n_cells, n_genes, n_cells + dim)
    plain = sd.BipartiteGraph.from_expression(x, device=DEV)
    split = sd.BipartiteGraph.from_expression(x, device=DEV).densify(thr, directions=("gene", "cell"))
    assert split.densified and split.cell_csr.dense is not None and split.gene_csr.dense is not None
    assert split.cell_csr.nnz + split.cell_csr.dense.nnz == plain.cell_csr.nnz
    assert split.gene_csr.nnz + split.gene_csr.dense.nnz == plain.gene_csr.nnz
    g = torch.Generator(device=DEV).manual_seed(7)
    xd = torch.from_numpy(x.toarray()).double()
    for which, ref_mat in (("cell_csr", xd), ("gene_csr", xd.t())):
        a, b = getattr(plain, which), getattr(split, which)
        hs = torch.randn(a.n_src, dim, device=DEV, generator=g)
        hself = torch.randn(a.n_dst, dim, device=DEV, generator=g)
        q = torch.randn(a.n_dst, dim, device=DEV, generator=g)
        dscale = torch.rand(a.n_dst, device=DEV, generator=g) + 0.5
        selfcoef = torch.rand(a.n_dst, device=DEV, generator=g)
        kw = dict(dscale=dscale, selfcoef=selfcoef, hself=hself, want_raw=True, q=q, want_dot=True)
        o1, r1, d1 = sd.spmm(a, hs, **kw)
        o2, r2, d2 = sd.spmm(b, hs, **kw)
        acc = ref_mat @ hs.double().cpu()
        ref_out = dscale.double().cpu()[:, None] * acc + selfcoef.double().cpu()[:, None] * hself.double().cpu()
        assert rel_err(r2.cpu(), acc) < 1e-5
        assert rel_err(o2.cpu(), ref_out) < 1e-5
        assert rel_err(d2.cpu(), (acc * q.double().cpu()).sum(1)) < 1e-5
        assert rel_err(o2.cpu(), o1.cpu()) < 1e-5 and rel_err(r2.cpu(), r1.cpu()) < 1e-5
        assert torch.equal(o2, sd.spmm(b, hs, **kw)[0])                  # deterministic
        with pytest.raises(RuntimeError):
            sd.spmm(b, hs, algo=1)                                       # the gather kernel cannot take the block


@pytest.mark.parametrize("n_layers", [1, 2, 3])
def test_full_graph_fwd_bwd_with_dense_block_vs_oracle(golden_train, n_layers):
    """Golden Muscle sub-sample: reference logits (L <= 2) and fp64 oracle gradients, popular genes densified."""
    z = golden_train
    gg = golden_graph(z)
    if n_layers <= 2:
        params = golden_state(z, f"L{n_layers}")
    else:
        params = gnn_oracle.init_params(int(z["dense_dim"]), int(z["hidden"]), int(z["num_labels"]), n_layers,
                                        gg.num_genes, perturb_alpha=True)
    bg = sd.BipartiteGraph.from_expression(golden_csr(z), device=DEV).densify(0.2, directions=("gene", "cell") if n_layers != 2 else ("gene",))
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
    """BASELINE configs[2] shape (100k x 20k, 2e8 edges): densified vs plain kernels, checksum, adjointness."""
    plain = synthetic_bipartite(100_000, 20_000, 2000, device=DEV)
    split = synthetic_bipartite(100_000, 20_000, 2000, device=DEV).densify(0.2, directions=("gene", "cell"))
    assert 0.3 < split.cell_csr.dense.nnz / plain.cell_csr.nnz < 0.6
    g = torch.Generator(device=DEV).manual_seed(5)
    for which in ("cell_csr", "gene_csr"):
        a, b = getattr(plain, which), getattr(split, which)
        hs = torch.randn(a.n_src, 400, device=DEV, generator=g)
        o1, o2 = sd.spmm(a, hs)[0], sd.spmm(b, hs)[0]
        assert float((o1 - o2).abs().max() / o1.abs().max()) < 5e-5
        ones = sd.spmm(b, torch.ones(a.n_src, 400, device=DEV))[0]
        rs = plain.rowsum_c if which == "cell_csr" else plain.local_colsum_g
        assert float((ones[:, ::57] - rs[:, None]).abs().max() / rs.max()) < 1e-4
    hg = torch.randn(plain.num_genes, 400, device=DEV, generator=g)
    yc = torch.randn(plain.num_cells, 400, device=DEV, generator=g)
    ah = sd.spmm(split.cell_csr, hg)[0]
    lhs = (ah.double() * yc.double()).sum()
    rhs = (hg.double() * sd.spmm(split.gene_csr, yc)[0].double()).sum()
    assert abs(float(lhs - rhs)) < 1e-6 * float(ah.double().norm() * yc.double().norm())


@pytest.mark.parametrize("dense", [False, True])
def test_full_graph_trainer_host_inputs_match_device_inputs(dense):
    """FullGraphTrainer.step with pinned HOST tensors (cell rows copied on a side stream under the first
    cell<-gene pass, self-loop term added after the copy event) == the same steps with device tensors."""
    from scdeepsort_b200.synthetic import synthetic_features
    from scdeepsort_b200.trainer import FullGraphTrainer
    losses = {}
    for mode in ("device", "host"):
        bg = synthetic_bipartite(3000, 800, 60, device=DEV)
        feats = synthetic_features(bg, 128)
        if dense:
            bg.densify(0.2)
        labels = torch.randint(0, 5, (3000,), generator=torch.Generator().manual_seed(1)).to(DEV)
        tr = FullGraphTrainer(bg, 5, dense_dim=128, hidden_dim=128, n_layers=2, seed=3)
        f, l = (feats, labels) if mode == "device" else (feats.cpu().pin_memory(), labels.cpu().pin_memory())
        losses[mode] = [tr.step(f, l) for _ in range(4)]
    assert losses["device"][0] > 0
    for a, b in zip(losses["device"], losses["host"]):
        assert abs(a - b) < 2e-5 * abs(a)
