"""Pins the CPU oracle against vectors produced by the UNMODIFIED reference code
(oracle/gen_golden.py: /root/reference/models/gnn.py + utils/preprocess*.py on the DGL shim)."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

from oracle import gnn_oracle, graph_oracle
from scds_helpers import (adipose_inputs, golden_csr, golden_grads, golden_graph, golden_state, rel_err, sampled_grad_err,
                          seeded_state)

TOL = 2e-6   # fp32 oracle vs fp32 reference: only summation order differs


def _same_edges(g, z):
    """Edge *sets* must match exactly.  Within one cell the reference inserts edges in the data
    file's gene-column order (preprocess_internal.py:157-160), which the saved support matrix
    does not keep, so edges are compared in canonical (dst, src) order; normalised weights then
    differ at most by the fp32 summation order of ``torch.sum`` (preprocess_internal.py:23)."""
    n = g.num_nodes
    ka = g.dst.numpy() * n + g.src.numpy()
    kb = z["graph/dst"].astype(np.int64) * n + z["graph/src"].astype(np.int64)
    oa, ob = np.argsort(ka, kind="stable"), np.argsort(kb, kind="stable")
    assert np.array_equal(ka[oa], kb[ob])
    wa, wb = g.weight.numpy()[oa], z["graph/weight"][ob]
    assert np.max(np.abs(wa - wb) / wb) < 4e-7
    assert np.mean(wa == wb) > 0.5          # the rest differ by ≤ 2 ulp (sum order)
    assert np.array_equal(g.node_id.numpy(), z["graph/node_id"])
    assert np.array_equal(g.src.numpy()[-n:], np.arange(n)) and np.all(g.weight.numpy()[-n:] == 1)


def test_graph_oracle_matches_reference_train_graph(golden_train):
    z = golden_train
    g = graph_oracle.build_graph(golden_csr(z))
    assert g.num_genes == int(z["num_genes"]) and g.num_cells == int(z["num_cells"])
    _same_edges(g, z)
    feats = graph_oracle.make_features(golden_csr(z), z["gene_feat"])
    assert rel_err(feats, z["graph/features"]) < 1e-6                     # gene_feat stored as fp32


def test_graph_oracle_matches_reference_test_graph(golden_test):
    z = golden_test
    g = graph_oracle.build_graph(golden_csr(z), golden_csr(z, "xt"))
    _same_edges(g, z)
    nid = torch.from_numpy(z["test_nid"])
    assert nid.min() == g.num_genes + int(z["n_support"]) and len(nid) == int(z["n_test"])
    x_all = sp.vstack([golden_csr(z), golden_csr(z, "xt")])
    assert rel_err(graph_oracle.make_features(x_all, z["gene_feat"]), z["graph/features"]) < 1e-6


@pytest.mark.parametrize("n_layers", [1, 2])
def test_oracle_logits_match_reference(golden_train, n_layers):
    z = golden_train
    g = golden_graph(z)
    params = golden_state(z, f"L{n_layers}")
    seeds = torch.arange(g.num_genes, g.num_nodes)
    got = []
    for s in range(0, len(seeds), 100):     # same batching as the golden run (order-independent anyway)
        flow = graph_oracle.full_neighbor_flow(g, seeds[s:s + 100], n_layers)
        got.append(gnn_oracle.forward(params, flow, g.num_genes))
    assert rel_err(torch.cat(got), z[f"L{n_layers}/logits"]) < TOL


@pytest.mark.parametrize("n_layers", [1, 2])
def test_oracle_grads_match_reference(golden_train, n_layers):
    z = golden_train
    g = golden_graph(z)
    params = golden_state(z, f"L{n_layers}")
    seeds = torch.from_numpy(z[f"L{n_layers}/grad_seeds"])
    labels = torch.from_numpy(z["labels"])
    flow = graph_oracle.full_neighbor_flow(g, seeds, n_layers)
    loss, logits, grads = gnn_oracle.loss_and_grads(params, flow, labels[seeds], g.num_genes)
    assert abs(float(loss) - float(z[f"L{n_layers}/loss"])) < 1e-4 * abs(float(z[f"L{n_layers}/loss"]))
    assert rel_err(logits, z[f"L{n_layers}/train_logits"]) < TOL
    ref = golden_grads(z, f"L{n_layers}")
    assert set(ref) == set(grads)
    for k in ref:
        assert rel_err(grads[k], ref[k]) < 2e-5, k


@pytest.mark.parametrize("n_layers", [1, 2])
def test_oracle_inference_logits_match_reference(golden_test, n_layers):
    z = golden_test
    g = golden_graph(z, num_genes=int(z["num_genes"]))
    params = golden_state(z, f"L{n_layers}")
    flow = graph_oracle.full_neighbor_flow(g, torch.from_numpy(z["test_nid"]), n_layers)
    assert rel_err(gnn_oracle.forward(params, flow, g.num_genes), z[f"L{n_layers}/logits"]) < TOL


@pytest.mark.parametrize("tag,n_layers", [(t, l) for t in ("tiny", "c1s") for l in (1, 2, 3)])
def test_oracle_synthetic_match_reference(golden_syn, tag, n_layers):
    z = golden_syn
    g = golden_graph(z, f"{tag}/graph/")
    params = golden_state(z, f"{tag}/L{n_layers}")
    seeds = torch.arange(g.num_genes, g.num_nodes)
    flow = graph_oracle.full_neighbor_flow(g, seeds, n_layers)
    assert rel_err(gnn_oracle.forward(params, flow, g.num_genes), z[f"{tag}/L{n_layers}/logits"]) < TOL
    labels = torch.from_numpy(z[f"{tag}/labels"])
    flow = graph_oracle.full_neighbor_flow(g, seeds[:17], n_layers)
    loss, _, grads = gnn_oracle.loss_and_grads(params, flow, labels[seeds[:17]], g.num_genes)
    assert abs(float(loss) - float(z[f"{tag}/L{n_layers}/loss"])) < 1e-5 * max(1.0, abs(float(loss)))
    for k, v in golden_grads(z, f"{tag}/L{n_layers}").items():
        assert rel_err(grads[k], v) < 2e-5, k


def test_tiny_graph_analytic(golden_syn):
    """2 genes × 3 cells: closed form of SURVEY §8a checked by hand-arithmetic in fp64."""
    z = golden_syn
    x = z["tiny/x"]                                  # [[1,2],[0,3],[4,0]]
    g = golden_graph(z, "tiny/graph/")
    params = golden_state(z, "tiny/L1")
    a = params["alpha"].double().squeeze(1)          # [G+2]
    feats = g.features.double()
    hg, hc = feats[:2], feats[2:]
    deg_c = (x > 0).sum(1); rowsum = x.sum(1)
    neigh = []
    for c in range(3):
        acc = a[3] * hc[c]                           # α_{G+1}·h_c  (cell self loop, weight 1)
        for gi in range(2):
            if x[c, gi] > 0:
                acc = acc + a[gi] * (x[c, gi] * deg_c[c] / rowsum[c]) * hg[gi]
        neigh.append(acc / (deg_c[c] + 1))
    neigh = torch.stack(neigh)
    h = torch.relu(neigh @ params["layers.0.fc_neigh.weight"].double().t() + params["layers.0.fc_neigh.bias"].double())
    logits = h @ params["linear.weight"].double().t() + params["linear.bias"].double()
    assert rel_err(logits, z["tiny/L1/logits"]) < 1e-6


def test_alpha_index_cascade():
    src = np.array([3, -1, 2, -1]); dst = np.array([-1, 5, 7, -1])
    assert gnn_oracle.alpha_index(src, dst, 10).tolist() == [3, 5, 10, 11]


def test_unsure_rule():
    logits = torch.tensor([[5.0, 0, 0, 0], [0.1, 0, 0, 0]])
    assert gnn_oracle.predict_labels(logits, 2.0).tolist() == [0, -1]    # 0.27 < 2/4 → unsure
    assert gnn_oracle.predict_labels(logits, 0.0).tolist() == [0, 0]


@pytest.mark.parametrize("n_layers", [1, 2, 3])
def test_closed_form_spmm_oracle_matches_literal_oracle(n_layers):
    """oracle/spmm_oracle.py (bench.py's optimised CPU baseline: sparse-CSR products, no message tensor) computes
    the same logits as the literal edge-materialising restatement of models/gnn.py:47-68."""
    import scipy.sparse as sp
    from oracle import spmm_oracle
    rng = np.random.RandomState(n_layers)
    c, g, d, h, k = 60, 40, 8, 12, 4
    x = sp.csr_matrix(np.where(rng.rand(c, g) < 0.2, rng.rand(c, g) + 0.1, 0).astype(np.float32))
    og = graph_oracle.build_graph(x)
    og.features = torch.randn(g + c, d, generator=torch.Generator().manual_seed(0))
    params = gnn_oracle.init_params(d, h, k, n_layers, g, perturb_alpha=True)
    flow = graph_oracle.full_neighbor_flow(og, torch.arange(g, g + c), n_layers)
    ref = gnn_oracle.forward(params, flow, g, dtype=torch.float64)
    out = spmm_oracle.forward({n: v.double() for n, v in params.items()}, spmm_oracle.SpmmGraph(x, torch.float64),
                              og.features.double(), n_layers)
    assert float((out - ref).abs().max() / ref.abs().max()) < 1e-6


# ---- BASELINE configs[1] at its stated shape: Adipose1372 (+ Pancreas11), dense_dim 400 -----------------------------
ADIPOSE_MODELS = ("L1H200", "L2H400", "L2H200")


def test_adipose_graph_oracle_matches_reference_graph_summaries(golden_adipose):
    """The full fixtures are too large to store edge by edge: edge counts and weight sums of the reference-built graphs."""
    z = golden_adipose
    x, xt, feats = adipose_inputs(z, with_test=True)
    g = graph_oracle.build_graph(x)
    assert (g.num_genes, g.num_cells) == (int(z["num_genes"]), int(z["num_cells"])) == (16397, 1354)
    assert g.src.shape[0] == int(z["n_edges"]) == 2 * x.nnz + g.num_nodes
    assert abs(float(g.weight.double().sum()) - float(z["weight_sum"])) < 1e-6 * float(z["weight_sum"])
    assert abs(float((g.weight.double() ** 2).sum()) - float(z["weight_sq_sum"])) < 1e-6 * float(z["weight_sq_sum"])
    gt = graph_oracle.build_graph(x, xt)
    assert gt.src.shape[0] == int(z["test_n_edges"]) == 2 * x.nnz + xt.nnz + gt.num_nodes and xt.nnz == 10296
    assert abs(float(gt.weight.double().sum()) - float(z["test_weight_sum"])) < 1e-6 * float(z["test_weight_sum"])
    assert np.array_equal(z["test_nid"], np.arange(gt.num_genes + g.num_cells, gt.num_nodes))
    assert rel_err(feats[: g.num_nodes].reshape(-1)[::997], z["feat_sample"]) < 1e-6


@pytest.mark.parametrize("tag", ADIPOSE_MODELS)
def test_adipose_oracle_test_cell_logits_match_reference(golden_adipose, tag):
    """'human test set, 2-layer hidden=400, inference vs reference CPU logits': the 11 Pancreas cells."""
    z = golden_adipose
    x, xt, feats = adipose_inputs(z, with_test=True)
    g = graph_oracle.build_graph(x, xt)
    g.features = feats
    params = seeded_state(z, tag, int(z["dense_dim"]), int(z["num_labels"]), g.num_genes)
    flow = graph_oracle.full_neighbor_flow(g, torch.from_numpy(z["test_nid"]).long(), int(z[f"{tag}/n_layers"]))
    assert rel_err(gnn_oracle.forward(params, flow, g.num_genes), z[f"{tag}/test_logits"]) < TOL


def test_adipose_oracle_train_logits_and_grads_match_reference(golden_adipose):
    """The reference's native shape (n_layers 1, hidden 200): logits of 300 training cells, loss and gradients of
    the 64-seed batch."""
    z, tag = golden_adipose, "L1H200"
    x, _, feats = adipose_inputs(z)
    g = graph_oracle.build_graph(x)
    g.features = feats
    params = seeded_state(z, tag, int(z["dense_dim"]), int(z["num_labels"]), g.num_genes)
    cells = torch.arange(g.num_genes, g.num_genes + 300)
    flow = graph_oracle.full_neighbor_flow(g, cells, 1)
    assert rel_err(gnn_oracle.forward(params, flow, g.num_genes), z[f"{tag}/logits"][:300]) < TOL
    seeds = torch.from_numpy(z[f"{tag}/grad_seeds"]).long()
    labels = torch.from_numpy(z["labels"].astype(np.int64))
    flow = graph_oracle.full_neighbor_flow(g, seeds, 1)
    loss, _, grads = gnn_oracle.loss_and_grads(params, flow, labels[seeds], g.num_genes)
    assert abs(float(loss) - float(z[f"{tag}/loss"])) < 1e-4 * abs(float(z[f"{tag}/loss"]))
    for k, v in grads.items():
        assert sampled_grad_err(v, z, tag, k) < 2e-5, k
