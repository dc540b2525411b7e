"""GPU parity on BASELINE configs[1] at its stated shape — the reference's full fixtures train/human/Adipose1372
(support) + test/human/Pancreas11 (test cells), dense_dim 400, models (L=1,H=200), (L=2,H=400), (L=2,H=200) — against
logits / losses / gradients produced by the UNMODIFIED reference (tests/golden/adipose.npz, oracle/gen_golden.py),
plus the parity rows that had no oracle comparison in round 1: predict_labels (bit-exact), dropout (product's masks
replayed through the oracle), a composed c4-shard training step in fp64, and DGL >= 0.5 block input.
Everything goes through the ctypes binding of the C ABI.  Tolerance: north_star's 1e-4 (max|a-b| / max|b|)."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

import scdeepsort_b200 as sd
from oracle import gnn_oracle, graph_oracle, spmm_oracle
from scds_helpers import (ReluMaskCapture, adipose_inputs, golden_csr, golden_graph, golden_state, rel_err, sampled_grad_err,
                          seeded_state)

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-4
MODELS = ("L1H200", "L2H400", "L2H200")


def _model(z, tag, dropout=0.0):
    state = seeded_state(z, tag, int(z["dense_dim"]), int(z["num_labels"]), int(z["num_genes"]))
    m = sd.GNN(int(z["dense_dim"]), int(z[f"{tag}/hidden"]), int(z["num_labels"]), int(z[f"{tag}/n_layers"]), int(z["num_genes"]),
               activation=torch.relu, dropout=dropout).to(DEV)
    m.load_state_dict(state)
    return m, state


@pytest.mark.parametrize("tag", MODELS)
def test_adipose_minibatch_logits_and_grads_match_reference(golden_adipose, tag):
    """train.py:71-85 / :94-105 with our sampler + GNN on the full Adipose graph: logits of every training cell
    (500-seed full-neighbour NodeFlows), CE(sum) loss and gradients of the reference's 64-seed batch."""
    z = golden_adipose
    n_layers = int(z[f"{tag}/n_layers"])
    x, _, feats = adipose_inputs(z)
    g = sd.DeepSortGraph.from_expression(x, features=feats).to(DEV)
    assert g.number_of_edges() == int(z["n_edges"])
    model, _ = _model(z, tag)
    model.eval()
    cells = torch.arange(g.num_genes, g.number_of_nodes())
    out = torch.zeros(g.number_of_nodes(), int(z["num_labels"]))
    for nf in sd.NeighborSampler(g, 500, g.number_of_nodes(), n_layers, 'in', shuffle=False, num_workers=8, seed_nodes=cells):
        nf.copy_from_parent()
        with torch.no_grad():
            out[nf.layer_parent_nid(-1).cpu()] = model(nf).cpu()
    err = rel_err(out[cells], z[f"{tag}/logits"])
    assert err < TOL and err < 1e-5, err
    model.train()
    seeds = torch.from_numpy(z[f"{tag}/grad_seeds"]).long()
    labels = torch.from_numpy(z["labels"].astype(np.int64))
    nf = next(iter(sd.NeighborSampler(g, len(seeds), g.number_of_nodes(), n_layers, 'in', seed_nodes=seeds)))
    nf.copy_from_parent()
    cap = ReluMaskCapture(model)
    loss = sd.optim.cross_entropy_sum(model(nf), labels.to(DEV)[nf.layer_parent_nid(-1)])
    cap.close()
    loss.backward()
    assert abs(float(loss) - float(z[f"{tag}/loss"])) < 1e-5 * float(z[f"{tag}/loss"])
    # gradients: (a) against the reference's stored ones — flip-tolerant bar, a ReLU decision that differs moves a weight
    # gradient by a whole term (ReluMaskCapture); (b) against the fp64 oracle under the product's own decisions: 1e-4
    for name, p in model.named_parameters():
        assert sampled_grad_err(p.grad, z, tag, name) < 1e-3, name
    og = graph_oracle.build_graph(x)
    og.features = feats
    flow = graph_oracle.full_neighbor_flow(og, seeds, n_layers)
    hidden = []
    state = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    gnn_oracle.forward(state, flow, og.num_genes, dtype=torch.float64, hidden_out=hidden)
    assert ReluMaskCapture.disagreement(cap.masks, hidden) < 1e-4
    pp = {k: v.clone().double().requires_grad_(True) for k, v in state.items()}
    ref = gnn_oracle.forward(pp, flow, og.num_genes, dtype=torch.float64, relu_masks=cap.masks)
    torch.nn.functional.cross_entropy(ref, labels[seeds], reduction="sum").backward()
    for name, p in model.named_parameters():
        assert rel_err(p.grad.cpu(), pp[name].grad) < TOL, name


@pytest.mark.parametrize("tag", MODELS)
@pytest.mark.parametrize("dense", [None, 0.05, 0.0])
def test_adipose_full_graph_logits_match_reference(golden_adipose, tag, dense):
    """The throughput form on the same fixture: every training cell in one layer-wise pass, CSR only / popular genes
    on the tensor cores (≥ 5 % density) / every gene on the tensor cores."""
    z = golden_adipose
    x, _, feats = adipose_inputs(z)
    bg = sd.BipartiteGraph.from_expression(x, device=DEV)
    if dense is not None:
        bg.densify(dense)
        assert (bg.cell_csr.nnz == 0) == (dense == 0.0)
    model, _ = _model(z, tag)
    model.eval()
    with torch.no_grad():
        logits = model(sd.FullGraphFlow(bg, feats.to(DEV))).cpu()
    err = rel_err(logits, z[f"{tag}/logits"])
    assert err < TOL and err < 2e-5, err


@pytest.mark.parametrize("tag", MODELS)
def test_pancreas_test_cells_logits_match_reference(golden_adipose, tag):
    """'human test set, 2-layer hidden=400, inference vs reference CPU logits' (predict.py:61-76): the 11 Pancreas
    cells on the Adipose support graph — mini-batch path, full-graph path, full-graph path with the dense block."""
    z = golden_adipose
    n_layers = int(z[f"{tag}/n_layers"])
    x, xt, feats = adipose_inputs(z, with_test=True)
    nid = torch.from_numpy(z["test_nid"]).long()
    model, _ = _model(z, tag)
    model.eval()
    g = sd.DeepSortGraph.from_expression(x, xt, features=feats).to(DEV)
    assert g.number_of_edges() == int(z["test_n_edges"])
    outs = []
    for nf in sd.NeighborSampler(g, 500, g.number_of_nodes(), n_layers, 'in', shuffle=False, seed_nodes=nid):
        nf.copy_from_parent()
        with torch.no_grad():
            outs.append(model(nf).cpu())
    err = rel_err(torch.cat(outs), z[f"{tag}/test_logits"])
    assert err < TOL and err < 1e-5, err
    for dense in (None, 0.03):
        bg = sd.BipartiteGraph.from_expression(x, xt, device=DEV)
        if dense is not None:
            bg.densify(dense)
        with torch.no_grad():
            full = model(sd.FullGraphFlow(bg, feats.to(DEV), seeds=(nid - bg.num_genes).to(DEV))).cpu()
        err = rel_err(full, z[f"{tag}/test_logits"])
        assert err < TOL and err < 2e-5, (dense, err)


def test_predict_labels_bit_equal_to_oracle(golden_adipose):
    """softmax / argmax / 'unsure' (train.py:106-113, predict.py:77-87): index work, bit-exact against the oracle on
    the same logits — golden logits, random logits, and rows sitting exactly on the unsure edge."""
    z = golden_adipose
    g = torch.Generator().manual_seed(0)
    cases = [torch.from_numpy(z["L2H400/logits"]), torch.from_numpy(z["L1H200/test_logits"]),
             torch.randn(5000, 16, generator=g) * 3, torch.randn(3000, 4, generator=g) * 0.05, torch.zeros(7, 11)]
    k = 8
    edge = torch.full((64, k), -20.0)
    edge[:, 0] = torch.linspace(-21.0, -19.0, 64)           # max prob sweeps through unsure_rate / k
    cases.append(edge)
    for logits in cases:
        for rate in (0.0, 1.0, 2.0, 3.5, float(logits.shape[1])):
            ref = gnn_oracle.predict_labels(logits, rate)
            got = sd.predict_labels(logits.to(DEV), rate).cpu()
            assert got.dtype == ref.dtype and torch.equal(got, ref), (tuple(logits.shape), rate)
    both = gnn_oracle.predict_labels(cases[2], 3.5)            # the rule fires for some rows and not for others
    assert (both >= 0).any() and (both < 0).any()


@pytest.mark.parametrize("path", ["nodeflow", "full", "full_dense"])
@pytest.mark.parametrize("n_layers", [1, 2])
def test_dropout_masks_replayed_through_oracle(golden_train, path, n_layers):
    """models/gnn.py:33-36,62-64: inverted dropout (p = 0.1, train.py:129) on layer-i node features before aggregation.
    The product's masks are captured with a forward hook on ``model.dropout`` and replayed through
    ``gnn_oracle.forward(dropout_masks=...)``: logits and gradients must agree."""
    z = golden_train
    gg = golden_graph(z)
    params = golden_state(z, f"L{n_layers}")
    model = sd.GNN(int(z["dense_dim"]), int(z["hidden"]), int(z["num_labels"]), n_layers, gg.num_genes,
                   activation=torch.relu, dropout=0.1).to(DEV)
    model.load_state_dict(params)
    model.train()
    masks, nonzero = [], []

    def capture(mod, inp, out):          # mask = out / in wherever the input is non-zero (elsewhere it cannot matter)
        masks.append(torch.where(inp[0] != 0, out / inp[0], torch.ones_like(out)).detach())
        nonzero.append((inp[0] != 0).detach())

    model.dropout.register_forward_hook(capture)
    torch.manual_seed(5)
    seeds = torch.arange(gg.num_genes, gg.num_nodes)
    labels = torch.from_numpy(z["labels"])[seeds]
    if path == "nodeflow":
        g = sd.DeepSortGraph.from_edges(gg.src, gg.dst, gg.weight, gg.node_id, gg.features, gg.num_genes).to(DEV)
        nf = next(iter(sd.NeighborSampler(g, len(seeds), g.number_of_nodes(), n_layers, 'in', seed_nodes=seeds)))
        nf.copy_from_parent()
        logits = model(nf)
        order = nf.layer_parent_nid(-1).cpu() - gg.num_genes
        layer_nids = [nf.layer_parent_nid(i).cpu() for i in range(n_layers)]
    else:
        bg = sd.BipartiteGraph.from_expression(golden_csr(z), device=DEV)
        if path == "full_dense":
            bg.densify(0.1)
        logits = model(sd.FullGraphFlow(bg, gg.features.to(DEV)))
        order = torch.arange(len(seeds))
        layer_nids = None
    loss = torch.nn.functional.cross_entropy(logits, labels[order].to(DEV), reduction="sum")
    loss.backward()
    assert len(masks) == n_layers
    for m, nz in zip(masks, nonzero):                      # an inverted-dropout mask: 0 or 1/(1-p), ~10 % zeros
        v = m[nz]
        assert bool(((v == 0) | ((v - 1 / 0.9).abs() < 1e-5)).all())
        assert 0.08 < float((v == 0).float().mean()) < 0.12
    flow = graph_oracle.full_neighbor_flow(gg, seeds[order], n_layers)
    if layer_nids is None:
        # the full-graph form drops out the [G + C, D] state of every layer; the oracle's flow lists the nodes each layer needs
        omasks = [m.cpu()[flow.layer_nid[i]] for i, m in enumerate(masks)]
    else:
        for i in range(n_layers):
            assert torch.equal(layer_nids[i], flow.layer_nid[i])
        omasks = [m.cpu() for m in masks]
    p = {k: v.detach().clone().double().requires_grad_(True) for k, v in params.items()}
    ref = gnn_oracle.forward(p, flow, gg.num_genes, dtype=torch.float64, dropout_masks=omasks)
    loss_ref = torch.nn.functional.cross_entropy(ref, labels[order], reduction="sum")
    loss_ref.backward()
    assert rel_err(logits.detach().cpu(), ref.detach()) < 1e-5
    assert abs(float(loss) - float(loss_ref)) < 1e-5 * float(loss_ref)
    for k, v in p.items():
        assert rel_err(dict(model.named_parameters())[k].grad.cpu(), v.grad) < TOL, k
    model.eval()                                            # eval mode: identity (models/gnn.py:33-36 via nn.Dropout)
    n_before = len(masks)
    with torch.no_grad():
        if path == "nodeflow":
            e = model(nf)
        else:
            e = model(sd.FullGraphFlow(bg, gg.features.to(DEV)))
    flow0 = graph_oracle.full_neighbor_flow(gg, seeds[order], n_layers)
    assert rel_err(e.cpu(), gnn_oracle.forward(params, flow0, gg.num_genes, dtype=torch.float64)) < 1e-5
    assert all(bool((m == 1).all()) for m in masks[n_before:])


@pytest.mark.parametrize("dense", [None, 0.0])
def test_c4_shard_composed_step_vs_fp64_closed_form(dense):
    """The bench's own workload, small enough for the CPU: the first 4096 cells of the c4 atlas (760k x 20k generator,
    avg-degree 2000), full 400/400 two-layer training step — logits, CE(sum) loss and every gradient against
    oracle/spmm_oracle.py in fp64 (itself pinned to the literal oracle in test_oracle_golden.py)."""
    from scdeepsort_b200.synthetic import synthetic_bipartite, synthetic_features
    n, genes, dim, k = 4096, 20_000, 400, 16
    bg = synthetic_bipartite(760_000, genes, 2000, seed=10086, device=DEV, cell_range=(0, n))
    feats = synthetic_features(bg, dim, seed=10086)
    cs = bg.cell_csr
    col = cs.col.cpu().numpy().view(np.uint16).astype(np.int64)
    x = sp.csr_matrix((cs.x.cpu().numpy(), col, cs.rowptr.cpu().numpy()), shape=(n, genes))
    if dense is not None:
        bg.densify(dense)
    params = gnn_oracle.init_params(dim, dim, k, 2, genes, seed=10086, perturb_alpha=True)
    model = sd.GNN(dim, dim, k, 2, genes, activation=torch.relu).to(DEV)
    model.load_state_dict(params)
    model.train()
    labels = torch.randint(0, k, (n,), generator=torch.Generator().manual_seed(10086))
    cap = ReluMaskCapture(model)
    logits = model(sd.FullGraphFlow(bg, feats))
    cap.close()
    loss = sd.optim.cross_entropy_sum(logits, labels.to(DEV))
    loss.backward()
    graph = spmm_oracle.SpmmGraph(x, torch.float64)
    hidden = []
    ref = spmm_oracle.forward({name: v.double() for name, v in params.items()}, graph, feats.cpu().double(), 2, hidden_out=hidden)
    loss_ref = torch.nn.functional.cross_entropy(ref, labels, reduction="sum")
    assert rel_err(logits.detach().cpu(), ref) < 2e-5
    assert abs(float(loss) - float(loss_ref)) < 1e-5 * float(loss_ref)
    # gradients under the product's own ReLU decisions (see ReluMaskCapture); the decisions themselves differ from the
    # oracle's in a negligible share of the 10^7 pre-activations
    assert ReluMaskCapture.disagreement(cap.masks, hidden) < 1e-4
    p = {name: v.detach().clone().double().requires_grad_(True) for name, v in params.items()}
    ref_m = spmm_oracle.forward(p, graph, feats.cpu().double(), 2, relu_masks=cap.masks)
    torch.nn.functional.cross_entropy(ref_m, labels, reduction="sum").backward()
    for name, v in p.items():
        assert rel_err(dict(model.named_parameters())[name].grad.cpu(), v.grad) < TOL, name


class _FakeDglBlock:
    """What GNN.forward needs from a DGL >= 0.5 message-flow block, nothing more (no DGL in this image)."""

    def __init__(self, src, dst, weight, src_id, dst_id, features=None):
        self._src, self._dst = src, dst
        self.edata = {"weight": weight}
        self.srcdata = {"id": src_id}
        self.dstdata = {"id": dst_id}
        if features is not None:
            self.srcdata["features"] = features
        self._n = (int(src_id.shape[0]), int(dst_id.shape[0]))

    def edges(self):
        return self._src, self._dst

    def num_src_nodes(self):
        return self._n[0]

    def num_dst_nodes(self):
        return self._n[1]


@pytest.mark.parametrize("n_layers", [1, 2])
def test_dgl_block_list_input_matches_nodeflow(golden_train, n_layers):
    """north_star's 'DGL NodeDataLoader mini-batch surface': ``model(blocks)`` with duck-typed DGL >= 0.5 blocks (COO
    edges in arbitrary order, [E, 1] weights, [N, 1] ids as the reference stores them) == the NodeFlow path == oracle."""
    z = golden_train
    gg = golden_graph(z)
    params = golden_state(z, f"L{n_layers}")
    seeds = torch.arange(gg.num_genes, gg.num_genes + 50)
    flow = graph_oracle.full_neighbor_flow(gg, seeds, n_layers)
    rng = np.random.RandomState(1)
    blocks = []
    for i, b in enumerate(flow.blocks):
        perm = torch.from_numpy(rng.permutation(b.src.shape[0]))
        blocks.append(_FakeDglBlock(b.src[perm], b.dst[perm], b.weight[perm].unsqueeze(1), flow.layer_id[i].unsqueeze(1),
                                    flow.layer_id[i + 1].unsqueeze(1), flow.features if i == 0 else None))
    model = sd.GNN(int(z["dense_dim"]), int(z["hidden"]), int(z["num_labels"]), n_layers, gg.num_genes, activation=torch.relu).to(DEV)
    model.load_state_dict(params)
    model.eval()
    with torch.no_grad():
        got = model(blocks).cpu()
    assert rel_err(got, gnn_oracle.forward(params, flow, gg.num_genes, dtype=torch.float64)) < 1e-5
    assert rel_err(got, z[f"L{n_layers}/logits"][:50]) < 1e-5
    with pytest.raises(ValueError, match="blocks"):
        model(blocks + blocks)
