"""GPU parity tests (run on the B200 box with -m gpu).  Everything goes through the package's
ctypes binding of the C ABI; the checker is the CPU oracle and the golden vectors minted by the
unmodified reference.  Tolerance: north_star's 1e-4 relative (max|a-b| / max|b|) for fp32."""
import numpy as np
import pytest
import torch

from oracle import gnn_oracle, graph_oracle
from oracle.gnn_oracle import OracleBlock
from scds_helpers import golden_csr, golden_grads, golden_graph, golden_state, rel_err

import scdeepsort_b200 as sd

pytestmark = pytest.mark.gpu
TOL = 1e-4          # north_star tolerance; observed errors are ~1e-6
DEV = "cuda:0"


def _random_block(n_src, n_dst, max_deg, gene_num, rng, force=None):
    deg = rng.integers(0, max_deg + 1, n_dst)
    if force is not None:
        deg[:len(force)] = force
    rowptr = np.zeros(n_dst + 1, dtype=np.int64); rowptr[1:] = np.cumsum(deg)
    col = rng.integers(0, n_src, rowptr[-1]).astype(np.int32)
    w = rng.uniform(0.1, 4.0, rowptr[-1]).astype(np.float32)
    src_id = np.where(rng.random(n_src) < 0.6, rng.integers(0, gene_num, n_src), -1).astype(np.int32)
    dst_id = np.where(rng.random(n_dst) < 0.5, rng.integers(0, gene_num, n_dst), -1).astype(np.int32)
    return rowptr, col, w, src_id, dst_id


@pytest.mark.parametrize("dim", [1, 18, 48, 132, 400, 516, 800])
def test_block_aggregate_fwd_bwd_vs_oracle(dim):
    rng = np.random.default_rng(dim)
    gene_num, n_src, n_dst = 50, 301, 77
    rowptr, col, w, src_id, dst_id = _random_block(n_src, n_dst, 70, gene_num, rng, force=[0, 1, 1, 33, 64, 0])
    h = torch.from_numpy(rng.normal(0, 1, (n_src, dim)).astype(np.float32))
    alpha = torch.from_numpy(rng.uniform(0.5, 1.5, (gene_num + 2, 1)).astype(np.float32))
    dst = np.repeat(np.arange(n_dst), np.diff(rowptr))
    ob = OracleBlock(torch.from_numpy(col.astype(np.int64)), torch.from_numpy(dst), torch.from_numpy(w), n_src, n_dst)
    h64, a64 = h.double().requires_grad_(True), alpha.double().requires_grad_(True)
    ref = gnn_oracle.block_aggregate(h64, a64, ob, torch.from_numpy(src_id), torch.from_numpy(dst_id), gene_num)
    dout = torch.from_numpy(rng.normal(0, 1, (n_dst, dim)).astype(np.float32))
    ref.backward(dout.double())

    blk = sd.Block(torch.from_numpy(rowptr).to(DEV), torch.from_numpy(col).to(DEV), torch.from_numpy(w).to(DEV), n_src, n_dst)
    hd, ad = h.to(DEV).requires_grad_(True), alpha.to(DEV).requires_grad_(True)
    out = sd.block_aggregate(hd, ad, blk, torch.from_numpy(src_id).to(DEV), torch.from_numpy(dst_id).to(DEV), gene_num)
    out.backward(dout.to(DEV))
    assert rel_err(out.detach().cpu(), ref.detach()) < 1e-5
    assert torch.all(out[torch.tensor([0, 5])] == 0)            # zero in-degree rows → 0
    assert rel_err(hd.grad.cpu(), h64.grad) < 1e-5
    assert rel_err(ad.grad.cpu(), a64.grad) < 1e-5


def test_block_single_edge():
    """E == 1: the reference's indices.squeeze() yields a 0-d index (models/gnn.py:54)."""
    blk = sd.Block(torch.tensor([0, 1], device=DEV), torch.tensor([0], dtype=torch.int32, device=DEV),
                   torch.tensor([2.0], device=DEV), 1, 1)
    h = torch.arange(8, dtype=torch.float32, device=DEV).reshape(1, 8)
    alpha = torch.tensor([[3.0], [5.0], [7.0]], device=DEV)       # G = 1
    out = sd.block_aggregate(h, alpha, blk, torch.tensor([0], dtype=torch.int32, device=DEV),
                             torch.tensor([-1], dtype=torch.int32, device=DEV), 1)
    assert torch.equal(out.cpu(), (h * 3.0 * 2.0).cpu())


def _model_from(params, n_layers, gene_num, dev=DEV):
    m = sd.GNN(params["layers.0.fc_neigh.weight"].shape[1], params["linear.weight"].shape[1],
               params["linear.weight"].shape[0], n_layers, gene_num, activation=torch.relu).to(dev)
    m.load_state_dict(params)
    return m


@pytest.mark.parametrize("n_layers", [1, 2])
def test_minibatch_logits_match_reference_golden(golden_train, n_layers):
    """The loop of train.py:94-105 with our sampler + GNN against logits produced by the reference."""
    z = golden_train
    gg = golden_graph(z)
    g = sd.DeepSortGraph.from_edges(gg.src, gg.dst, gg.weight, gg.node_id, gg.features, gg.num_genes).to(DEV)
    model = _model_from(golden_state(z, f"L{n_layers}"), n_layers, gg.num_genes).eval()
    seeds = torch.arange(gg.num_genes, gg.num_nodes)
    out = torch.zeros(gg.num_nodes, int(z["num_labels"]))
    for nf in sd.NeighborSampler(g, 100, g.number_of_nodes(), n_layers, 'in', shuffle=False, num_workers=8, seed_nodes=seeds):
        nf.copy_from_parent()
        with torch.no_grad():
            logits = model(nf).cpu()
        out[nf.layer_parent_nid(-1).cpu()] = logits
    assert rel_err(out[seeds], z[f"L{n_layers}/logits"]) < TOL
    assert rel_err(out[seeds], z[f"L{n_layers}/logits"]) < 1e-5     # what we actually get


@pytest.mark.parametrize("n_layers", [1, 2])
def test_minibatch_grads_match_reference_golden(golden_train, n_layers):
    z = golden_train
    gg = golden_graph(z)
    g = sd.DeepSortGraph.from_edges(gg.src, gg.dst, gg.weight, gg.node_id, gg.features, gg.num_genes).to(DEV)
    model = _model_from(golden_state(z, f"L{n_layers}"), n_layers, gg.num_genes).train()
    seeds = torch.from_numpy(z[f"L{n_layers}/grad_seeds"])
    labels = torch.from_numpy(z["labels"]).to(DEV)
    nf = next(iter(sd.NeighborSampler(g, len(seeds), g.number_of_nodes(), n_layers, 'in', seed_nodes=seeds)))
    nf.copy_from_parent()
    logits = model(nf)
    loss = torch.nn.CrossEntropyLoss(reduction='sum')(logits, labels[nf.layer_parent_nid(-1)])
    loss.backward()
    assert abs(float(loss) - float(z[f"L{n_layers}/loss"])) < 1e-4 * float(z[f"L{n_layers}/loss"])
    assert rel_err(logits.detach().cpu(), z[f"L{n_layers}/train_logits"]) < 1e-5
    for k, v in golden_grads(z, f"L{n_layers}").items():
        got = dict(model.named_parameters())[k].grad.cpu()
        assert rel_err(got, v) < TOL, k


@pytest.mark.parametrize("n_layers", [1, 2])
def test_inference_graph_logits_match_reference_golden(golden_test, n_layers):
    """predict.py:61-76 on the reference-built inference graph (test cells: gene→cell edges only)."""
    z = golden_test
    gg = golden_graph(z, num_genes=int(z["num_genes"]))
    params = golden_state(z, f"L{n_layers}")
    nid = torch.from_numpy(z["test_nid"])
    # (a) mini-batch path on the reference's edges
    g = sd.DeepSortGraph.from_edges(gg.src, gg.dst, gg.weight, gg.node_id, gg.features, gg.num_genes).to(DEV)
    model = _model_from(params, n_layers, gg.num_genes).eval()
    outs = []
    for nf in sd.NeighborSampler(g, 16, g.number_of_nodes(), n_layers, 'in', shuffle=False, seed_nodes=nid):
        nf.copy_from_parent()
        with torch.no_grad():
            outs.append(model(nf).cpu())
    assert rel_err(torch.cat(outs), z[f"L{n_layers}/logits"]) < 1e-5
    # (b) full-graph path rebuilt from the expression matrices (our builder, our factorisation)
    bg = sd.BipartiteGraph.from_expression(golden_csr(z), golden_csr(z, "xt"), device=DEV)
    flow = sd.FullGraphFlow(bg, gg.features.to(DEV), seeds=(nid - gg.num_genes).to(DEV))
    with torch.no_grad():
        full = model(flow).cpu()
    assert rel_err(full, z[f"L{n_layers}/logits"]) < 1e-5
    assert torch.equal(flow.layer_parent_nid(-1).cpu(), nid)


@pytest.mark.parametrize("n_layers,algo", [(1, 1), (2, 1), (3, 1), (2, 0), (2, 2)])
def test_full_graph_path_fwd_bwd(golden_train, n_layers, algo):
    """Full-graph layer-wise path vs reference logits (L ≤ 2) and vs oracle gradients of CE(sum) over all cells."""
    z = golden_train
    gg = golden_graph(z)
    if n_layers <= 2:
        params = golden_state(z, f"L{n_layers}")
    else:
        params = gnn_oracle.init_params(int(z["dense_dim"]), int(z["hidden"]), int(z["num_labels"]), n_layers,
                                        gg.num_genes, perturb_alpha=True)
    bg = sd.BipartiteGraph.from_expression(golden_csr(z), device=DEV)
    model = _model_from(params, n_layers, gg.num_genes).train()
    model.spmm_algo = algo
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
    for k, v in grads_ref.items():
        assert rel_err(dict(model.named_parameters())[k].grad.cpu(), v) < TOL, k


@pytest.mark.parametrize("tag,n_layers", [(t, l) for t in ("tiny", "c1s") for l in (1, 2, 3)])
def test_synthetic_golden_both_paths(golden_syn, tag, n_layers):
    """Hand-sized graph and odd feature width (D0 = 18 → scalar kernels), generic and full-graph paths."""
    z = golden_syn
    gg = golden_graph(z, f"{tag}/graph/")
    params = golden_state(z, f"{tag}/L{n_layers}")
    model = _model_from(params, n_layers, gg.num_genes).eval()
    seeds = torch.arange(gg.num_genes, gg.num_nodes)
    g = sd.DeepSortGraph.from_edges(gg.src, gg.dst, gg.weight, gg.node_id, gg.features, gg.num_genes).to(DEV)
    nf = next(iter(sd.NeighborSampler(g, len(seeds), g.number_of_nodes(), n_layers, 'in', seed_nodes=seeds)))
    nf.copy_from_parent()
    with torch.no_grad():
        a = model(nf).cpu()
        bg = sd.BipartiteGraph.from_expression(z[f"{tag}/x"], device=DEV)
        b = model(sd.FullGraphFlow(bg, gg.features.to(DEV))).cpu()
    assert rel_err(a, z[f"{tag}/L{n_layers}/logits"]) < 1e-5
    assert rel_err(b, z[f"{tag}/L{n_layers}/logits"]) < 1e-5


@pytest.mark.parametrize("n_src,dim,algo", [(300, 64, 1), (70000, 32, 1), (300, 18, 1), (300, 400, 0), (300, 400, 2), (5000, 128, 2),
                                            (2000, 200, 2), (700, 132, 2), (900, 512, 2), (70000, 400, 2), (43, 400, 2), (1, 64, 2)])
def test_spmm_all_outputs(n_src, dim, algo):
    """wsage_spmm epilogue options (dscale / self / raw / dot / row_perm, uint16 and int32 columns)."""
    rng = np.random.default_rng(n_src + dim)
    n_dst = 211
    deg = rng.integers(0, 60, n_dst); deg[:3] = [0, 1, 59]
    rowptr = np.zeros(n_dst + 1, dtype=np.int64); rowptr[1:] = np.cumsum(deg)
    deg = np.minimum(deg, n_src)
    rowptr[1:] = np.cumsum(deg)
    col = np.concatenate([np.sort(rng.choice(n_src, d, replace=False)) for d in deg]).astype(np.int64)
    x = rng.uniform(0.05, 9, rowptr[-1]).astype(np.float32)
    hs = torch.from_numpy(rng.normal(0, 1, (n_src, dim)).astype(np.float32))
    hself = torch.from_numpy(rng.normal(0, 1, (n_dst, dim)).astype(np.float32))
    q = torch.from_numpy(rng.normal(0, 1, (n_dst, dim)).astype(np.float32))
    dscale = torch.from_numpy(rng.uniform(0.1, 2, n_dst).astype(np.float32))
    selfcoef = torch.from_numpy(rng.uniform(0.1, 2, n_dst).astype(np.float32))
    dst = np.repeat(np.arange(n_dst), deg)
    acc = torch.zeros(n_dst, dim, dtype=torch.float64).index_add(0, torch.from_numpy(dst), hs.double()[col] * torch.from_numpy(x).double()[:, None])
    ref_out = dscale.double()[:, None] * acc + selfcoef.double()[:, None] * hself.double()
    if n_src <= 65536:
        colt, bits = torch.from_numpy(col.astype(np.uint16).view(np.int16)), 16
    else:
        colt, bits = torch.from_numpy(col.astype(np.int32)), 32
    perm = torch.from_numpy(rng.permutation(n_dst).astype(np.int32))
    csr = sd.Csr(torch.from_numpy(rowptr).to(DEV), colt.to(DEV), torch.from_numpy(x).to(DEV), n_src, n_dst, bits, perm.to(DEV))
    out, raw, dot = sd.spmm(csr, hs.to(DEV), dscale=dscale.to(DEV), selfcoef=selfcoef.to(DEV), hself=hself.to(DEV),
                            want_raw=True, q=q.to(DEV), want_dot=True, algo=algo)
    assert rel_err(out.cpu(), ref_out) < 1e-5
    assert rel_err(raw.cpu(), acc) < 1e-5
    assert rel_err(dot.cpu(), (acc * q.double()).sum(1)) < 1e-5
    out2, _, _ = sd.spmm(csr, hs.to(DEV), algo=algo)      # plain Σ, no epilogue terms
    assert rel_err(out2.cpu(), acc) < 1e-5
    csr.row_perm = None                                   # identity row order
    out3, _, _ = sd.spmm(csr, hs.to(DEV), algo=algo)
    assert rel_err(out3.cpu(), acc) < 1e-5


def test_wrapper_rejects_bad_inputs():
    blk = sd.Block(torch.tensor([0, 1], device=DEV), torch.tensor([0], dtype=torch.int32, device=DEV),
                   torch.tensor([2.0], device=DEV), 1, 1)
    ids = torch.tensor([0], dtype=torch.int32, device=DEV)
    alpha = torch.ones(3, 1, device=DEV)
    with pytest.raises(ValueError):
        sd.block_aggregate(torch.ones(1, 8), alpha, blk, ids, ids, 1)                       # CPU tensor
    with pytest.raises(ValueError):
        sd.block_aggregate(torch.ones(1, 8, device=DEV, dtype=torch.float64), alpha, blk, ids, ids, 1)
    with pytest.raises(ValueError):
        sd.block_aggregate(torch.ones(2, 8, device=DEV), alpha, blk, ids, ids, 1)           # wrong n_src


def test_launch_counter_counts_our_kernels():
    sd._lib.launch_count(reset=True)
    test_block_single_edge()
    assert sd._lib.launch_count() == 1
