"""GPU neighbour sampler (wsage_sample_neighbors) and sampled NodeFlows: exact-k uniform draws without
replacement, NodeFlow consistency with the parent graph, and logits parity with the oracle on the very
blocks that were sampled (sampled runs are only comparable to the reference statistically)."""
import numpy as np
import pytest
import torch

from oracle import gnn_oracle
from oracle.gnn_oracle import OracleBlock, OracleFlow
from scds_helpers import golden_graph, golden_state, rel_err

import scdeepsort_b200 as sd
from scdeepsort_b200.nodeflow import _sample_edges_cuda

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _toy_graph(degs):
    rowptr = torch.zeros(len(degs) + 1, dtype=torch.int64)
    rowptr[1:] = torch.cumsum(torch.tensor(degs), 0)
    e = int(rowptr[-1])
    g = sd.DeepSortGraph(0, len(degs), rowptr, torch.arange(e) % len(degs), torch.ones(e), {})
    return g.to(DEV)


def test_kernel_draws_uniform_k_subsets():
    degs = [100, 3, 0, 10, 11, 1000]
    g = _toy_graph(degs)
    nodes = torch.arange(len(degs), device=DEV)
    k, trials = 10, 4000
    counts = torch.zeros(100, device=DEV)
    for t in range(trials):
        eid, deg = _sample_edges_cuda(g, nodes, k, seed=12345 + t)
        assert deg.tolist() == [10, 3, 0, 10, 10, 10]
        parts = torch.split(eid, deg.tolist())
        for v, part in enumerate(parts):
            lo, hi = int(g.in_rowptr[v]), int(g.in_rowptr[v + 1])
            assert torch.all((part >= lo) & (part < hi))
            assert torch.all(part[1:] > part[:-1])                       # ascending, hence no duplicates
        assert torch.equal(parts[1], torch.arange(100, 103, device=DEV))  # deg <= fanout keeps every in-edge
        counts[parts[0]] += 1
    freq = counts / trials                                               # each edge of node 0: p = k/deg = 0.1
    assert float(freq.mean()) == pytest.approx(0.1, abs=1e-6)
    sd_expected = (0.1 * 0.9 / trials) ** 0.5                            # 0.0047
    assert float((freq - 0.1).abs().max()) < 5 * sd_expected
    a, _ = _sample_edges_cuda(g, nodes, k, seed=7)
    b, _ = _sample_edges_cuda(g, nodes, k, seed=7)
    assert torch.equal(a, b)                                             # reproducible
    with pytest.raises(RuntimeError, match="fanout"):
        _sample_edges_cuda(g, nodes, 33, seed=1)


@pytest.mark.parametrize("fanouts", [[7, 7], [12, 5, 3]])
def test_sampled_nodeflow_matches_oracle_on_same_blocks(golden_train, fanouts):
    z = golden_train
    gg = golden_graph(z)
    n_layers = len(fanouts)
    params = golden_state(z, "L2") if n_layers == 2 else gnn_oracle.init_params(
        int(z["dense_dim"]), int(z["hidden"]), int(z["num_labels"]), n_layers, gg.num_genes, perturb_alpha=True)
    g = sd.DeepSortGraph.from_edges(gg.src, gg.dst, gg.weight, gg.node_id, gg.features, gg.num_genes).to(DEV)
    seeds = torch.from_numpy(z["train_ids"])[:50]
    sampler = sd.NeighborSampler(g, 50, num_hops=n_layers, neighbor_type='in', seed_nodes=seeds, fanouts=fanouts, seed=3)
    nf = next(iter(sampler))
    nf.copy_from_parent()
    full_deg = g.in_rowptr[1:] - g.in_rowptr[:-1]
    blocks = []
    for i, b in enumerate(nf.blocks):
        deg = b.rowptr[1:] - b.rowptr[:-1]
        assert torch.equal(deg, torch.clamp(full_deg[nf.layer_parent_nid(i + 1)], max=fanouts[n_layers - 1 - i]))
        eid = nf.block_parent_eid(i)
        assert torch.equal(b.weight, g.in_weight[eid])
        assert torch.equal(nf.layer_parent_nid(i)[b.col.long()], g.in_src[eid])
        dst = torch.repeat_interleave(torch.arange(b.n_dst, device=DEV), deg)
        blocks.append(OracleBlock(b.col.long().cpu(), dst.cpu(), b.weight.cpu(), b.n_src, b.n_dst))
    flow = OracleFlow([nf.layer_parent_nid(i).cpu() for i in range(n_layers + 1)],
                      [nf.layers[i].data["id"].cpu() for i in range(n_layers + 1)],
                      nf.layers[0].data["features"].cpu(), blocks)
    labels = torch.from_numpy(z["labels"])[seeds]
    loss_ref, logits_ref, grads_ref = gnn_oracle.loss_and_grads(params, flow, labels, gg.num_genes, dtype=torch.float64)
    model = sd.GNN(int(z["dense_dim"]), int(z["hidden"]), int(z["num_labels"]), n_layers, gg.num_genes,
                   activation=torch.relu).to(DEV)
    model.load_state_dict(params)
    logits = model(nf)
    torch.nn.functional.cross_entropy(logits, labels.to(DEV), reduction="sum").backward()
    assert rel_err(logits.detach().cpu(), logits_ref) < 1e-5
    for k, v in grads_ref.items():
        assert rel_err(dict(model.named_parameters())[k].grad.cpu(), v) < 1e-4, k
    # a second batch draws a different sample
    nf2 = sampler.build(seeds.to(DEV))
    assert not torch.equal(nf2.block_parent_eid(n_layers - 1), nf.block_parent_eid(n_layers - 1))


def test_trainer_draws_new_neighbour_samples_every_epoch(monkeypatch):
    """train.py:71-78 re-samples with fresh randomness each epoch; the device sampler is keyed by
    (seed, batch, hop, node), so Trainer must hand every epoch its own seed, and the seed must follow random_seed."""
    import scipy.sparse as sp
    from scdeepsort_b200 import trainer as trainer_mod
    rng = np.random.RandomState(0)
    x = sp.csr_matrix(np.where(rng.rand(120, 80) < 0.3, rng.rand(120, 80) + 0.1, 0).astype(np.float32))
    feats = torch.randn(200, 16, generator=torch.Generator().manual_seed(0))
    graph = sd.DeepSortGraph.from_expression(x, features=feats)
    labels = torch.cat([torch.full((80,), -1, dtype=torch.int64), torch.randint(0, 3, (120,))])
    seen = []
    real = trainer_mod.NeighborSampler

    def spy(*a, **kw):
        seen.append(kw.get("seed"))
        return real(*a, **kw)

    monkeypatch.setattr(trainer_mod, "NeighborSampler", spy)

    def epochs(seed):
        torch.manual_seed(seed)
        tr = trainer_mod.Trainer(graph, labels, torch.arange(80, 180), torch.arange(180, 200), 3, dense_dim=16, hidden_dim=16,
                                 n_layers=2, dropout=0.0, batch_size=50, num_neighbors=4, device="cuda:0")
        del seen[:]
        tr.train(); tr.train(); tr.train()
        return list(seen)

    a, b = epochs(1), epochs(2)
    assert len(set(a)) == 3 and None not in a          # three epochs, three different sampler seeds
    assert a == epochs(1) and set(a).isdisjoint(b)     # reproducible under the same random_seed, different under another
