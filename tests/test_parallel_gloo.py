"""Multi-rank host logic on CPU (gloo, world_size 2): cell sharding, global gene normalisers, the
autograd all-reduce and the flattened gradient all-reduce.  The CUDA kernel is replaced IN THE TEST
PROCESS ONLY by a torch restatement of wsage_spmm's contract (monkeypatched `ops.spmm`), so what is
exercised is the product's sharding arithmetic, not its kernels (those are covered by -m gpu tests)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import scdeepsort_b200 as sd
from scdeepsort_b200 import parallel
from scdeepsort_b200.synthetic import synthetic_bipartite, synthetic_features

C, G, DEG, D0, H, K, LAYERS = 90, 70, 12, 8, 12, 4, 3


def _cols(csr):
    c = csr.col.to(torch.int64)
    return c & 0xFFFF if csr.col_bits == 16 else c


def _spmm_cpu(csr, hs, *, dscale=None, selfcoef=None, hself=None, out=None, want_out=True, raw=None,
              want_raw=False, q=None, want_dot=False, algo=0, src_scale=None):
    if src_scale is not None:
        hs = hs * src_scale[:, None]
    seg = torch.repeat_interleave(torch.arange(csr.n_dst), csr.rowptr[1:] - csr.rowptr[:-1])
    acc = torch.zeros(csr.n_dst, hs.shape[1]).index_add_(0, seg, hs[_cols(csr)] * csr.x[:, None])
    if getattr(csr, "dense", None) is not None:      # what wsage_dense16 + wsage_spmm(init=...) add: the block's entries
        from scds_helpers import csr_dense_matrix
        acc = acc + csr_dense_matrix(csr) @ hs
    o = acc if dscale is None else acc * dscale[:, None]
    if selfcoef is not None:
        o = o + selfcoef[:, None] * hself
    if out is not None:
        out.copy_(o)
        o = out
    return (o if want_out or out is not None else None), (acc if want_raw else None), ((acc * q).sum(1) if want_dot else None)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _model():
    torch.manual_seed(7)
    m = sd.GNN(D0, H, K, LAYERS, G, activation=torch.relu)
    with torch.no_grad():
        m.alpha.copy_(0.5 + torch.rand(m.alpha.shape))
    return m


def _labels():
    return torch.randint(0, K, (C,), generator=torch.Generator().manual_seed(3))


def _full_reference():
    sd.ops.spmm = _spmm_cpu
    bg = synthetic_bipartite(C, G, DEG, device="cpu")
    feats = synthetic_features(bg, D0)
    m = _model()
    logits = parallel.sharded_forward(m, bg, feats)          # world_size 1: all-reduces are no-ops
    loss = torch.nn.functional.cross_entropy(logits, _labels(), reduction="sum")
    loss.backward()
    return bg, feats, logits.detach(), float(loss), {k: p.grad.clone() for k, p in m.named_parameters()}


def _worker(rank, world, port, q, densify=None):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sd.ops.spmm = _spmm_cpu
        lo, hi = parallel.cell_ranges(C, world)[rank]
        bg = synthetic_bipartite(C, G, DEG, device="cpu", cell_range=(lo, hi))
        parallel.globalize_gene_normalisers(bg)
        feats = synthetic_features(bg, D0)
        if densify:            # every rank picks ITS OWN popular genes from its local degrees: no coordination needed
            bg.densify(densify[rank])
            assert bg.densified and bg.gene_csr.dense is not None
        m = _model()
        parallel.broadcast_params(m)
        logits = parallel.sharded_forward(m, bg, feats)
        loss = torch.nn.functional.cross_entropy(logits, _labels()[lo:hi], reduction="sum")
        loss.backward()
        parallel.allreduce_grads(m)
        total = loss.detach().clone()
        dist.all_reduce(total)
        # autograd all-reduce: y = sum_r x_r ; dL/dx_r = sum_r g_r
        x = torch.full((3,), float(rank + 1), requires_grad=True)
        y = parallel.AllReduceSum.apply(x)
        (y * (rank + 1)).sum().backward()
        # numpy, not tensors: tensors travel through the queue as shared-memory fds, which die with this process
        q.put((rank, lo, hi, bg.norm_g.numpy(), bg.mean_g.numpy(), logits.detach().numpy(), float(total),
               {k: p.grad.numpy() for k, p in m.named_parameters()}, y.detach().numpy(), x.grad.numpy()))
    finally:
        dist.destroy_process_group()


def test_cell_ranges_partition():
    for n, w in ((10, 3), (760000, 8), (5, 8), (0, 2)):
        r = parallel.cell_ranges(n, w)
        assert len(r) == w and r[0][0] == 0 and r[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
        assert max(h - l for l, h in r) - min(h - l for l, h in r) <= 1


@pytest.mark.parametrize("densify", [None, (0.15, 0.3)])
def test_two_rank_sharded_step_matches_single_process(densify):
    """densify: each rank moves its own (different) popular-gene set into dense blocks — the all-reduced gene sums
    and the gradients must not change."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, densify)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=600) for _ in range(world)], key=lambda t: t[0])
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    bg, feats, logits_full, loss_full, grads_full = _full_reference()
    t = torch.from_numpy
    for rank, lo, hi, norm_g, mean_g, logits, total, grads, y, xg in res:
        norm_g, mean_g, logits, y, xg = t(norm_g), t(mean_g), t(logits), t(y), t(xg)
        grads = {k: t(v) for k, v in grads.items()}
        assert torch.allclose(norm_g, bg.norm_g, rtol=1e-6) and torch.allclose(mean_g, bg.mean_g)
        assert torch.allclose(logits, logits_full[lo:hi], rtol=1e-4, atol=1e-5)
        assert abs(total - loss_full) < 1e-4 * abs(loss_full)
        for k, g in grads_full.items():
            assert torch.allclose(grads[k], g, rtol=2e-4, atol=1e-5), k
        assert torch.equal(y, torch.full((3,), 3.0)) and torch.equal(xg, torch.full((3,), 3.0))
    assert np.array_equal(res[0][7]["alpha"], res[1][7]["alpha"])        # identical after the all-reduce


def test_sharded_math_matches_oracle():
    """The world_size-1 path of parallel.sharded_forward (same code the ranks run) against the oracle."""
    import scipy.sparse as sp
    from oracle import gnn_oracle, graph_oracle
    bg, feats, logits_full, loss_full, grads_full = _full_reference()
    cs = bg.cell_csr
    x = sp.csr_matrix((cs.x.numpy(), _cols(cs).numpy(), cs.rowptr.numpy()), shape=(C, G))
    og = graph_oracle.build_graph(x)
    og.features = feats
    flow = graph_oracle.full_neighbor_flow(og, torch.arange(G, G + C), LAYERS)
    params = {k: v.detach() for k, v in _model().state_dict().items()}
    loss, logits, grads = gnn_oracle.loss_and_grads(params, flow, _labels(), G, dtype=torch.float64)
    assert float((logits - logits_full).abs().max() / logits.abs().max()) < 1e-5
    for k, g in grads.items():
        assert float((grads_full[k] - g).abs().max() / g.abs().max().clamp(min=1e-30)) < 1e-4, k


def _dropout_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sd.ops.spmm = _spmm_cpu
        torch.manual_seed(100 + rank)                      # the ranks' default generators differ, as in real runs
        lo, hi = parallel.cell_ranges(C, world)[rank]
        bg = synthetic_bipartite(C, G, DEG, device="cpu", cell_range=(lo, hi))
        parallel.globalize_gene_normalisers(bg)
        feats = synthetic_features(bg, D0)
        m = sd.GNN(D0, H, K, 2, G, activation=torch.relu, dropout=0.3)
        parallel.broadcast_params(m)
        m.train()
        seen = []
        real = sd.gnn._LayerAggregate.apply

        def spy(h, alpha, graph, algo, gene_too, ready, gmask, cmask, sharded):
            seen.append((gmask.clone(), cmask.clone()))
            return real(h, alpha, graph, algo, gene_too, ready, gmask, cmask, sharded)

        sd.gnn._LayerAggregate.apply = spy
        logits = parallel.sharded_forward(m, bg, feats)
        q.put((rank, [g.numpy() for g, _ in seen], [c.numpy() for _, c in seen], logits.detach().numpy()))
    finally:
        dist.destroy_process_group()


def test_sharded_dropout_uses_one_mask_for_the_replicated_gene_rows():
    """the gene rows are replicated state — under dropout every rank must drop the SAME gene entries (a shared
    generator), while the cell rows of each shard get their own masks."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dropout_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=600) for _ in range(world)], key=lambda t: t[0])
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    (_, g0, c0, _), (_, g1, c1, _) = res
    assert len(g0) == 2
    for a, b in zip(g0, g1):
        assert a.shape == (G, a.shape[1]) and np.array_equal(a, b)           # same gene mask on both ranks
        assert 0.2 < float((a == 0).mean()) < 0.4 and np.allclose(a[a != 0], 1 / 0.7)
    assert not np.array_equal(g0[0], g0[1])                                   # a fresh mask per layer
    assert c0[0].shape != c1[0].shape or not np.array_equal(c0[0], c1[0])    # cell masks are per shard
