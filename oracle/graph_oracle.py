"""CPU oracle for the graph contract the hot path receives.  TEST INFRASTRUCTURE ONLY.

Literal restatement of the parts of the reference's graph builders that define the hot
path's inputs (node order, ``ndata['id']``, edge order, per-destination weight
normalisation, self-loops, features):
  * training graph   /root/reference/utils/preprocess_internal.py:107-110,156-173,183-215
  * ``normalize_weight``  /root/reference/utils/preprocess_internal.py:15-23  (per-node loop, fp32)
  * inference graph  /root/reference/utils/preprocess.py:102-134,167-187,192-221
    (support cells: both directions; test cells: gene→cell only; self-loops after normalisation)
and of DGL-0.4 full-neighbour NodeFlow construction (``oracle/dgl_shim`` documents the
semantics).  File parsing / label bookkeeping / PCA fitting are out of scope (SURVEY §2
rows 7-8): callers pass the expression matrix and the gene features in.
"""
from dataclasses import dataclass
from typing import List, Optional

import numpy as np
import scipy.sparse as sp
import torch

from .gnn_oracle import OracleBlock, OracleFlow


@dataclass
class OracleGraph:
    num_genes: int
    num_cells: int
    src: torch.Tensor        # int64 [E]  (edge id = position, insertion order as in the reference)
    dst: torch.Tensor        # int64 [E]
    weight: torch.Tensor     # fp32  [E]
    node_id: torch.Tensor    # int32 [N]  gene → index, cell → -1
    features: Optional[torch.Tensor] = None   # fp32 [N, D0]

    @property
    def num_nodes(self):
        return self.num_genes + self.num_cells


def _normalize_weight(src, dst, weight, n_nodes):
    """preprocess_internal.py:15-23: w_e ← indeg(v)·w_e / Σ_{e'→v} w_e', fp32, per-node loop."""
    weight = weight.clone()
    order = torch.sort(dst, stable=True).indices
    counts = torch.bincount(dst, minlength=n_nodes)
    ptr = torch.zeros(n_nodes + 1, dtype=torch.int64)
    ptr[1:] = torch.cumsum(counts, 0)
    for v in range(n_nodes):
        eid = order[ptr[v]:ptr[v + 1]]
        if eid.shape[0] == 0:
            continue
        w = weight[eid].unsqueeze(1)
        weight[eid] = (counts[v] * w / torch.sum(w)).squeeze(1)
    return weight


def build_graph(x_support: sp.csr_matrix, x_test: Optional[sp.csr_matrix] = None,
                threshold: float = 0.0) -> OracleGraph:
    """x_support [C, G] gets both edge directions; x_test [Ct, G] (inference graphs only,
    preprocess.py:185-187) gets gene→cell edges only.  Returns the graph *after*
    normalisation and self-loop insertion."""
    x_support = sp.csr_matrix(x_support)
    num_genes = x_support.shape[1]
    srcs, dsts, ws = [], [], []
    n_cells = 0
    for mat, both in ((x_support, True), (x_test, False)):
        if mat is None:
            continue
        coo = sp.csr_matrix(mat).tocoo()
        keep = coo.data > threshold                       # preprocess_internal.py:158
        order = np.lexsort((coo.col[keep], coo.row[keep]))  # row-major, as np.nonzero(arr > t)
        row = coo.row[keep][order].astype(np.int64)
        col = coo.col[keep][order].astype(np.int64)
        val = torch.tensor(coo.data[keep][order], dtype=torch.float32)   # :171,173
        cell = torch.from_numpy(row) + num_genes + n_cells
        gene = torch.from_numpy(col)
        if both:
            srcs.append(cell); dsts.append(gene); ws.append(val)         # cell → gene (:170)
        srcs.append(gene); dsts.append(cell); ws.append(val)             # gene → cell (:172)
        n_cells += mat.shape[0]
    n_nodes = num_genes + n_cells
    src, dst, w = torch.cat(srcs), torch.cat(dsts), torch.cat(ws)
    w = _normalize_weight(src, dst, w, n_nodes)                          # :211
    loops = torch.arange(n_nodes, dtype=torch.int64)                     # :213-214
    src, dst = torch.cat([src, loops]), torch.cat([dst, loops])
    w = torch.cat([w, torch.ones(n_nodes, dtype=torch.float32)])
    node_id = torch.cat([torch.arange(num_genes, dtype=torch.int32),
                         torch.full((n_cells,), -1, dtype=torch.int32)])
    return OracleGraph(num_genes, n_cells, src, dst, w, node_id)


def make_features(x_all: sp.csr_matrix, gene_feat: np.ndarray) -> torch.Tensor:
    """preprocess_internal.py:194-202: cell_feat = (X / (rowsum(X)+1e-6)) · gene_feat in fp64,
    features = cat[gene_feat; cell_feat] cast to fp32.  x_all stacks support then test cells."""
    dense = np.asarray(sp.csr_matrix(x_all).todense(), dtype=np.float64)
    dense = dense / (np.sum(dense, axis=1, keepdims=True) + 1e-6)
    cell_feat = dense.dot(np.asarray(gene_feat, dtype=np.float64))
    return torch.cat([torch.from_numpy(np.asarray(gene_feat, dtype=np.float64)),
                      torch.from_numpy(cell_feat)], dim=0).type(torch.float)


def full_neighbor_flow(g: OracleGraph, seeds: torch.Tensor, num_hops: int,
                       fanout: Optional[int] = None, rng: Optional[np.random.Generator] = None) -> OracleFlow:
    """DGL-0.4 NodeFlow for ``seeds``: every in-edge per hop (``fanout`` None / ≥ degree), else
    ``fanout`` in-edges drawn uniformly without replacement (self-loop is an ordinary edge)."""
    n = g.num_nodes
    order = torch.sort(g.dst, stable=True).indices
    ptr = torch.zeros(n + 1, dtype=torch.int64)
    ptr[1:] = torch.cumsum(torch.bincount(g.dst, minlength=n), 0)
    layer_nid: List[torch.Tensor] = [None] * (num_hops + 1)
    blocks: List[OracleBlock] = [None] * num_hops
    layer_nid[num_hops] = seeds.to(torch.int64)
    for hop in range(num_hops, 0, -1):
        dst_nodes = layer_nid[hop]
        chunks = []
        for v in dst_nodes.tolist():
            e = order[ptr[v]:ptr[v + 1]]
            if fanout is not None and e.shape[0] > fanout:
                pick = np.sort(rng.choice(e.shape[0], fanout, replace=False))
                e = e[torch.from_numpy(pick)]
            chunks.append(e)
        eid = torch.cat(chunks) if chunks else torch.zeros(0, dtype=torch.int64)
        src_parent, dst_parent = g.src[eid], g.dst[eid]
        src_nodes = torch.unique(src_parent)
        layer_nid[hop - 1] = src_nodes
        lookup = torch.full((n,), -1, dtype=torch.int64)
        lookup[dst_nodes] = torch.arange(dst_nodes.shape[0])
        blocks[hop - 1] = OracleBlock(torch.searchsorted(src_nodes, src_parent), lookup[dst_parent],
                                      g.weight[eid], int(src_nodes.shape[0]), int(dst_nodes.shape[0]))
    layer_id = [g.node_id[nid] for nid in layer_nid]
    feats = g.features[layer_nid[0]] if g.features is not None else None
    return OracleFlow(layer_nid, layer_id, feats, blocks)
