"""Mini-batch surface: ``NeighborSampler`` → ``NodeFlow`` (mirrors the DGL 0.4.3 objects the
reference drives, train.py:71-81, predict.py:64-75) and the full-graph flow used when every
seed can be served by one layer-wise pass.

Same constructor keywords and iterator protocol as ``dgl.contrib.sampling.NeighborSampler``;
a yielded ``NodeFlow`` offers ``copy_from_parent()``, ``layer_parent_nid(i)``, ``layers[i].data``
and per-block CSR (``blocks[i]``) instead of DGL's edge frames.  Index manipulation is plain
``torch`` on whatever device the parent graph lives on (plumbing); the arithmetic on the
flow happens in ``libwsage.so`` via ``gnn.GNN``.
"""
from typing import List, Optional

import numpy as np
import torch

from .graph import BipartiteGraph, DeepSortGraph
from .ops import Block


class _Layer:
    def __init__(self):
        self.data = {}


class NodeFlow:
    def __init__(self, parent: DeepSortGraph, layer_nid: List[torch.Tensor], blocks: List[Block],
                 block_eid: List[torch.Tensor]):
        self._parent = parent
        self._layer_nid = layer_nid
        self._block_eid = block_eid
        self.blocks = blocks
        self.layers = [_Layer() for _ in layer_nid]

    @property
    def num_layers(self):
        return len(self._layer_nid)

    @property
    def num_blocks(self):
        return len(self.blocks)

    def layer_size(self, i):
        return int(self._layer_nid[i].shape[0])

    def block_size(self, i):
        return int(self.blocks[i].col.shape[0])

    def layer_parent_nid(self, i):
        return self._layer_nid[i]

    def block_parent_eid(self, i):
        return self._block_eid[i]

    def copy_from_parent(self):
        """Gather ``ndata`` per layer (train.py:79).  Edge weights were sliced when the block was built."""
        for i, nid in enumerate(self._layer_nid):
            for key, col in self._parent.ndata.items():
                if key == "features" and i != 0:
                    continue            # only layer 0 features are ever read (models/gnn.py:59)
                self.layers[i].data[key] = col[nid]

    def to(self, device):
        """Move a host-built flow to ``device`` (what ``copy_from_parent`` does for a CPU-resident DGL graph)."""
        nf = NodeFlow(self._parent, [t.to(device) for t in self._layer_nid],
                      [Block(b.rowptr.to(device), b.col.to(device), b.weight.to(device), b.n_src, b.n_dst)
                       for b in self.blocks], [t.to(device) for t in self._block_eid])
        for dst, src in zip(nf.layers, self.layers):
            dst.data = {k: v.to(device) for k, v in src.data.items()}
        return nf


def _sample_edges_cuda(g: DeepSortGraph, nodes: torch.Tensor, fanout: int, seed: int):
    """K6: wsage_sample_neighbors — warp-per-node Floyd sampling on the device (fanout <= 32)."""
    import ctypes
    from . import _lib
    n = int(nodes.shape[0])
    eid = torch.empty(n, fanout, dtype=torch.int64, device=nodes.device)
    deg = torch.empty(n, dtype=torch.int32, device=nodes.device)
    lib = _lib.load()
    p = lambda t: ctypes.c_void_p(t.data_ptr())      # noqa: E731
    _lib.check(lib.wsage_sample_neighbors(p(g.in_rowptr), p(nodes), n, fanout, ctypes.c_uint64(seed & (2 ** 64 - 1)),
                                          p(eid), p(deg), ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice()))),
               "wsage_sample_neighbors")
    deg = deg.to(torch.int64)
    keep = torch.arange(fanout, device=nodes.device)[None, :] < deg[:, None]
    return eid[keep], deg                                # row-major: grouped by node, ascending inside a node


def _in_edges(g: DeepSortGraph, nodes: torch.Tensor, fanout: Optional[int], gen: Optional[torch.Generator],
              seed: Optional[int] = None):
    """Edge positions (into g.in_src) of the in-edges of ``nodes``, grouped by node in order."""
    if fanout is not None and nodes.is_cuda and fanout <= 32 and nodes.shape[0] > 0:
        return _sample_edges_cuda(g, nodes.contiguous(), fanout, seed if seed is not None else 0)
    start = g.in_rowptr[nodes]
    deg = g.in_rowptr[nodes + 1] - start
    total = int(deg.sum())
    dev = nodes.device
    seg = torch.repeat_interleave(torch.arange(nodes.shape[0], device=dev), deg, output_size=total)
    first = torch.cumsum(deg, 0) - deg
    rank = torch.arange(total, device=dev) - first[seg]
    eid = start[seg] + rank
    if fanout is not None and total and int(deg.max()) > fanout:
        # uniform without replacement: keep the `fanout` smallest random keys of every row
        key = torch.rand(total, device=dev, generator=gen)
        order = torch.argsort(seg.to(torch.float64) + key.to(torch.float64), stable=True)
        keep = rank < fanout                      # position k inside a row of the sorted order = k-th smallest key
        sel = torch.sort(order[keep]).values      # chosen positions, back in (row, ascending edge) order
        eid = eid[sel]
        deg = torch.clamp(deg, max=fanout)
    return eid, deg


class NeighborSampler:
    """``for nf in NeighborSampler(g, batch_size, expand_factor, num_hops, 'in', shuffle=..., seed_nodes=...)``.

    ``expand_factor`` ≥ max in-degree (the reference passes ``num_cells + num_genes``) means full
    neighbourhood.  ``num_workers`` is accepted for signature compatibility; batches are built on
    the graph's device."""

    def __init__(self, g: DeepSortGraph, batch_size, expand_factor=None, num_hops=1, neighbor_type='in',
                 transition_prob=None, seed_nodes=None, shuffle=False, num_workers=1, prefetch=False,
                 add_self_loop=False, generator: Optional[torch.Generator] = None, fanouts=None, seed=10086):
        if neighbor_type != 'in':
            raise NotImplementedError("only in-neighbour sampling is on the hot path (train.py:75)")
        if transition_prob is not None or add_self_loop:
            raise NotImplementedError("transition_prob / add_self_loop are not used by the reference")
        self.g = g
        self.batch_size = int(batch_size)
        self.num_hops = int(num_hops)
        max_deg = int((g.in_rowptr[1:] - g.in_rowptr[:-1]).max()) if g.number_of_nodes() else 0
        clip = lambda f: None if f is None or int(f) >= max_deg else int(f)      # noqa: E731
        # reference: ONE expand_factor for every hop (train.py:73); `fanouts` (seed hop first, e.g. [25, 10, 5])
        # is the per-hop extension BASELINE.json's sampled configuration names
        self.fanouts = [clip(expand_factor)] * self.num_hops if fanouts is None else [clip(f) for f in fanouts]
        if len(self.fanouts) != self.num_hops:
            raise ValueError("fanouts needs one entry per hop")
        self.seed, self._batch = int(seed), 0
        if seed_nodes is None:
            seed_nodes = torch.arange(g.number_of_nodes())
        self.seed_nodes = torch.as_tensor(seed_nodes, dtype=torch.int64).reshape(-1).to(g.device)
        self.shuffle = shuffle
        self.generator = generator

    def __len__(self):
        return (self.seed_nodes.shape[0] + self.batch_size - 1) // self.batch_size

    def build(self, seeds: torch.Tensor) -> NodeFlow:
        g = self.g
        layer_nid = [None] * (self.num_hops + 1)
        blocks = [None] * self.num_hops
        eids = [None] * self.num_hops
        layer_nid[self.num_hops] = seeds
        self._batch += 1
        for hop in range(self.num_hops, 0, -1):
            dst_nodes = layer_nid[hop]
            eid, deg = _in_edges(g, dst_nodes, self.fanouts[self.num_hops - hop], self.generator,
                                 seed=self.seed + 1000003 * self._batch + hop)
            src_parent = g.in_src[eid].to(torch.int64)
            src_nodes = torch.unique(src_parent)                   # sorted parent ids
            layer_nid[hop - 1] = src_nodes
            col = torch.searchsorted(src_nodes, src_parent).to(torch.int32)
            rowptr = torch.zeros(dst_nodes.shape[0] + 1, dtype=torch.int64, device=dst_nodes.device)
            rowptr[1:] = torch.cumsum(deg, 0)
            blocks[hop - 1] = Block(rowptr, col, g.in_weight[eid], int(src_nodes.shape[0]), int(dst_nodes.shape[0]))
            eids[hop - 1] = eid
        return NodeFlow(g, layer_nid, blocks, eids)

    def __iter__(self):
        seeds = self.seed_nodes
        if self.shuffle:
            perm = torch.randperm(seeds.shape[0], generator=self.generator,
                                  device=self.generator.device if self.generator is not None else "cpu")
            seeds = seeds[perm.to(seeds.device)]
        for s in range(0, seeds.shape[0], self.batch_size):
            yield self.build(seeds[s:s + self.batch_size])


class FullGraphFlow:
    """All cells (or a subset ``seeds``, as cell indices 0..C-1) served by one layer-wise pass over
    the whole bipartite graph: the closed form of SURVEY §8a.  Numerically this is what the
    reference computes for those seeds with full-neighbour NodeFlows, without re-deriving the
    shared lower layers once per 500-seed batch."""

    def __init__(self, graph: BipartiteGraph, features: torch.Tensor, seeds: Optional[torch.Tensor] = None,
                 cells_ready: Optional[torch.cuda.Event] = None):
        if features.shape[0] != graph.num_genes + graph.num_cells:
            raise ValueError("features must have one row per gene then one per cell")
        self.graph = graph
        self.features = features
        self.seeds = seeds
        # set by FullGraphTrainer when the CELL rows of ``features`` are still arriving from the host on a
        # copy stream: the first cell<-gene pass only needs the gene rows, so it runs under the copy and
        # the event is waited on just before the first read of a cell row
        self.cells_ready = cells_ready

    def layer_parent_nid(self, i):
        if i not in (-1,):
            raise NotImplementedError("FullGraphFlow only exposes the seed layer")
        c = torch.arange(self.graph.num_cells, device=self.graph.device) if self.seeds is None else self.seeds
        return c + self.graph.num_genes
