"""``dgl.contrib.sampling.NeighborSampler`` restatement (DGL 0.4), CPU only.

Yields one :class:`NodeFlow` per ``batch_size`` seeds.  ``expand_factor`` ≥ in-degree keeps
every in-edge (the reference's full-neighbour mode, train.py:37-40,96; predict.py:41,66);
otherwise ``expand_factor`` in-edges are drawn uniformly without replacement per node per
hop.  ``num_workers`` only affects DGL's prefetch threading and is ignored here.
"""
import numpy as np
import torch

from ..graph import _as_index
from ..nodeflow import NodeFlow


class NeighborSampler:
    def __init__(self, g, batch_size, expand_factor=None, num_hops=1, neighbor_type='in',
                 transition_prob=None, seed_nodes=None, shuffle=False, num_workers=1,
                 prefetch=False, add_self_loop=False):
        assert neighbor_type == 'in', "the reference only samples in-neighbours"
        assert transition_prob is None and not add_self_loop
        self.g = g
        self.batch_size = int(batch_size)
        self.expand_factor = g.number_of_nodes() if expand_factor is None else int(expand_factor)
        self.num_hops = int(num_hops)
        self.seed_nodes = g.nodes() if seed_nodes is None else _as_index(seed_nodes)
        self.shuffle = shuffle

    def _in_edges_sampled(self, nodes):
        ptr, order = self.g._build_in_csr()
        eids = []
        for v in nodes.tolist():
            e = order[ptr[v]:ptr[v + 1]]
            if e.shape[0] > self.expand_factor:
                pick = np.random.choice(e.shape[0], self.expand_factor, replace=False)
                e = e[torch.as_tensor(np.sort(pick))]
            eids.append(e)
        return torch.cat(eids) if eids else torch.zeros(0, dtype=torch.int64)

    def _build(self, seeds):
        layer_nids = [None] * (self.num_hops + 1)
        block_edges = [None] * self.num_hops
        layer_nids[self.num_hops] = seeds
        for hop in range(self.num_hops, 0, -1):
            dst_nodes = layer_nids[hop]
            eid = self._in_edges_sampled(dst_nodes)
            src_parent, dst_parent = self.g._src[eid], self.g._dst[eid]
            src_nodes = torch.unique(src_parent)  # sorted parent ids
            layer_nids[hop - 1] = src_nodes
            src_local = torch.searchsorted(src_nodes, src_parent)
            # dst nodes keep the order of the upper layer (seed order for the last layer)
            lookup = torch.full((self.g.number_of_nodes(),), -1, dtype=torch.int64)
            lookup[dst_nodes] = torch.arange(dst_nodes.shape[0])
            dst_local = lookup[dst_parent]
            block_edges[hop - 1] = (src_local, dst_local, eid)
        return NodeFlow(self.g, layer_nids, block_edges)

    def __iter__(self):
        seeds = self.seed_nodes
        if self.shuffle:
            seeds = seeds[torch.as_tensor(np.random.permutation(seeds.shape[0]))]
        for start in range(0, seeds.shape[0], self.batch_size):
            yield self._build(seeds[start:start + self.batch_size])
