"""``NodeFlow`` restatement (DGL 0.4 ``dgl.nodeflow.NodeFlow``), CPU PyTorch only.

A NodeFlow with ``L`` hops has ``L+1`` node layers (layer ``L`` = seeds) and ``L`` blocks; block
``i`` holds the sampled in-edges of layer ``i+1`` nodes, with sources in layer ``i``.
"""
import torch

from .graph import EdgeBatch, NodeBatch


class _LayerView:
    def __init__(self):
        self.data = {}


class _BlockView:
    def __init__(self):
        self.data = {}


class NodeFlow:
    def __init__(self, parent, layer_nids, block_edges):
        """layer_nids[i]: int64 parent node ids of layer i.
        block_edges[i]: (src_local, dst_local, parent_eid) int64 tensors for block i."""
        self._parent = parent
        self._layer_nids = layer_nids
        self._block_edges = block_edges
        self.layers = [_LayerView() for _ in layer_nids]
        self.blocks = [_BlockView() for _ in block_edges]

    @property
    def num_layers(self):
        return len(self._layer_nids)

    @property
    def num_blocks(self):
        return len(self._block_edges)

    def layer_size(self, i):
        return int(self._layer_nids[i].shape[0])

    def block_size(self, i):
        return int(self._block_edges[i][0].shape[0])

    def layer_parent_nid(self, i):
        return self._layer_nids[i]

    def block_parent_eid(self, i):
        return self._block_edges[i][2]

    def copy_from_parent(self, node_embed_names=None, edge_embed_names=None, ctx=None):
        for i, nid in enumerate(self._layer_nids):
            for key, col in self._parent.ndata.items():
                self.layers[i].data[key] = col[nid]
        for i, (_, _, eid) in enumerate(self._block_edges):
            for key, col in self._parent.edata.items():
                self.blocks[i].data[key] = col[eid]

    def block_compute(self, block_id, message_func, reduce_func, apply_node_func=None):
        src, dst, _ = self._block_edges[block_id]
        src_layer, dst_layer = self.layers[block_id], self.layers[block_id + 1]
        n_dst = self.layer_size(block_id + 1)
        edges = EdgeBatch({k: v[src] for k, v in src_layer.data.items()},
                          {k: v[dst] for k, v in dst_layer.data.items()},
                          dict(self.blocks[block_id].data))
        msgs = message_func(edges)
        m = msgs[reduce_func.msg_field]
        red = torch.zeros((n_dst,) + tuple(m.shape[1:]), dtype=m.dtype, device=m.device)
        red = red.index_add(0, dst, m)
        if reduce_func.name == "mean":
            deg = torch.bincount(dst, minlength=n_dst).clamp(min=1).to(m.dtype)
            red = red / deg.reshape((-1,) + (1,) * (m.dim() - 1))
        dst_layer.data[reduce_func.out_field] = red
        if apply_node_func is not None:
            out = apply_node_func(NodeBatch(dst_layer.data))
            dst_layer.data.update(out)
