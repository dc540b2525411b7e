"""``GNN`` / ``NodeUpdate``: host-side mirror of /root/reference/models/gnn.py.

Same constructor, same ``state_dict`` keys and shapes (``alpha [G+2,1]``,
``layers.{i}.fc_neigh.{weight,bias}``, ``linear.{weight,bias}``; models/gnn.py:13,37-44), same
initialisation (xavier_uniform, gain √2; α = 1; models/gnn.py:16,43,45), so checkpoints written
by the reference's ``Trainer.save_model`` (train.py:117-123) load unchanged.  ``forward`` takes
our ``NodeFlow`` (any sampled / full-neighbour mini-batch) or a ``FullGraphFlow`` and returns
``logits[B, n_classes]`` with rows ordered as ``nf.layer_parent_nid(-1)`` (train.py:80-82).

The DGL UDF protocol (message_func / fn.mean / apply) is replaced wholesale by fused kernels in
``libwsage.so``; nothing here has a CPU code path.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import dense, ops
from .graph import BipartiteGraph
from .nodeflow import FullGraphFlow, NodeFlow
from .ops import Csr


class NodeUpdate(nn.Module):
    """models/gnn.py:10-25: ``activation(fc_neigh(neigh))`` then optional norm.  No self/neigh
    concatenation: the node's own state enters only through its self-loop edge."""

    def __init__(self, in_feats, out_feats, activation=None, norm=None):
        super().__init__()
        self.fc_neigh = nn.Linear(in_features=in_feats, out_features=out_feats)
        self.activation = activation
        self.norm = norm
        nn.init.xavier_uniform_(self.fc_neigh.weight, gain=nn.init.calculate_gain('relu'))

    use_tensor_cores = True     # tcgen05 kernel (tf32 hi/lo split, three products: fp32-grade accuracy) when the shape allows

    def forward(self, h_neigh):
        fc = self.fc_neigh
        relu = self.activation in (F.relu, torch.relu)
        if (self.use_tensor_cores and h_neigh.is_cuda and (relu or self.activation is None)
                and dense.tc_supported(fc.in_features, fc.out_features)):
            h_neigh = dense.linear_relu(h_neigh, fc.weight, fc.bias, relu=relu)     # Linear + bias + ReLU fused
        else:
            h_neigh = fc(h_neigh)
            if self.activation is not None:
                h_neigh = self.activation(h_neigh)
        if self.norm is not None:
            h_neigh = self.norm(h_neigh)
        return h_neigh


def _all_reduce_(t):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t)
    return t


class _LayerAggregate(torch.autograd.Function):
    """Both aggregations of one layer over the whole bipartite graph (SURVEY §8a closed form), written into ONE
    ``[G + C, D]`` buffer (``[C, D]`` for the last layer, which needs no gene states):

        neigh_c = s_c·[Σ_g α_g·w_{g→c}·h_g + α_{G+1}·h_c]          every cell
        neigh_g = s_g·[α_g·Σ_c w_{c→g}·h_c + α_G·h_g]               every gene; only support cells send

    i.e. message_func + fn.mean of /root/reference/models/gnn.py:47-56,65 for every destination at once.  Backward
    runs the two transposed passes with the gradient scalings, the self-loop terms and dα_g = <h_g, T_g> fused into
    the kernels' epilogues.  ``sharded``: the graph holds this rank's cells; the raw gene sums (and, when the cell
    features need a gradient, their incoming gradient) are all-reduced."""

    @staticmethod
    def forward(ctx, h, alpha, graph: BipartiteGraph, algo, gene_too, cells_ready, gene_mask, cell_mask, sharded):
        g, ns, c = graph.num_genes, graph.num_support, graph.num_cells
        a = alpha.reshape(-1)
        hg, hc = h[:g], h[g:]
        if gene_mask is not None:
            hg = hg * gene_mask                             # dropout of the gene rows (models/gnn.py:62-63)
        off = g if gene_too else 0
        neigh = torch.empty(off + c, h.shape[1], device=h.device, dtype=torch.float32)
        alpha_g = a[:g].contiguous()                        # α enters as a per-source-row factor of the (small) gene table
        coef_c = graph.mean_c * a[g + 1]
        if cells_ready is None and cell_mask is None:
            ops.spmm(graph.cell_csr, hg, src_scale=alpha_g, dscale=graph.mean_c * graph.norm_c, selfcoef=coef_c, hself=hc,
                     out=neigh[off:], algo=algo)
        else:
            # hc is still being copied from the host: the gene->cell sum needs only the gene table, so it runs
            # under the copy and the self-loop term is added once the rows have landed
            ops.spmm(graph.cell_csr, hg, src_scale=alpha_g, dscale=graph.mean_c * graph.norm_c, out=neigh[off:], algo=algo)
            if cells_ready is not None:
                torch.cuda.current_stream().wait_event(cells_ready)
            if cell_mask is not None:
                hc = hc * cell_mask                         # dropout of the cell rows (models/gnn.py:62-63), after arrival
            neigh[off:].addcmul_(hc, coef_c[:, None])
        raw = None
        if gene_too:
            need_raw = ctx.needs_input_grad[1]
            dscale_g = graph.mean_g * graph.norm_g * a[:g]
            pg = getattr(graph, "peer_group", None) if sharded else None        # parallel.enable_peer_exchange
            if pg is not None and g * h.shape[1] > pg.max_elems:
                raise RuntimeError(f"the peer exchange buffers hold {pg.max_elems} values, this layer needs {g * h.shape[1]}")
            if pg is not None:
                # split-K slabs of this shard -> sum over the ranks -> scale + self loop: one kernel over peer memory
                d = graph.gene_csr.dense
                slabs = ops.dense16(d, 1, hc[:ns], n_src_cells=ns)
                raw = torch.empty(g, h.shape[1], device=h.device, dtype=torch.float32) if need_raw else None
                pg.reduce(slabs, g, slot_of_row=d.slot_of_gene, dscale=dscale_g, selfcoef=graph.mean_g * a[g], hself=hg,
                          out=neigh[:g], raw=raw)
            elif sharded:
                _, raw, _ = ops.spmm(graph.gene_csr, hc[:ns], want_out=False, want_raw=True, algo=algo)
                _all_reduce_(raw)
                torch.mul(raw, dscale_g[:, None], out=neigh[:g])
                neigh[:g].addcmul_(hg, (graph.mean_g * a[g])[:, None])
            else:
                _, raw, _ = ops.spmm(graph.gene_csr, hc[:ns], dscale=dscale_g, selfcoef=graph.mean_g * a[g], hself=hg,
                                     out=neigh[:g], want_raw=need_raw, algo=algo)
        ctx.save_for_backward(hg, hc, a, raw, gene_mask, cell_mask)
        ctx.graph, ctx.algo, ctx.alpha_shape, ctx.gene_too, ctx.sharded = graph, algo, alpha.shape, gene_too, sharded
        return neigh

    @staticmethod
    def backward(ctx, dn):
        hg, hc, a, raw, gene_mask, cell_mask = ctx.saved_tensors       # hg / hc: after their dropout masks, if any
        graph, gene_too = ctx.graph, ctx.gene_too
        g, ns, c = graph.num_genes, graph.num_support, graph.num_cells
        need_h, need_a = ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        off = g if gene_too else 0
        dn = dn.contiguous()
        dn_c = dn[off:]
        dn_g = dn[:g] if gene_too else None
        dh = torch.empty(g + c, dn.shape[1], device=dn.device, dtype=torch.float32) if need_h else None
        da = None
        coef_c = graph.mean_c * a[g + 1]
        dot = None
        if need_h or need_a:
            # T_g = Σ_c x_cg·(s_c·norm_c·dn_c): the transposed pass; dα_g = <h_g, T_g> is its row-dot epilogue and
            # dh_g = α_g·T_g + s_g·α_G·dn_g its scale / self epilogue
            kw = {}
            if need_h:
                kw = dict(dscale=a[:g].contiguous(), out=dh[:g])
                if gene_too:
                    kw.update(selfcoef=graph.mean_g * a[g], hself=dn_g)
            _, _, dot = ops.spmm(graph.transpose_of_cell_csr(), dn_c, src_scale=graph.mean_c * graph.norm_c, q=hg, want_dot=need_a,
                                 want_out=need_h, algo=ctx.algo, **kw)
        if need_h:
            if gene_too:
                # dh_c = X_sup·(s_g·norm_g·α_g·dn_g) + s_c·α_{G+1}·dn_c  (self term fused); test cells send nothing to genes
                scale_g = graph.mean_g * graph.norm_g * a[:g]
                src, scale = dn_g, scale_g
                if ctx.sharded:                             # backward of the forward all-reduce of the raw gene sums
                    src, scale = _all_reduce_(dn_g * scale_g[:, None]), None
                ops.spmm(graph.support_cell_csr(), src, src_scale=scale, selfcoef=coef_c[:ns].contiguous(), hself=dn_c[:ns],
                         out=dh[g:g + ns], algo=ctx.algo)
                if ns < c:
                    torch.mul(dn_c[ns:], coef_c[ns:, None], out=dh[g + ns:])
            else:
                torch.mul(dn_c, coef_c[:, None], out=dh[g:])
            if cell_mask is not None:
                dh[g:].mul_(cell_mask)
            if gene_mask is not None:
                dh[:g].mul_(gene_mask)
        if need_a:
            da = torch.zeros_like(a)
            da[:g] = dot
            da[g + 1] = (ops.rowdot(hc, dn_c) * graph.mean_c).sum()
            if gene_too:
                da[:g] += ops.rowdot(raw, dn_g) * graph.mean_g * graph.norm_g
                da[g] = (ops.rowdot(hg, dn_g) * graph.mean_g).sum()
            da = da.reshape(ctx.alpha_shape)
        return dh, da, None, None, None, None, None, None, None


def _is_dgl_block(b):
    return hasattr(b, "srcdata") and hasattr(b, "dstdata") and hasattr(b, "edges") and hasattr(b, "edata")


def blocks_to_flow(blocks, device):
    """DGL >= 0.5 message-flow blocks (what ``dgl.dataloading.NodeDataLoader`` yields as its third element) →
    the ``(features, [(ops.Block, src_id, dst_id)])`` form the kernels take.  Duck-typed: each block needs
    ``srcdata`` / ``dstdata`` (``'id'``, and ``'features'`` on the first block), ``edges()`` → (src, dst) local
    indices, ``edata['weight']`` and ``num_src_nodes()`` / ``num_dst_nodes()``.  No DGL import."""
    out = []
    for b in blocks:
        src, dst = b.edges()
        src = torch.as_tensor(src, dtype=torch.int64, device=device)
        dst = torch.as_tensor(dst, dtype=torch.int64, device=device)
        n_src, n_dst = int(b.num_src_nodes()), int(b.num_dst_nodes())
        w = torch.as_tensor(b.edata["weight"], dtype=torch.float32, device=device).reshape(-1)
        order = torch.sort(dst, stable=True).indices
        rowptr = torch.zeros(n_dst + 1, dtype=torch.int64, device=device)
        rowptr[1:] = torch.cumsum(torch.bincount(dst, minlength=n_dst), 0)
        blk = ops.Block(rowptr, src[order].to(torch.int32).contiguous(), w[order].contiguous(), n_src, n_dst)
        sid = torch.as_tensor(b.srcdata["id"], device=device).reshape(-1).to(torch.int32).contiguous()
        did = torch.as_tensor(b.dstdata["id"], device=device).reshape(-1).to(torch.int32).contiguous()
        out.append((blk, sid, did))
    feats = torch.as_tensor(blocks[0].srcdata["features"], dtype=torch.float32, device=device).contiguous()
    return feats, out


class GNN(nn.Module):
    def __init__(self, in_feats, n_hidden, n_classes, n_layers, gene_num, activation=None, norm=None, dropout=0.0):
        super().__init__()
        self.n_layers = n_layers
        self.gene_num = gene_num
        if dropout != 0:
            self.dropout = nn.Dropout(p=dropout)
        else:
            self.dropout = None
        self.layers = nn.ModuleList()
        self.layers.append(NodeUpdate(in_feats=in_feats, out_feats=n_hidden, activation=activation, norm=norm))
        for _ in range(n_layers - 1):
            self.layers.append(NodeUpdate(in_feats=n_hidden, out_feats=n_hidden, activation=activation, norm=norm))
        # [gene_num] is alpha of gene-gene, [gene_num+1] is alpha of cell-cell self loop
        self.alpha = nn.Parameter(torch.tensor([1] * (self.gene_num + 2), dtype=torch.float32).unsqueeze(-1))
        self.linear = nn.Linear(n_hidden, n_classes)
        nn.init.xavier_uniform_(self.linear.weight, gain=nn.init.calculate_gain('relu'))
        self.spmm_algo = 0     # wsage_spmm algo for the full-graph path (0 = auto)

    # -- mini-batch path: any NodeFlow, or a list of DGL >= 0.5 blocks -------------------------
    def _forward_blocks(self, h, blocks):
        for layer, (blk, src_id, dst_id) in zip(self.layers, blocks):
            if self.dropout:
                h = self.dropout(h)          # on node features, before aggregation (models/gnn.py:62-64)
            h = layer(ops.block_aggregate(h, self.alpha, blk, src_id, dst_id, self.gene_num))
        return self._classify(h)

    def _forward_nodeflow(self, nf: NodeFlow):
        if "features" not in nf.layers[0].data:
            raise RuntimeError("call nf.copy_from_parent() before the forward pass (train.py:79)")
        dev = self.alpha.device
        if nf.layers[0].data["features"].device != dev:
            nf = nf.to(dev)
        blocks = [(nf.blocks[i], nf.layers[i].data["id"].reshape(-1), nf.layers[i + 1].data["id"].reshape(-1))
                  for i in range(len(self.layers))]
        return self._forward_blocks(nf.layers[0].data["features"], blocks)

    def _classify(self, h):
        """Final linear (models/gnn.py:67), on the tensor-core kernel when the shape allows."""
        fc = self.linear
        if NodeUpdate.use_tensor_cores and h.is_cuda and dense.tc_supported(fc.in_features, fc.out_features):
            return dense.linear_relu(h, fc.weight, fc.bias, relu=False)
        return fc(h)

    # -- throughput path: whole bipartite graph, layer by layer ----------------------------
    def _forward_full(self, flow: FullGraphFlow):
        graph = flow.graph
        g = graph.num_genes
        h = flow.features
        ready = flow.cells_ready
        sharded = bool(getattr(flow, "sharded", False))
        for i, layer in enumerate(self.layers):
            gmask = mask = None
            if self.dropout:
                if self.training and (ready is not None or sharded):
                    # the masks do not depend on the data: drawn now, applied inside the aggregate — the gene rows at
                    # once (replicated state: the same mask on every rank when sharded), the cell rows once they
                    # have arrived from the host, so the first gene->cell pass still runs under the copy
                    gmask = self._dropout_mask((g, h.shape[1]), h.device, shared=sharded)
                    mask = self._dropout_mask((h.shape[0] - g, h.shape[1]), h.device, shared=False)
                else:
                    h = self.dropout(h)
            last = i == self.n_layers - 1
            neigh = _LayerAggregate.apply(h, self.alpha, graph, self.spmm_algo, not last, ready, gmask, mask, sharded)
            ready = None
            if last and flow.seeds is not None:
                neigh = neigh[flow.seeds]
            h = layer(neigh)
        return self._classify(h)

    def _dropout_mask(self, shape, device, shared):
        """Inverted-dropout mask (0 or 1/(1-p)).  ``shared``: from a generator seeded alike on every rank."""
        p = self.dropout.p
        gen = None
        if shared:
            gen = getattr(self, "_shared_gen", None)
            if gen is None:
                gen = self._shared_gen = torch.Generator(device=device).manual_seed(10086)
        return (torch.rand(shape, device=device, generator=gen) >= p).to(torch.float32) / (1.0 - p)

    def forward(self, nf):
        if not self.alpha.is_cuda:
            raise RuntimeError("scdeepsort_b200.GNN runs on CUDA only (no CPU fallback): call .to('cuda')")
        if isinstance(nf, FullGraphFlow):
            return self._forward_full(nf)
        if isinstance(nf, (list, tuple)) and len(nf) and all(_is_dgl_block(b) for b in nf):
            if len(nf) != len(self.layers):
                raise ValueError(f"expected {len(self.layers)} blocks, got {len(nf)}")
            feats, blocks = blocks_to_flow(nf, self.alpha.device)
            return self._forward_blocks(feats, blocks)
        return self._forward_nodeflow(nf)


def predict_labels(logits: torch.Tensor, unsure_rate: float):
    """softmax → argmax with the reference's 'unsure' rule (train.py:106-113, predict.py:77-87):
    returns class indices, -1 where max prob < unsure_rate / num_classes.  Vectorised on device."""
    prob = F.softmax(logits, dim=1)
    max_prob, arg = prob.max(dim=1)
    return torch.where(max_prob < unsure_rate / logits.shape[1], torch.full_like(arg, -1), arg)
