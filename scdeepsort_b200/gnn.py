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

    use_tensor_cores = True     # tcgen05 kernel (bf16x3 split, fp32-grade accuracy) when the shape allows

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


class _CellAggregate(torch.autograd.Function):
    """neigh_c = s_c·[Σ_g α_g·w_{g→c}·h_g + α_{G+1}·h_c]  for every cell (SURVEY §8a closed form)."""

    @staticmethod
    def forward(ctx, hg, hc, alpha, graph: BipartiteGraph, algo, cells_ready=None):
        g = graph.num_genes
        a = alpha.reshape(-1)
        hs = hg * a[:g, None]                               # α folded into the (small) gene table
        if cells_ready is None:
            out, _, _ = ops.spmm(graph.cell_csr, hs, dscale=graph.mean_c * graph.norm_c,
                                 selfcoef=graph.mean_c * a[g + 1], hself=hc, algo=algo)
        else:
            # hc is still being copied from the host: the gene->cell sum needs only the gene table, so it runs
            # under the copy and the self-loop term is added once the rows have landed
            out, _, _ = ops.spmm(graph.cell_csr, hs, dscale=graph.mean_c * graph.norm_c, algo=algo)
            torch.cuda.current_stream().wait_event(cells_ready)
            out.addcmul_(hc, (graph.mean_c * a[g + 1])[:, None])
        ctx.save_for_backward(hg, hc, a)
        ctx.graph, ctx.algo, ctx.alpha_shape = graph, algo, alpha.shape
        return out

    @staticmethod
    def backward(ctx, dn):
        hg, hc, a = ctx.saved_tensors
        graph, g = ctx.graph, ctx.graph.num_genes
        need_hg, need_hc, need_a = ctx.needs_input_grad[:3]
        dhg = dhc = da = None
        if need_hg or need_a:
            # T_g = Σ_c x_cg·(s_c·norm_c·dn_c): the transposed pass; dα_g = <h_g, T_g> fused as a row-dot
            dsrc = dn * (graph.mean_c * graph.norm_c)[:, None]
            t, _, dot = ops.spmm(graph.transpose_of_cell_csr(), dsrc, q=hg, want_dot=need_a, algo=ctx.algo)
            if need_hg:
                dhg = t * a[:g, None]
        if need_hc:
            dhc = dn * (graph.mean_c * a[g + 1])[:, None]
        if need_a:
            da = torch.zeros_like(a)
            da[:g] = dot
            da[g + 1] = ((hc * dn).sum(dim=1) * graph.mean_c).sum()
            da = da.reshape(ctx.alpha_shape)
        return dhg, dhc, da, None, None, None


class _GeneAggregate(torch.autograd.Function):
    """neigh_g = s_g·[α_g·Σ_c w_{c→g}·h_c + α_G·h_g]  for every gene; only support cells send."""

    @staticmethod
    def forward(ctx, hg, hc_support, alpha, graph: BipartiteGraph, algo):
        g = graph.num_genes
        a = alpha.reshape(-1)
        need_raw = ctx.needs_input_grad[2]
        out, raw, _ = ops.spmm(graph.gene_csr, hc_support, dscale=graph.mean_g * graph.norm_g * a[:g],
                               selfcoef=graph.mean_g * a[g], hself=hg, want_raw=need_raw, algo=algo)
        ctx.save_for_backward(hg, a, raw)
        ctx.graph, ctx.algo, ctx.alpha_shape = graph, algo, alpha.shape
        return out

    @staticmethod
    def backward(ctx, dn):
        hg, a, raw = ctx.saved_tensors
        graph, g = ctx.graph, ctx.graph.num_genes
        need_hg, need_hc, need_a = ctx.needs_input_grad[:3]
        dhg = dhc = da = None
        if need_hg:
            dhg = dn * (graph.mean_g * a[g])[:, None]
        if need_hc:
            cs = graph.cell_csr
            ns = graph.num_support
            sup = cs if ns == graph.num_cells else Csr(cs.rowptr[:ns + 1], cs.col, cs.x, cs.n_src, ns, cs.col_bits, None)
            dsrc = dn * (graph.mean_g * graph.norm_g * a[:g])[:, None]
            dhc, _, _ = ops.spmm(sup, dsrc, algo=ctx.algo)
        if need_a:
            da = torch.zeros_like(a)
            da[:g] = (raw * dn).sum(dim=1) * graph.mean_g * graph.norm_g
            da[g] = ((hg * dn).sum(dim=1) * graph.mean_g).sum()
            da = da.reshape(ctx.alpha_shape)
        return dhg, dhc, da, None, None


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

    # -- mini-batch path: any NodeFlow -----------------------------------------------------
    def _forward_nodeflow(self, nf: NodeFlow):
        if "features" not in nf.layers[0].data:
            raise RuntimeError("call nf.copy_from_parent() before the forward pass (train.py:79)")
        dev = self.alpha.device
        if nf.layers[0].data["features"].device != dev:
            nf = nf.to(dev)
        h = nf.layers[0].data["features"]
        for i, layer in enumerate(self.layers):
            if self.dropout:
                h = self.dropout(h)          # on node features, before aggregation (models/gnn.py:62-64)
            neigh = ops.block_aggregate(h, self.alpha, nf.blocks[i], nf.layers[i].data["id"].reshape(-1),
                                        nf.layers[i + 1].data["id"].reshape(-1), self.gene_num)
            h = layer(neigh)
        return self._classify(h)

    def _classify(self, h):
        """Final linear (models/gnn.py:67), on the tensor-core kernel when the shape allows."""
        fc = self.linear
        if NodeUpdate.use_tensor_cores and h.is_cuda and dense.tc_supported(fc.in_features, fc.out_features):
            return dense.linear_relu(h, fc.weight, fc.bias, relu=False)
        return fc(h)

    # -- throughput path: whole bipartite graph, layer by layer ----------------------------
    def _forward_full(self, flow: FullGraphFlow):
        graph = flow.graph
        g, ns = graph.num_genes, graph.num_support
        h = flow.features
        ready = flow.cells_ready
        for i, layer in enumerate(self.layers):
            if self.dropout:
                if ready is not None:
                    torch.cuda.current_stream().wait_event(ready)
                    ready = None
                h = self.dropout(h)
            hg, hc = h[:g], h[g:]
            last = i == self.n_layers - 1
            neigh_c = _CellAggregate.apply(hg, hc, self.alpha, graph, self.spmm_algo, ready)
            ready = None
            if last:
                if flow.seeds is not None:
                    neigh_c = neigh_c[flow.seeds]
                h = layer(neigh_c)
            else:
                neigh_g = _GeneAggregate.apply(hg, hc[:ns], self.alpha, graph, self.spmm_algo)
                h = layer(torch.cat([neigh_g, neigh_c], dim=0))
        return self._classify(h)

    def forward(self, nf):
        if not self.alpha.is_cuda:
            raise RuntimeError("scdeepsort_b200.GNN runs on CUDA only (no CPU fallback): call .to('cuda')")
        if isinstance(nf, FullGraphFlow):
            return self._forward_full(nf)
        return self._forward_nodeflow(nf)


def predict_labels(logits: torch.Tensor, unsure_rate: float):
    """softmax → argmax with the reference's 'unsure' rule (train.py:106-113, predict.py:77-87):
    returns class indices, -1 where max prob < unsure_rate / num_classes.  Vectorised on device."""
    prob = F.softmax(logits, dim=1)
    max_prob, arg = prob.max(dim=1)
    return torch.where(max_prob < unsure_rate / logits.shape[1], torch.full_like(arg, -1), arg)
