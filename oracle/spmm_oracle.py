"""Optimised CPU variant of the oracle.  TEST / BENCH INFRASTRUCTURE ONLY (never imported by scdeepsort_b200).

`gnn_oracle.forward` is the literal restatement of /root/reference/models/gnn.py:47-68: it materialises the
per-edge message tensor exactly as the reference does.  SURVEY §8(d) asks for a second CPU number that is not
a strawman: the same arithmetic in its closed form, with the two normalised adjacency matrices as
torch sparse CSR tensors and one sparse x dense product per direction per layer — what a CPU user who rewrote
the UDF would run.  Full-neighbour, every cell a seed, support cells only.

    cell c:  neigh_c = s_c [ sum_g alpha_g w_{g->c} h_g + alpha_{G+1} h_c ]     w_{g->c} = x_cg deg_c / sum_g x_cg
    gene g:  neigh_g = s_g [ alpha_g sum_c w_{c->g} h_c + alpha_G h_g ]         w_{c->g} = x_cg deg_g / sum_c x_cg
    h' = relu(W neigh + b);  logits = W_o h_cells + b_o                          s_v = 1 / (deg_v + 1)
"""
import numpy as np
import scipy.sparse as sp
import torch


class SpmmGraph:
    def __init__(self, x: sp.csr_matrix, dtype=torch.float32):
        x = sp.csr_matrix(x).astype(np.float64)
        x.sort_indices()
        self.num_cells, self.num_genes = x.shape
        deg_c = np.diff(x.indptr).astype(np.float64)
        xt = x.T.tocsr()
        xt.sort_indices()
        deg_g = np.diff(xt.indptr).astype(np.float64)
        rs = np.asarray(x.sum(axis=1)).ravel()
        cs = np.asarray(x.sum(axis=0)).ravel()
        with np.errstate(divide="ignore", invalid="ignore"):
            norm_c = np.where(deg_c > 0, deg_c / rs, 0.0)
            norm_g = np.where(deg_g > 0, deg_g / cs, 0.0)
        wc = sp.diags(norm_c) @ x               # [C, G]  w_{g->c}
        wg = sp.diags(norm_g) @ xt              # [G, C]  w_{c->g}

        def csr(m):
            m = sp.csr_matrix(m)
            return torch.sparse_csr_tensor(torch.from_numpy(m.indptr.astype(np.int64)), torch.from_numpy(m.indices.astype(np.int64)),
                                           torch.from_numpy(m.data).to(dtype), size=m.shape)
        self.wc, self.wg = csr(wc), csr(wg)
        self.mean_c = torch.from_numpy(1.0 / (deg_c + 1)).to(dtype)[:, None]
        self.mean_g = torch.from_numpy(1.0 / (deg_g + 1)).to(dtype)[:, None]


def forward(params: dict, graph: SpmmGraph, features: torch.Tensor, n_layers: int, relu_masks=None, hidden_out=None) -> torch.Tensor:
    """relu_masks[i] (optional, 0/1, shape of layer i's output) replaces that layer's ReLU — see gnn_oracle.forward."""
    act = (lambda y, i: torch.relu(y)) if relu_masks is None else (lambda y, i: y * relu_masks[i].to(y.dtype))
    g = graph.num_genes
    a = params["alpha"].reshape(-1, 1)
    h = features
    for i in range(n_layers):
        hg, hc = h[:g], h[g:]
        neigh_c = graph.mean_c * (torch.sparse.mm(graph.wc, hg * a[:g]) + a[g + 1] * hc)
        w, b = params[f"layers.{i}.fc_neigh.weight"], params[f"layers.{i}.fc_neigh.bias"]
        if i == n_layers - 1:
            h = act(neigh_c @ w.t() + b, i)
        else:
            neigh_g = graph.mean_g * (a[:g] * torch.sparse.mm(graph.wg, hc) + a[g] * hg)
            h = act(torch.cat([neigh_g, neigh_c], dim=0) @ w.t() + b, i)
        if hidden_out is not None:
            hidden_out.append(h.detach())
    return h @ params["linear.weight"].t() + params["linear.bias"]
