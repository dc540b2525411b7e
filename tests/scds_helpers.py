"""Shared helpers for tests: golden → oracle structures."""
import numpy as np
import scipy.sparse as sp
import torch

from oracle.graph_oracle import OracleGraph


def golden_state(z, tag):
    pre = tag + "/"
    keys = ("alpha", "linear.weight", "linear.bias")
    out = {}
    for k in z.files:
        if k.startswith(pre):
            name = k[len(pre):]
            if name in keys or name.startswith("layers."):
                out[name] = torch.from_numpy(z[k])
    return out


def golden_grads(z, tag):
    pre = tag + "/grad/"
    return {k[len(pre):]: torch.from_numpy(z[k]) for k in z.files if k.startswith(pre)}


def golden_graph(z, prefix="graph/", num_genes=None):
    node_id = torch.from_numpy(z[prefix + "node_id"])
    g = int((node_id >= 0).sum()) if num_genes is None else num_genes
    return OracleGraph(g, node_id.shape[0] - g, torch.from_numpy(z[prefix + "src"]),
                       torch.from_numpy(z[prefix + "dst"]), torch.from_numpy(z[prefix + "weight"]),
                       node_id, torch.from_numpy(z[prefix + "features"]))


def golden_csr(z, pre="x"):
    return sp.csr_matrix((z[pre + "_data"], z[pre + "_indices"], z[pre + "_indptr"]), shape=tuple(z[pre + "_shape"]))


def rel_err(a, b):
    """max|a-b| / max|b| — logits can be ≈0 so element-wise relative error is ill-posed."""
    a = torch.as_tensor(np.asarray(a), dtype=torch.float64)
    b = torch.as_tensor(np.asarray(b), dtype=torch.float64)
    return float((a - b).abs().max() / b.abs().max().clamp(min=1e-30))


def dense_block_matrix(block):
    """[cells, gene_slots] fp32 values a ``DenseBlock`` stands for: planes
    ``plane[cell // 128][slot // 32][cell % 128][slot % 32]`` decoded as (hi + lo) / x_scale (include/wsage.h)."""
    nb = block.slots_pad // 32
    n_tiles = (block.cells + 127) // 128
    f16 = block.lo is not None
    dt = torch.float16 if f16 else torch.bfloat16

    def plane(t):
        v = t.cpu().view(dt).to(torch.float64).view(n_tiles, nb, 128, 32)
        return v.permute(0, 2, 1, 3).reshape(n_tiles * 128, nb * 32)[:block.cells, :block.gene_slots]

    m = plane(block.hi)
    if f16:
        m = m + plane(block.lo)
    return (m / block.x_scale).to(torch.float32)


def csr_dense_matrix(csr):
    """[n_dst, n_src] matrix of the entries ``csr.dense`` holds for this CSR (zeros elsewhere)."""
    d = csr.dense
    out = torch.zeros(csr.n_dst, csr.n_src)
    if d is None:
        return out
    m = dense_block_matrix(d)
    ids = d.gene_ids.cpu().to(torch.int64)
    if csr.dense_side == 0:         # rows = cells, block sources = genes
        out[:, ids] = m[:csr.n_dst]
    else:                            # rows = genes, block sources = cells
        out[ids, :] = m[:csr.n_src].t()
    return out


def recipe_features(x_all, num_genes, dense_dim, seed):
    """The seeded stand-in for the PCA features of tests/golden/adipose.npz (same function as in
    oracle/gen_golden.py): gene rows ~ N(0, 0.66²) from a torch CPU generator; cell rows = (x / (rowsum + 1e-6)) ·
    gene_feat in float64, as /root/reference/utils/preprocess_internal.py:190-196."""
    gen = torch.Generator().manual_seed(int(seed))
    gene_feat = (torch.randn(num_genes, dense_dim, generator=gen) * 0.66).numpy().astype(np.float64)
    dense = np.asarray(x_all.todense(), dtype=np.float64)
    dense = dense / (np.sum(dense, axis=1, keepdims=True) + 1e-6)
    cell_feat = dense.dot(gene_feat)
    return torch.cat([torch.from_numpy(gene_feat), torch.from_numpy(cell_feat)], dim=0).type(torch.float)


def seeded_state(z, tag, dense_dim, num_labels, num_genes):
    """State dict of golden model ``tag``, regenerated from its seed exactly as oracle/gen_golden.py:_make_model built
    it with the reference's GNN class (scdeepsort_b200.GNN has the same constructor, so it draws the same numbers);
    the stored per-parameter sums guard against RNG drift."""
    import torch.nn.functional as F
    import scdeepsort_b200 as sd
    torch.manual_seed(int(z[f"{tag}/seed"]))
    m = sd.GNN(in_feats=dense_dim, n_hidden=int(z[f"{tag}/hidden"]), n_classes=num_labels, n_layers=int(z[f"{tag}/n_layers"]),
               gene_num=num_genes, activation=F.relu, dropout=0.0)
    with torch.no_grad():
        m.alpha.copy_(0.5 + torch.rand(m.alpha.shape))
        m.linear.bias.copy_(torch.rand(m.linear.bias.shape) - 0.5)
    state = {k: v.detach().clone() for k, v in m.state_dict().items()}
    sums = np.array([float(v.double().sum()) for v in state.values()])
    assert np.allclose(sums, z[f"{tag}/param_sums"], rtol=1e-9, atol=1e-9), "seeded weights differ from the golden run's"
    return state


def adipose_inputs(z, with_test=False):
    """(x_support, x_test or None, features) of tests/golden/adipose.npz."""
    x = golden_csr(z)
    xt = golden_csr(z, "xt") if with_test else None
    x_all = x if xt is None else sp.vstack([x, xt]).tocsr()
    return x, xt, recipe_features(x_all, int(z["num_genes"]), int(z["dense_dim"]), int(z["feat_seed"]))


def sampled_grad_err(grad, z, tag, name, step=53):
    """Gradient check against a golden entry stored whole (small tensors) or as a strided sample + norm."""
    g = torch.as_tensor(grad).detach().cpu()
    if f"{tag}/grad/{name}" in z.files:
        return rel_err(g, z[f"{tag}/grad/{name}"])
    ref = z[f"{tag}/grad_sample/{name}"]
    e1 = float((g.reshape(-1)[::step].double() - torch.from_numpy(ref).double()).abs().max() / g.double().abs().max())
    e2 = abs(float(g.double().norm()) - float(z[f"{tag}/grad_norm/{name}"])) / float(z[f"{tag}/grad_norm/{name}"])
    return max(e1, e2)


class ReluMaskCapture:
    """Records the 0/1 ReLU decisions of every NodeUpdate of a ``scdeepsort_b200.GNN`` during one forward pass.

    Gradients are discontinuous where a pre-activation crosses zero: two correct fp32 implementations whose forward
    values differ by 1e-6 take different branches for a few entries out of millions, and each such entry moves a weight
    gradient by a whole term of its sum.  Gradient parity is therefore checked against the fp64 oracle evaluated under
    the product's own decisions (``relu_masks=``), and ``disagreement`` bounds how many decisions differ from the
    oracle's own ReLU (they must be a negligible share, all at |pre-activation| ~ rounding error)."""

    def __init__(self, model):
        self.masks = []
        self._handles = [layer.register_forward_hook(lambda mod, inp, out: self.masks.append((out > 0).detach().cpu()))
                         for layer in model.layers]

    def close(self):
        for h in self._handles:
            h.remove()

    @staticmethod
    def disagreement(masks, oracle_hidden):
        """Largest share of entries, over the layers, where the product's decision differs from ``oracle_hidden[i] > 0``."""
        return max(float((m != (h > 0)).float().mean()) for m, h in zip(masks, oracle_hidden))
