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
