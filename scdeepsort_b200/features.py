"""Node features of the graph contract without the dense ``[C, G]`` array.

The reference builds ``ndata['features']`` as ``cat[PCA(dense_dim).fit_transform(Xᵀ); (X / (rowsum + 1e-6)) · gene_feat]``
(/root/reference/utils/preprocess_internal.py:183-202, utils/preprocess.py:192-208) from ``vstack(...).toarray()`` — 60 GB at
atlas scale.  Both pieces are products of the sparse expression matrix with a thin dense matrix, i.e. the aggregation
kernels themselves:

* ``cell_features``      one cell←gene pass of ``wsage_spmm`` with ``dscale = 1 / (rowsum + 1e-6)``;
* ``pca_gene_features``  randomized range finder (Halko et al. 2011; what sklearn's ``PCA`` runs at these sizes) on the
                         centred ``A = Xᵀ - 1 μᵀ`` using only ``A·M = XᵀM - 1(μᵀM)`` (gene←cell pass) and
                         ``Aᵀ·M = XM - μ(1ᵀM)`` (cell←gene pass), QR / a ``k×k`` eigen-problem on the device.

PCA is input preparation, not part of the parity contract: components are determined up to the usual randomized-SVD
accuracy and sign; the sign convention follows sklearn's ``svd_flip`` (largest entry of each right singular vector positive).
"""
import torch

from .graph import BipartiteGraph
from .ops import spmm


def cell_features(graph: BipartiteGraph, gene_feat: torch.Tensor) -> torch.Tensor:
    """(X / (rowsum + 1e-6)) · gene_feat for every cell of ``graph`` (support then test cells)."""
    rowsum = _row_sums(graph.cell_csr, graph.device, gene_feat.shape[1])
    out, _, _ = spmm(graph.cell_csr, gene_feat.contiguous(), dscale=1.0 / (rowsum + 1e-6))
    return out


def _row_sums(csr, device, dim_hint):
    ones = torch.ones(csr.n_src, 4, device=device, dtype=torch.float32)
    s, _, _ = spmm(csr, ones)
    return s[:, 0].contiguous()


def pca_gene_features(graph: BipartiteGraph, n_components: int, seed: int = 10086, n_iter: int = 6, oversample: int = 10) -> torch.Tensor:
    """``PCA(n_components, random_state=seed).fit_transform(X_supportᵀ)``: ``[G, n_components]`` fp32 on the device.
    ``graph`` must hold the RAW expression CSRs (``BipartiteGraph.from_expression``); only support cells take part
    (utils/preprocess.py:196: ``sparse_feat[:support_num]``)."""
    dev = graph.device
    g, c = graph.num_genes, graph.num_support
    k = min(n_components, g, c)
    kk = max(4, min((k + oversample + 3) // 4 * 4, min(g, c) // 4 * 4))   # multiple of 4: 16-byte rows for the kernels
    k = min(k, kk)
    gen = torch.Generator(device=dev).manual_seed(int(seed))
    sup = graph.support_cell_csr()
    mu = _row_sums(sup, dev, kk) / float(g)                              # mean of every cell's column of A over the genes

    def a_times(m):                 # A·m,  m [C, kk] -> [G, kk]
        y, _, _ = spmm(graph.gene_csr, m.contiguous())
        return y - (mu[None, :] @ m)

    def at_times(m):                # Aᵀ·m, m [G, kk] -> [C, kk]
        z, _, _ = spmm(sup, m.contiguous())
        return z - mu[:, None] * m.sum(dim=0, keepdim=True)

    y = a_times(torch.randn(c, kk, device=dev, generator=gen))
    for _ in range(n_iter):
        q, _ = torch.linalg.qr(y)
        q2, _ = torch.linalg.qr(at_times(q))
        y = a_times(q2)
    q, _ = torch.linalg.qr(y)                                            # [G, kk] orthonormal basis of the range of A
    bt = at_times(q)                                                     # Bᵀ = AᵀQ  [C, kk]
    evals, evecs = torch.linalg.eigh((bt.t().double() @ bt.double()))    # B Bᵀ = U_B S² U_Bᵀ
    order = torch.argsort(evals, descending=True)[:k]
    s = evals[order].clamp(min=0).sqrt()
    ub = evecs[:, order]
    v = (bt.double() @ ub) / s.clamp(min=1e-30)[None, :]                 # right singular vectors [C, k]
    sign = torch.sign(v[v.abs().argmax(dim=0), torch.arange(k, device=dev)])
    sign = torch.where(sign == 0, torch.ones_like(sign), sign)
    feat = (q.double() @ ub) * (s * sign)[None, :]                       # U·S with sklearn's sign convention
    out = torch.zeros(g, n_components, device=dev, dtype=torch.float32)
    out[:, :k] = feat.float()
    return out
