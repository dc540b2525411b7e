"""Feature construction on the aggregation kernels (scdeepsort_b200/features.py): no dense [C, G] array.
PCA is input preparation (components are defined up to sign and randomized-SVD accuracy), so it is compared with sklearn's
exact solver through sign-free quantities; the cell features are plain arithmetic and are compared with the oracle."""
import numpy as np
import pytest
import scipy.sparse as sp
import torch

import scdeepsort_b200 as sd
from oracle import graph_oracle
from scdeepsort_b200.features import cell_features, pca_gene_features
from scds_helpers import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _expression(c, g, seed, rank=10):
    """Sparse, non-negative, with ``rank`` well-separated leading directions (geometrically decaying strengths) over noise."""
    rng = np.random.RandomState(seed)
    u = np.abs(rng.normal(0, 1, (c, rank))) * (0.7 ** np.arange(rank))[None, :]
    low = u @ np.abs(rng.normal(0, 1, (rank, g)))
    dense = np.where(rng.rand(c, g) < 0.3, low + 0.02 * rng.rand(c, g) + 0.01, 0.0)
    return sp.csr_matrix(dense.astype(np.float32))


@pytest.mark.parametrize("c,g,k,with_test", [(600, 900, 40, False), (300, 500, 64, True), (90, 70, 400, False)])
def test_pca_gene_features_match_sklearn(c, g, k, with_test):
    """Against sklearn's exact solver.  A randomized range finder resolves the directions that stand out of the noise floor
    (here the first 5, each ≥ 1.3x the next): singular values within 1e-3 of the largest, identical embedding up to sign
    (Gram matrices within 1 %); the flat tail is only checked through the captured variance (within 2 %)."""
    from sklearn.decomposition import PCA
    x = _expression(c, g, c + g)
    xt = _expression(50, g, 7) if with_test else None
    bg = sd.BipartiteGraph.from_expression(x, xt, device=DEV)
    got = pca_gene_features(bg, k, seed=3).cpu().double()
    assert got.shape == (g, k)
    k_eff = min(k, c, g)
    ref = torch.from_numpy(PCA(k_eff, svd_solver="full").fit_transform(np.asarray(x.todense(), dtype=np.float64).T))   # support cells only
    sv_got, sv_ref = got[:, :k_eff].norm(dim=0), ref.norm(dim=0)
    top = 5
    assert float((sv_got[:top] - sv_ref[:top]).abs().max() / sv_ref[0]) < 1e-3
    gram_got, gram_ref = got[:, :top] @ got[:, :top].t(), ref[:, :top] @ ref[:, :top].t()
    assert float((gram_got - gram_ref).norm() / gram_ref.norm()) < 1e-2
    n_cmp = min(k_eff, int((sv_got > 0).sum()))
    assert abs(float(sv_got[:n_cmp].square().sum() / sv_ref[:n_cmp].square().sum()) - 1) < 2e-2
    assert bool((got[:, k_eff:] == 0).all())                        # tiny inputs: zero-padded to the requested width
    # sklearn's sign convention on the leading component (largest |v| entry positive <=> same sign of the scores)
    assert float((got[:, 0] * ref[:, 0]).sum()) > 0
    assert torch.equal(got, pca_gene_features(bg, k, seed=3).cpu().double())       # seeded: reproducible


def test_cell_features_match_oracle():
    x, xt = _expression(400, 300, 1), _expression(30, 300, 2)
    bg = sd.BipartiteGraph.from_expression(x, xt, device=DEV)
    gf = torch.randn(300, 48, generator=torch.Generator().manual_seed(0))
    got = cell_features(bg, gf.to(DEV)).cpu()
    ref = graph_oracle.make_features(sp.vstack([x, xt]).tocsr(), gf.numpy())[300:]
    assert rel_err(got, ref) < 1e-5
