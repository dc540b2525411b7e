"""Synthetic expression matrices of the shape BASELINE.json's configs name.

There is no network for real atlases, so bench.py and the large parity/property tests use a
seeded generator with the statistics of the reference fixtures: Zipf-like gene popularity
(p_g ∝ rank^-0.8, ranks scattered over gene ids as in an alphabetical gene list), log-normal
cell depth (σ = 0.35), values ~ clip(N(3.0, 0.5), 0.05, 9).

Edge (c, g) exists iff ``hash32(c, g, seed) < min(1, depth_c · pop_g) · 2^32`` — a pure integer
test per pair, so the cell-major CSR and its gene-major transpose are generated independently,
each already sorted, without ever sorting or transposing 10^9 edges, and identically on CPU
and GPU (plain torch integer ops; plumbing, not the product).
"""
import math
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib
from .graph import BipartiteGraph, _balanced_row_perm
from .ops import Csr

_M32 = 0xFFFFFFFF


def _mix32(h: torch.Tensor) -> torch.Tensor:
    """32-bit finaliser on int64 lanes; every product stays < 2^59 so int64 never overflows."""
    h = ((h ^ (h >> 16)) * 0x45D9F3B) & _M32
    h = ((h ^ (h >> 16)) * 0x45D9F3B) & _M32
    return h ^ (h >> 16)


def _pair_hash(cells: torch.Tensor, genes: torch.Tensor, seed: int) -> torch.Tensor:
    """cells [A,1] (or [1,A]) and genes [1,B] (or [B,1]) int64 → hash of every pair, in [0, 2^32)."""
    return _mix32((cells * 0x9E3779B1 + genes * 0x85EBCA77 + (seed & _M32)) & _M32)


def _value_from_hash(h: torch.Tensor, seed: int) -> torch.Tensor:
    """x ≈ clip(N(3, 0.5), 0.05, 9): Irwin–Hall sum of four 16-bit uniforms from two re-hashes."""
    h1 = _mix32((h + 0x68E31DA4 + (seed & 0xFFFF)) & _M32)
    h2 = _mix32((h1 + 0xB5297A4D) & _M32)
    s = ((h1 & 0xFFFF) + (h1 >> 16) + (h2 & 0xFFFF) + (h2 >> 16)).to(torch.float32)
    z = (s * (1.0 / 65536.0) - 2.0) * 1.7320508
    return torch.clamp(z * 0.5 + 3.0, 0.05, 9.0)


def population(num_cells: int, num_genes: int, avg_degree: float, seed: int):
    """Host-side (numpy) per-cell depth factors and per-gene popularities, fp64."""
    rng = np.random.RandomState(seed)
    rank = rng.permutation(num_genes) + 1                    # popularity rank of gene id g
    shape = rank.astype(np.float64) ** -0.8
    lo, hi = 0.0, float(num_genes)
    for _ in range(200):                                      # Σ_g min(1, c·shape_g) = avg_degree
        mid = 0.5 * (lo + hi)
        if np.minimum(1.0, mid * shape).sum() < avg_degree:
            lo = mid
        else:
            hi = mid
    pop = 0.5 * (lo + hi) * shape
    sigma = 0.35
    depth = np.exp(rng.normal(-0.5 * sigma * sigma, sigma, num_cells))
    return depth, pop


def _threshold(depth: torch.Tensor, pop: torch.Tensor) -> torch.Tensor:
    """int64 threshold on the 32-bit hash for every (cell, gene) pair of the broadcast."""
    return (torch.clamp(depth * pop, max=1.0) * 4294967296.0).to(torch.int64)


def synthetic_bipartite(num_cells: int, num_genes: int, avg_degree: float, seed: int = 10086,
                        device="cpu", chunk_elems: int = 1 << 25,
                        cell_range: Optional[Tuple[int, int]] = None) -> BipartiteGraph:
    """Full-graph structures for a synthetic atlas.  ``cell_range=(lo, hi)`` builds the shard that
    owns cells [lo, hi) (multi-GPU cell sharding): its cell-major CSR has hi-lo rows, its
    gene-major CSR has local cell columns, and the gene normalisers are LOCAL partial sums that
    the caller all-reduces (see parallel.shard_graph)."""
    depth, pop = population(num_cells, num_genes, avg_degree, seed)
    lo, hi = (0, num_cells) if cell_range is None else cell_range
    n_local = hi - lo
    gbits = _lib.COL_U16 if num_genes <= 65536 else _lib.COL_I32
    cbits = _lib.COL_U16 if n_local <= 65536 else _lib.COL_I32
    dt = lambda bits: torch.int16 if bits == _lib.COL_U16 else torch.int32     # noqa: E731

    def to_col(t, bits):            # uint16 values stored in int16 lanes (two's complement wrap)
        return ((t + 32768) % 65536 - 32768).to(torch.int16) if bits == _lib.COL_U16 else t.to(torch.int32)

    # cell-major: rows = local cells; the hash uses GLOBAL cell ids so shards tile the same atlas
    depth_l = depth[lo:hi]
    depth_t = torch.from_numpy(depth_l).to(device); pop_t = torch.from_numpy(pop).to(device)
    genes = torch.arange(num_genes, device=device, dtype=torch.int64)[None, :]
    chunk = max(1, int(chunk_elems // num_genes))
    cnt, cols, vals = [], [], []
    for r0 in range(0, n_local, chunk):
        r1 = min(n_local, r0 + chunk)
        cells = torch.arange(lo + r0, lo + r1, device=device, dtype=torch.int64)[:, None]
        h = _pair_hash(cells, genes, seed)
        mask = h < _threshold(depth_t[r0:r1, None], pop_t[None, :])
        cnt.append(mask.sum(dim=1))
        cols.append(to_col(mask.nonzero(as_tuple=False)[:, 1], gbits))
        vals.append(_value_from_hash(h[mask], seed))
        del h, mask
    deg_c = torch.cat(cnt) if cnt else torch.zeros(0, dtype=torch.int64, device=device)
    rp_c = torch.zeros(n_local + 1, dtype=torch.int64, device=device); rp_c[1:] = torch.cumsum(deg_c, 0)
    cell_csr = Csr(rp_c, torch.cat(cols), torch.cat(vals), num_genes, n_local, gbits, _balanced_row_perm(deg_c))
    del cols, vals

    # gene-major transpose, generated independently from the same hash
    cells_all = torch.arange(lo, hi, device=device, dtype=torch.int64)[None, :]
    chunk = max(1, int(chunk_elems // max(1, n_local)))
    cnt, cols, vals = [], [], []
    for g0 in range(0, num_genes, chunk):
        g1 = min(num_genes, g0 + chunk)
        gs = torch.arange(g0, g1, device=device, dtype=torch.int64)[:, None]
        h = _pair_hash(cells_all, gs, seed)
        mask = h < _threshold(depth_t[None, :], pop_t[g0:g1, None])
        cnt.append(mask.sum(dim=1))
        cols.append(to_col(mask.nonzero(as_tuple=False)[:, 1], cbits))
        vals.append(_value_from_hash(h[mask], seed))
        del h, mask
    deg_g = torch.cat(cnt)
    rp_g = torch.zeros(num_genes + 1, dtype=torch.int64, device=device); rp_g[1:] = torch.cumsum(deg_g, 0)
    gene_csr = Csr(rp_g, torch.cat(cols), torch.cat(vals), n_local, num_genes, cbits, _balanced_row_perm(deg_g))
    del cols, vals

    def seg_sum(csr, n):
        out = torch.zeros(n, dtype=torch.float32, device=device)
        seg = torch.repeat_interleave(torch.arange(n, device=device), csr.rowptr[1:] - csr.rowptr[:-1], output_size=csr.nnz)
        return out.index_add_(0, seg, csr.x)

    rs, cs = seg_sum(cell_csr, n_local), seg_sum(gene_csr, num_genes)
    norm_c = torch.where(deg_c > 0, deg_c.float() / rs.clamp(min=1e-30), torch.zeros_like(rs))
    norm_g = torch.where(deg_g > 0, deg_g.float() / cs.clamp(min=1e-30), torch.zeros_like(cs))
    bg = BipartiteGraph(num_genes, n_local, n_local, cell_csr, gene_csr,
                        norm_c, 1.0 / (deg_c + 1).float(), norm_g, 1.0 / (deg_g + 1).float())
    bg.local_deg_g, bg.local_colsum_g, bg.rowsum_c = deg_g, cs, rs
    return bg


def synthetic_features(graph: BipartiteGraph, dense_dim: int, seed: int = 10086) -> torch.Tensor:
    """features = cat[gene_feat; cell_feat]: gene_feat ~ N(0, 0.66²) stands in for the PCA embedding
    (utils/preprocess_internal.py:186-187); cell_feat = (X / (rowsum+1e-6)) · gene_feat as in
    preprocess_internal.py:194-196 — which is the cell-destination aggregation itself, so on a
    GPU it runs through wsage_spmm."""
    dev = graph.device
    gen = torch.Generator().manual_seed(seed)
    gene_feat = (torch.randn(graph.num_genes, dense_dim, generator=gen) * 0.66).to(dev)
    cs = graph.cell_csr
    rowsum = torch.zeros(cs.n_dst, dtype=torch.float32, device=dev)
    seg = torch.repeat_interleave(torch.arange(cs.n_dst, device=dev), cs.rowptr[1:] - cs.rowptr[:-1], output_size=cs.nnz)
    rowsum.index_add_(0, seg, cs.x)
    inv = 1.0 / (rowsum + 1e-6)
    if dev.type == "cuda":
        from .ops import spmm
        cell_feat, _, _ = spmm(cs, gene_feat, dscale=inv)
    else:
        col = cs.col.to(torch.int64)
        if cs.col_bits == _lib.COL_U16:
            col = col & 0xFFFF
        cell_feat = torch.zeros(cs.n_dst, dense_dim).index_add_(0, seg, gene_feat[col] * cs.x[:, None]) * inv[:, None]
    return torch.cat([gene_feat, cell_feat], dim=0)
