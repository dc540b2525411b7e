"""Graph containers for the hot path.

``DeepSortGraph`` is the parent graph exactly as the reference's builders leave it
(utils/preprocess_internal.py:107-110,168-173,211-215; utils/preprocess.py:129-134,185-187,
216-221): genes ``0..G-1`` then cells, ``ndata['id']`` (gene → index, cell → -1),
``ndata['features']``, per-destination-normalised ``edata['weight']`` and a unit self-loop on
every node — stored as a destination-major CSR instead of a DGLGraph.  The mini-batch surface
(``nodeflow.NeighborSampler``) slices it.

``BipartiteGraph`` is the same graph factored for the full-graph throughput path: one CSR of
RAW expression values per direction plus per-node scale vectors, because both directions'
normalised weights are per-destination rescalings of the same ``x_cg``
(``w_{g→c} = x·deg_c/Σ_g x``, ``w_{c→g} = x·deg_g/Σ_c x``).

Builders here are vectorised restatements of the reference's per-node Python loop
(``normalize_weight``, preprocess_internal.py:15-23); file parsing / PCA stay with the caller
(out of scope).
"""
from dataclasses import dataclass, field
from typing import Dict, Optional

import numpy as np
import scipy.sparse as sp
import torch

from . import _lib
from .ops import Csr, DenseBlock


def _segment_sum_f32(data: np.ndarray, ptr: np.ndarray) -> np.ndarray:
    out = np.zeros(len(ptr) - 1, dtype=np.float32)
    nz = ptr[1:] > ptr[:-1]
    if data.size:
        out[nz] = np.add.reduceat(data.astype(np.float32), ptr[:-1][nz]).astype(np.float32)
    return out


def _filter_csr(x, threshold):
    x = sp.csr_matrix(x).astype(np.float32)
    x.sort_indices()
    coo = x.tocoo()
    keep = coo.data > threshold                       # preprocess_internal.py:158
    return sp.csr_matrix((coo.data[keep], (coo.row[keep], coo.col[keep])), shape=x.shape)


@dataclass
class DeepSortGraph:
    num_genes: int
    num_cells: int
    in_rowptr: torch.Tensor      # int64 [N+1]   in-edges of node v: [in_rowptr[v], in_rowptr[v+1])
    in_src: torch.Tensor         # int64 [E]     parent id of the source node
    in_weight: torch.Tensor      # fp32  [E]     edata['weight']
    ndata: Dict[str, torch.Tensor] = field(default_factory=dict)   # 'id' int32 [N], 'features' fp32 [N, D]

    def number_of_nodes(self):
        return self.num_genes + self.num_cells

    def number_of_edges(self):
        return int(self.in_src.shape[0])

    @property
    def device(self):
        return self.in_rowptr.device

    def to(self, device):
        return DeepSortGraph(self.num_genes, self.num_cells, self.in_rowptr.to(device), self.in_src.to(device),
                             self.in_weight.to(device), {k: v.to(device) for k, v in self.ndata.items()})

    @classmethod
    def from_bipartite(cls, bg: "BipartiteGraph", features=None):
        """Device-side build of the reference-contract graph (normalised weights, unit self-loop last in every
        row) from the factored full-graph structures — the vectorised form of normalize_weight + add_edges
        (utils/preprocess_internal.py:15-23,170-173,211-214) for atlases whose edge list never exists on the
        host.  Every cell must be a support cell.  ``in_src`` is int32 (N < 2^31) to halve its footprint."""
        assert bg.num_support == bg.num_cells, "test cells (gene->cell only) need the host builder"
        dev, g, c = bg.device, bg.num_genes, bg.num_cells
        n = g + c

        def cols(csr):
            col = csr.col.to(torch.int32)
            return col & 0xFFFF if csr.col_bits == _lib.COL_U16 else col

        deg = torch.cat([bg.gene_csr.rowptr[1:] - bg.gene_csr.rowptr[:-1], bg.cell_csr.rowptr[1:] - bg.cell_csr.rowptr[:-1]])
        rowptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
        rowptr[1:] = torch.cumsum(deg + 1, 0)
        e_total = int(rowptr[-1])
        src = torch.empty(e_total, dtype=torch.int32, device=dev)
        w = torch.empty(e_total, dtype=torch.float32, device=dev)
        for csr, norm, row0, col_off in ((bg.gene_csr, bg.norm_g, 0, g), (bg.cell_csr, bg.norm_c, g, 0)):
            d = csr.rowptr[1:] - csr.rowptr[:-1]
            row_of = torch.repeat_interleave(torch.arange(csr.n_dst, device=dev), d, output_size=csr.nnz)
            pos = torch.arange(csr.nnz, device=dev) + row_of + int(rowptr[row0])      # one self-loop slot per earlier row
            src[pos] = cols(csr) + col_off
            w[pos] = csr.x * norm[row_of]
            del row_of, pos
        loops = rowptr[1:] - 1
        src[loops] = torch.arange(n, dtype=torch.int32, device=dev)
        w[loops] = 1.0
        node_id = torch.cat([torch.arange(g, dtype=torch.int32, device=dev), torch.full((c,), -1, dtype=torch.int32, device=dev)])
        nd = {"id": node_id}
        if features is not None:
            nd["features"] = features
        return cls(g, c, rowptr, src, w, nd)

    @classmethod
    def from_edges(cls, src, dst, weight, node_id, features, num_genes):
        """From a COO edge list with final weights (e.g. arrays exported from a reference DGLGraph)."""
        src = torch.as_tensor(src, dtype=torch.int64)
        dst = torch.as_tensor(dst, dtype=torch.int64)
        n = int(node_id.shape[0])
        order = torch.sort(dst, stable=True).indices
        rowptr = torch.zeros(n + 1, dtype=torch.int64)
        rowptr[1:] = torch.cumsum(torch.bincount(dst, minlength=n), 0)
        nd = {"id": torch.as_tensor(node_id, dtype=torch.int32).reshape(-1)}
        if features is not None:
            nd["features"] = torch.as_tensor(features, dtype=torch.float32)
        return cls(num_genes, n - num_genes, rowptr, src[order], torch.as_tensor(weight, dtype=torch.float32)[order], nd)

    @classmethod
    def from_expression(cls, x_support, x_test=None, threshold=0.0, features=None):
        """x_support [C, G]: both edge directions; x_test [Ct, G]: gene→cell only (inference)."""
        xs = _filter_csr(x_support, threshold)
        num_genes = xs.shape[1]
        xa = xs if x_test is None else sp.vstack([xs, _filter_csr(x_test, threshold)]).tocsr()
        n_cells = xa.shape[0]
        n = num_genes + n_cells
        # gene → cell edges, cell-major:  w = deg_c * x / rowsum_c   (fp32, same operation order)
        deg_c = np.diff(xa.indptr).astype(np.int64)
        rowsum_c = _segment_sum_f32(xa.data, xa.indptr)
        row_of = np.repeat(np.arange(n_cells), deg_c)
        w_gc = (deg_c[row_of].astype(np.float32) * xa.data) / rowsum_c[row_of]
        # cell → gene edges, gene-major (support cells only)
        xg = xs.tocsc(); xg.sort_indices()
        deg_g = np.diff(xg.indptr).astype(np.int64)
        colsum_g = _segment_sum_f32(xg.data, xg.indptr)
        col_of = np.repeat(np.arange(num_genes), deg_g)
        w_cg = (deg_g[col_of].astype(np.float32) * xg.data) / colsum_g[col_of]
        # in-CSR over all nodes; the unit self-loop is the last in-edge of every node
        indeg = np.concatenate([deg_g, deg_c]) + 1
        rowptr = np.zeros(n + 1, dtype=np.int64)
        rowptr[1:] = np.cumsum(indeg)
        e_total = int(rowptr[-1])
        src = np.empty(e_total, dtype=np.int64)
        w = np.empty(e_total, dtype=np.float32)
        pos_g = np.arange(xg.nnz) + col_of                      # each earlier gene added one self-loop slot
        src[pos_g] = xg.indices.astype(np.int64) + num_genes
        w[pos_g] = w_cg
        pos_c = np.arange(xa.nnz) + row_of + int(rowptr[num_genes])
        src[pos_c] = xa.indices
        w[pos_c] = w_gc
        loops = rowptr[1:] - 1
        src[loops] = np.arange(n)
        w[loops] = 1.0
        node_id = np.concatenate([np.arange(num_genes, dtype=np.int32), np.full(n_cells, -1, dtype=np.int32)])
        nd = {"id": torch.from_numpy(node_id)}
        if features is not None:
            nd["features"] = torch.as_tensor(features, dtype=torch.float32)
        return cls(num_genes, n_cells, torch.from_numpy(rowptr), torch.from_numpy(src), torch.from_numpy(w), nd)


def _balanced_row_perm(deg: torch.Tensor) -> torch.Tensor:
    """Heavy rows first: warps pick rows in this order, so the long rows start early and the tail
    of the grid is made of short rows."""
    return torch.sort(deg, descending=True, stable=True).indices.to(torch.int32)


def _to_csr(rowptr, col, x, n_src, n_dst, device):
    rowptr = torch.as_tensor(rowptr, dtype=torch.int64)
    deg = rowptr[1:] - rowptr[:-1]
    if n_src <= 65536:
        colt = torch.from_numpy(np.asarray(col).astype(np.uint16).view(np.int16))
        bits = _lib.COL_U16
    else:
        colt = torch.as_tensor(np.asarray(col), dtype=torch.int32)
        bits = _lib.COL_I32
    return Csr(rowptr.to(device), colt.to(device), torch.as_tensor(x, dtype=torch.float32).to(device),
               n_src, n_dst, bits, _balanced_row_perm(deg).to(device))


def _split_dense(csr: Csr, slot: torch.Tensor, side: str, tile: int, chunk_edges: int = 1 << 27) -> Csr:
    """Moves the entries of the popular genes (``slot[gene] >= 0``) out of ``csr`` into a ``DenseBlock``.

    side == 'src': the genes are the COLUMNS of csr (cell-destination CSR); the block's sources are the
    popular genes, its destinations every row.  side == 'dst': the genes are the ROWS (gene-destination
    CSR); the block's sources are every column (cell), its destinations the popular genes' slots.
    Works in row chunks so the index temporaries stay bounded at atlas scale."""
    dev = csr.x.device
    md = int((slot >= 0).sum())
    n_dst, n_src = csr.n_dst, csr.n_src
    if side == 'src':
        k, t = md, n_dst
        ids = torch.nonzero(slot >= 0).flatten()
        order = torch.argsort(slot[ids])
        src_ids, dst_map = ids[order].to(torch.int32), None
    else:
        k, t = n_src, md
        src_ids, dst_map = None, slot.to(torch.int32)
    n_tiles = (t + tile - 1) // tile
    xd = torch.zeros(n_tiles * k * tile, dtype=torch.float32, device=dev)
    rowptr_h = csr.rowptr.cpu()
    deg = csr.rowptr[1:] - csr.rowptr[:-1]
    new_deg = torch.zeros(n_dst, dtype=torch.int64, device=dev)
    cols, vals = [], []
    moved = 0
    r0 = 0
    while r0 < n_dst:
        # largest r1 with rowptr[r1] - rowptr[r0] <= chunk_edges (at least one row)
        r1 = int(torch.searchsorted(rowptr_h, rowptr_h[r0] + chunk_edges, right=True)) - 1
        r1 = min(max(r1, r0 + 1), n_dst)
        e0, e1 = int(rowptr_h[r0]), int(rowptr_h[r1])
        if e1 > e0:
            col = csr.col[e0:e1].to(torch.int64)
            if csr.col_bits == _lib.COL_U16:
                col = col & 0xFFFF
            row = torch.repeat_interleave(torch.arange(r0, r1, device=dev), deg[r0:r1], output_size=e1 - e0)
            x = csr.x[e0:e1]
            if side == 'src':
                s_e = slot[col]
                hit = s_e >= 0
                idx = ((row[hit] // tile) * k + s_e[hit]) * tile + row[hit] % tile
            else:
                s_e = slot[row]
                hit = s_e >= 0
                idx = ((s_e[hit] // tile) * k + col[hit]) * tile + s_e[hit] % tile
            xd[idx] = x[hit]
            moved += int(hit.sum())
            keep = ~hit
            cols.append(csr.col[e0:e1][keep])
            vals.append(x[keep])
            new_deg[r0:r1] = torch.zeros(r1 - r0, dtype=torch.int64, device=dev).index_add_(0, row - r0, keep.to(torch.int64))
            del col, row, s_e, hit, idx, keep
        r0 = r1
    rp = torch.zeros(n_dst + 1, dtype=torch.int64, device=dev)
    rp[1:] = torch.cumsum(new_deg, 0)
    col_new = torch.cat(cols) if cols else csr.col[:0]
    x_new = torch.cat(vals) if vals else csr.x[:0]
    return Csr(rp, col_new, x_new, n_src, n_dst, csr.col_bits, _balanced_row_perm(new_deg),
               DenseBlock(xd, k, t, src_ids, dst_map, moved))


@dataclass
class BipartiteGraph:
    """Full-graph factorisation.  ``cell_csr`` rows = all cells (support then test), columns =
    genes; ``gene_csr`` rows = genes, columns = SUPPORT cells (test cells send nothing to genes,
    utils/preprocess.py:185-187).  When there are no test cells each CSR is the other's
    transpose and the backward passes reuse them; otherwise ``cell_csr_t`` holds the transpose
    of ``cell_csr`` (only needed for training, which the reference never does with test cells)."""
    num_genes: int
    num_cells: int          # all cells
    num_support: int
    cell_csr: Csr
    gene_csr: Csr
    norm_c: torch.Tensor    # fp32 [C]  deg_c / Σ_g x_cg     (0 for empty rows)
    mean_c: torch.Tensor    # fp32 [C]  1 / (deg_c + 1)
    norm_g: torch.Tensor    # fp32 [G]  deg_g / Σ_c x_cg     (0 for zero-degree genes)
    mean_g: torch.Tensor    # fp32 [G]  1 / (deg_g + 1)
    cell_csr_t: Optional[Csr] = None

    @property
    def device(self):
        return self.norm_c.device

    @property
    def nnz(self):
        return self.cell_csr.nnz + (self.cell_csr.dense.nnz if self.cell_csr.dense is not None else 0)

    @classmethod
    def from_expression(cls, x_support, x_test=None, threshold=0.0, device="cpu"):
        xs = _filter_csr(x_support, threshold)
        g = xs.shape[1]
        xa = xs if x_test is None else sp.vstack([xs, _filter_csr(x_test, threshold)]).tocsr()
        xa.sort_indices()
        xg = xs.tocsc(); xg.sort_indices()
        deg_c = np.diff(xa.indptr); deg_g = np.diff(xg.indptr)
        rs = _segment_sum_f32(xa.data, xa.indptr); cs = _segment_sum_f32(xg.data, xg.indptr)
        with np.errstate(divide="ignore", invalid="ignore"):
            norm_c = np.where(deg_c > 0, deg_c.astype(np.float32) / rs, 0).astype(np.float32)
            norm_g = np.where(deg_g > 0, deg_g.astype(np.float32) / cs, 0).astype(np.float32)
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)   # noqa: E731
        out = cls(g, xa.shape[0], xs.shape[0],
                  _to_csr(xa.indptr, xa.indices, xa.data, g, xa.shape[0], device),
                  _to_csr(xg.indptr, xg.indices, xg.data, xs.shape[0], g, device),
                  t(norm_c), t((1.0 / (deg_c + 1)).astype(np.float32)),
                  t(norm_g), t((1.0 / (deg_g + 1)).astype(np.float32)))
        if x_test is not None:
            xt = xa.tocsc(); xt.sort_indices()
            out.cell_csr_t = _to_csr(xt.indptr, xt.indices, xt.data, xa.shape[0], g, device)
        return out

    def densify(self, threshold: float = 0.3, max_bytes: int = 32 << 30, directions=("gene",)) -> "BipartiteGraph":
        """Splits X = X_sparse + X_dense IN PLACE for the full-graph path: genes expressed in at least
        ``threshold`` of the (local) support cells leave the CSRs and become zero-filled dense blocks
        (``Csr.dense``), which wsage_spmm runs on the FMA-bound dense-block kernel before the CSR walk.
        Results are unchanged up to fp32 summation order.  The popular set is capped so that the blocks
        stay within ``max_bytes``.  ``directions``: "gene" splits the gene-destination CSR(s) (whole dense
        destination tiles leave the CSR walk, whose remaining tiles are unaffected: the profitable case),
        "cell" also splits the cell-destination CSR (every row gets thinner, which costs the CSR walk
        efficiency: measured no gain at 10 % density, kept for denser atlases).  Not for the mini-batch
        surface (``DeepSortGraph.from_bipartite`` needs the complete CSRs)."""
        if getattr(self, "densified", False) or self.gene_csr.n_src == 0 or self.nnz == 0:
            return self
        dev = self.device
        gcsr = self.gene_csr
        deg_g = (gcsr.rowptr[1:] - gcsr.rowptr[:-1])
        rho = deg_g.to(torch.float64) / float(gcsr.n_src)
        cand = torch.nonzero(rho >= threshold).flatten()
        do_gene, do_cell = "gene" in directions, "cell" in directions
        n_blocks = (int(do_gene) * (2 if self.cell_csr_t is not None else 1)) + int(do_cell)
        if n_blocks == 0:
            return self
        cap = int(max_bytes // (4 * n_blocks * max(1, self.num_cells)))
        if cand.numel() > cap:                       # keep the most popular ones
            cand = cand[torch.argsort(deg_g[cand], descending=True)[:cap]]
            cand = torch.sort(cand).values
        if cand.numel() == 0:
            return self
        slot = torch.full((self.num_genes,), -1, dtype=torch.int64, device=dev)
        slot[cand] = torch.arange(cand.numel(), device=dev)
        tile = int(_lib.load().wsage_dense_tile())
        new_cell = _split_dense(self.cell_csr, slot, 'src', tile) if do_cell else self.cell_csr
        new_gene = _split_dense(self.gene_csr, slot, 'dst', tile) if do_gene else self.gene_csr
        new_t = self.cell_csr_t
        if do_gene and new_t is not None:
            new_t = _split_dense(new_t, slot, 'dst', tile)
        if new_cell.nnz == 0 or new_gene.nnz == 0 or (new_t is not None and new_t.nnz == 0):
            return self                              # the CSR walk needs a non-empty remainder
        self.cell_csr, self.gene_csr, self.cell_csr_t = new_cell, new_gene, new_t
        self.densified = True
        self.dense_genes = cand
        return self

    def transpose_of_cell_csr(self) -> Csr:
        if self.cell_csr_t is not None:
            return self.cell_csr_t
        assert self.num_support == self.num_cells
        return self.gene_csr
