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
        if getattr(bg, "densified", False) or bg.cell_csr.dense is not None or bg.gene_csr.dense is not None:
            raise ValueError("DeepSortGraph.from_bipartite needs the complete CSRs: build it before BipartiteGraph.densify() "
                             "(a densified graph keeps the popular genes' edges outside the CSRs)")
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


def _strip_dense(csr: Csr, slot: torch.Tensor, side: str, fill=None, chunk_edges: int = 1 << 27) -> Csr:
    """Removes the entries of the popular genes (``slot[gene] >= 0``) from ``csr``.

    side == 'src': the genes are the COLUMNS of csr (cell-destination CSR); side == 'dst': the genes are the ROWS
    (gene-destination CSR).  ``fill(row, slot, x)`` receives the removed entries (side 'src' only: row = cell).
    Works in row chunks so the index temporaries stay bounded at atlas scale."""
    dev = csr.x.device
    n_dst, n_src = csr.n_dst, csr.n_src
    rowptr_h = csr.rowptr.cpu()
    deg = csr.rowptr[1:] - csr.rowptr[:-1]
    new_deg = torch.zeros(n_dst, dtype=torch.int64, device=dev)
    cols, vals = [], []
    r0 = 0
    while r0 < n_dst:
        # largest r1 with rowptr[r1] - rowptr[r0] <= chunk_edges (at least one row)
        r1 = int(torch.searchsorted(rowptr_h, rowptr_h[r0] + chunk_edges, right=True)) - 1
        r1 = min(max(r1, r0 + 1), n_dst)
        e0, e1 = int(rowptr_h[r0]), int(rowptr_h[r1])
        if e1 > e0:
            row = torch.repeat_interleave(torch.arange(r0, r1, device=dev), deg[r0:r1], output_size=e1 - e0)
            x = csr.x[e0:e1]
            if side == 'src':
                col = csr.col[e0:e1].to(torch.int64)
                if csr.col_bits == _lib.COL_U16:
                    col = col & 0xFFFF
                s_e = slot[col]
                del col
            else:
                s_e = slot[row]
            hit = s_e >= 0
            if fill is not None:
                fill(row[hit], s_e[hit], x[hit])
            keep = ~hit
            cols.append(csr.col[e0:e1][keep])
            vals.append(x[keep])
            new_deg[r0:r1] = torch.zeros(r1 - r0, dtype=torch.int64, device=dev).index_add_(0, row - r0, keep.to(torch.int64))
            del row, s_e, hit, keep
        r0 = r1
    rp = torch.zeros(n_dst + 1, dtype=torch.int64, device=dev)
    rp[1:] = torch.cumsum(new_deg, 0)
    col_new = torch.cat(cols) if cols else csr.col[:0]
    x_new = torch.cat(vals) if vals else csr.x[:0]
    return Csr(rp, col_new, x_new, n_src, n_dst, csr.col_bits, _balanced_row_perm(new_deg))


def _build_dense_block(cell_csr: Csr, slot: torch.Tensor, gene_ids: torch.Tensor, fmt: int):
    """(stripped cell-destination CSR, DenseBlock): the popular genes' entries of every cell as 16-bit planes
    ``plane[cell // 128][slot // 32][cell % 128][slot % 32]`` (include/wsage.h).  fp16 hi + lo of x * 2^k with
    max|x| * 2^k in [2^13, 2^14) (fp32-grade), or bf16(x) (BASELINE configs[2])."""
    dev = cell_csr.x.device
    lib = _lib.load()
    n_cells, gd = cell_csr.n_dst, int(gene_ids.numel())
    slots_pad = int(lib.wsage_dense16_slots_pad(gd))
    nb = slots_pad // 32
    n_tiles = (n_cells + 127) // 128
    numel = n_tiles * nb * 128 * 32
    f16 = fmt == _lib.D16_F16X2
    x_scale = 1.0
    if f16 and cell_csr.nnz:
        xmax = float(cell_csr.x.abs().max())
        if xmax > 0 and np.isfinite(xmax):
            x_scale = float(2.0 ** (14 - np.frexp(xmax)[1]))
    hi = torch.zeros(numel, dtype=torch.float16 if f16 else torch.bfloat16, device=dev)
    lo = torch.zeros(numel, dtype=torch.float16, device=dev) if f16 else None
    moved = [0]

    def fill(row, s, x):
        idx = (((row // 128) * nb + s // 32) * 128 + row % 128) * 32 + s % 32
        if f16:
            v = x * x_scale
            h = v.to(torch.float16)
            hi[idx] = h
            lo[idx] = (v - h.to(torch.float32)).to(torch.float16)
        else:
            hi[idx] = x.to(torch.bfloat16)
        moved[0] += int(x.numel())

    stripped = _strip_dense(cell_csr, slot, 'src', fill)
    block = DenseBlock(hi.view(torch.int16), lo.view(torch.int16) if lo is not None else None, fmt, n_cells, gd, slots_pad,
                       gene_ids.to(torch.int32).contiguous(), slot.to(torch.int32).contiguous(), x_scale, moved[0])
    return stripped, block


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

    def densify(self, threshold: float = 0.05, max_bytes: int = 96 << 30, fmt: str = "f16x2") -> "BipartiteGraph":
        """Splits X = X_sparse + X_dense IN PLACE for the full-graph path: genes expressed in at least
        ``threshold`` of the (local) support cells leave every CSR and become ONE zero-filled block of 16-bit
        tiles (``Csr.dense``, shared by all directions) that wsage_dense16 runs on the tensor cores before the
        CSR walk — for such a gene, multiplying through its zeros on tcgen05 is cheaper than walking its edges.
        ``fmt``: "f16x2" (fp16 hi + lo, three products: results agree with the plain CSR path to fp32 rounding)
        or "bf16" (one product; BASELINE configs[2]).  The popular set is capped so that the planes stay within
        ``max_bytes``.  A gene set that covers every entry leaves empty CSRs (wsage_spmm then only reduces).
        Not for the mini-batch surface: ``DeepSortGraph.from_bipartite`` needs the complete CSRs and refuses a
        densified graph."""
        if getattr(self, "densified", False) or self.gene_csr.n_src == 0 or self.nnz == 0:
            return self
        fmt_id = {"f16x2": _lib.D16_F16X2, "bf16": _lib.D16_BF16}[fmt]
        dev = self.device
        gcsr = self.gene_csr
        deg_g = (gcsr.rowptr[1:] - gcsr.rowptr[:-1])
        rho = deg_g.to(torch.float64) / float(gcsr.n_src)
        cand = torch.nonzero((rho >= threshold) & (deg_g > 0)).flatten()
        planes = 2 if fmt_id == _lib.D16_F16X2 else 1
        cells_pad = (self.num_cells + 127) // 128 * 128
        cap = int(max_bytes // (2 * planes * max(1, cells_pad))) // 128 * 128
        if cand.numel() > cap:                       # keep the most popular ones
            cand = cand[torch.argsort(deg_g[cand], descending=True)[:cap]]
            cand = torch.sort(cand).values
        if cand.numel() == 0:
            return self
        slot = torch.full((self.num_genes,), -1, dtype=torch.int64, device=dev)
        slot[cand] = torch.arange(cand.numel(), device=dev)
        new_cell, block = _build_dense_block(self.cell_csr, slot, cand, fmt_id)
        new_cell.dense, new_cell.dense_side = block, 0
        new_gene = _strip_dense(self.gene_csr, slot, 'dst')
        new_gene.dense, new_gene.dense_side = block, 1
        new_t = self.cell_csr_t
        if new_t is not None:
            new_t = _strip_dense(new_t, slot, 'dst')
            new_t.dense, new_t.dense_side = block, 1
        self.cell_csr, self.gene_csr, self.cell_csr_t = new_cell, new_gene, new_t
        self.densified = True
        self.dense_genes = cand
        return self

    def support_cell_csr(self) -> Csr:
        """Rows of ``cell_csr`` that belong to support cells (all of them unless the graph carries test cells): the
        transpose of ``gene_csr``, used by the backward of the gene aggregation.  Keeps the dense block."""
        cs, ns = self.cell_csr, self.num_support
        if ns == self.num_cells:
            return cs
        cached = getattr(self, "_support_csr", None)
        if cached is None or cached[0] is not cs:
            e = int(cs.rowptr[ns])
            deg = cs.rowptr[1:ns + 1] - cs.rowptr[:ns]
            sub = Csr(cs.rowptr[:ns + 1].contiguous(), cs.col[:e], cs.x[:e], cs.n_src, ns, cs.col_bits,
                      _balanced_row_perm(deg), cs.dense, cs.dense_side)
            cached = self._support_csr = (cs, sub)
        return cached[1]

    def transpose_of_cell_csr(self) -> Csr:
        if self.cell_csr_t is not None:
            return self.cell_csr_t
        assert self.num_support == self.num_cells
        return self.gene_csr
