"""CPU-only checks: the C-ABI library loads and exports what include/wsage.h declares, and the
host-side graph / mini-batch logic matches the oracle's restatement of the reference contract."""
import re
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import graph_oracle
from scds_helpers import golden_csr, golden_graph

import scdeepsort_b200 as sd

ROOT = Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    header = (ROOT / "include" / "wsage.h").read_text()
    declared = set(re.findall(r"\b(wsage_[a-z_0-9]+)\s*\(", header))
    assert declared == set(sd._lib.EXPORTS)
    lib = sd._lib.load()
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.wsage_version() >= 1000
    assert lib.wsage_last_error() is not None


def test_spmm_args_struct_layout_matches_header():
    header = (ROOT / "include" / "wsage.h").read_text()
    body = header[header.index("typedef struct wsage_spmm_args {"):header.index("} wsage_spmm_args;")]
    names = re.findall(r"\b([a-z_]+);", body)
    assert names == [f[0] for f in sd._lib.SpmmArgs._fields_]
    body = header[header.index("typedef struct wsage_dense16_args {"):header.index("} wsage_dense16_args;")]
    names = re.findall(r"\b([a-z_]+);", body)
    assert names == [f[0] for f in sd._lib.Dense16Args._fields_]
    body = header[header.index("typedef struct wsage_peer_reduce_args {"):header.index("} wsage_peer_reduce_args;")]
    names = re.findall(r"\b([a-z_]+);", body)
    assert names == [f[0] for f in sd._lib.PeerReduceArgs._fields_]


def test_argument_errors_are_reported_not_thrown():
    lib = sd._lib.load()
    a = sd._lib.SpmmArgs()
    a.n_dst, a.n_src, a.dim, a.col_bits = 4, 4, 0, 32
    import ctypes
    assert lib.wsage_spmm(ctypes.byref(a), None) == sd._lib.EINVAL
    assert b"dim" in lib.wsage_last_error()
    a.dim, a.col_bits, a.n_src = 8, 16, 70000
    assert lib.wsage_spmm(ctypes.byref(a), None) == sd._lib.EINVAL
    assert b"uint16" in lib.wsage_last_error()
    with pytest.raises(RuntimeError):
        sd._lib.check(sd._lib.EINVAL, "x")


def test_model_refuses_cpu():
    m = sd.GNN(8, 4, 3, 1, 10, activation=torch.relu)
    with pytest.raises(RuntimeError, match="CUDA only"):
        m(object())


def test_state_dict_keys_match_reference(golden_train):
    z = golden_train
    m = sd.GNN(int(z["dense_dim"]), int(z["hidden"]), int(z["num_labels"]), 2, int(z["num_genes"]), activation=torch.relu)
    ref = {k[len("L2/"):]: z[k].shape for k in z.files
           if k.startswith("L2/") and (k.startswith("L2/layers.") or k in ("L2/alpha", "L2/linear.weight", "L2/linear.bias"))}
    mine = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert mine == ref
    assert torch.all(m.alpha == 1)


def _edge_set(rowptr, src, w, n):
    deg = (rowptr[1:] - rowptr[:-1])
    dst = torch.repeat_interleave(torch.arange(n), deg)
    key = dst * n + src
    o = torch.argsort(key)
    return key[o], w[o]


def test_graph_builder_matches_oracle(golden_test):
    z = golden_test
    og = graph_oracle.build_graph(golden_csr(z), golden_csr(z, "xt"))
    g = sd.DeepSortGraph.from_expression(golden_csr(z), golden_csr(z, "xt"))
    n = og.num_nodes
    assert g.number_of_nodes() == n and g.number_of_edges() == og.src.shape[0]
    k1, w1 = _edge_set(g.in_rowptr, g.in_src, g.in_weight, n)
    o = torch.argsort(og.dst * n + og.src)
    assert torch.equal(k1, (og.dst * n + og.src)[o])
    assert float(((w1 - og.weight[o]).abs() / og.weight[o]).max()) < 4e-7
    assert torch.equal(g.ndata["id"], og.node_id)
    # from_edges on the reference's own arrays gives the same CSR
    gg = golden_graph(z, num_genes=int(z["num_genes"]))
    g2 = sd.DeepSortGraph.from_edges(gg.src, gg.dst, gg.weight, gg.node_id, gg.features, gg.num_genes)
    k2, w2 = _edge_set(g2.in_rowptr, g2.in_src, g2.in_weight, n)
    assert torch.equal(k1, k2)


def test_bipartite_factorisation_reproduces_reference_weights(golden_train):
    z = golden_train
    x = golden_csr(z)
    bg = sd.BipartiteGraph.from_expression(x)
    gg = golden_graph(z)
    G = bg.num_genes
    # gene→cell edges: w = x * norm_c ;  cell→gene edges: w = x * norm_g
    cs = bg.cell_csr
    col = torch.from_numpy(cs.col.numpy().view(np.uint16).astype(np.int64))
    deg = cs.rowptr[1:] - cs.rowptr[:-1]
    row = torch.repeat_interleave(torch.arange(bg.num_cells), deg)
    w = cs.x * bg.norm_c[row]
    n = gg.num_nodes
    key = (row + G) * n + col
    mask = (gg.dst >= G) & (gg.src < G)
    rkey = (gg.dst * n + gg.src)[mask]
    o1, o2 = torch.argsort(key), torch.argsort(rkey)
    assert torch.equal(key[o1], rkey[o2])
    assert float(((w[o1] - gg.weight[mask][o2]).abs() / gg.weight[mask][o2]).max()) < 5e-7
    assert torch.allclose(bg.mean_c, 1.0 / (deg + 1).float())
    gs = bg.gene_csr
    degg = gs.rowptr[1:] - gs.rowptr[:-1]
    rowg = torch.repeat_interleave(torch.arange(G), degg)
    colg = gs.col.to(torch.int64) if gs.col_bits == 32 else torch.from_numpy(gs.col.numpy().view(np.uint16).astype(np.int64))
    wg = gs.x * bg.norm_g[rowg]
    keyg = rowg * n + (colg + G)
    maskg = (gg.dst < G) & (gg.src >= G)
    rkeyg = (gg.dst * n + gg.src)[maskg]
    o1, o2 = torch.argsort(keyg), torch.argsort(rkeyg)
    assert torch.equal(keyg[o1], rkeyg[o2])
    assert float(((wg[o1] - gg.weight[maskg][o2]).abs() / gg.weight[maskg][o2]).max()) < 5e-7
    assert int((degg == 0).sum()) > 0 and torch.all(bg.norm_g[degg == 0] == 0)
    # columns ascending inside each row; permutation is a permutation
    assert torch.equal(torch.sort(cs.row_perm.long()).values, torch.arange(bg.num_cells))
    d = col[1:] - col[:-1]
    inner = torch.ones_like(d, dtype=torch.bool)
    inner[(cs.rowptr[1:-1] - 1).clamp(min=0)] = False
    assert torch.all(d[inner] > 0)


@pytest.mark.parametrize("n_layers", [1, 2])
def test_sampler_full_neighbour_matches_oracle_flow(golden_train, n_layers):
    z = golden_train
    gg = golden_graph(z)
    g = sd.DeepSortGraph.from_edges(gg.src, gg.dst, gg.weight, gg.node_id, gg.features, gg.num_genes)
    seeds = torch.from_numpy(z["train_ids"])[:37]
    nf = next(iter(sd.NeighborSampler(g, 37, g.number_of_nodes(), n_layers, 'in', shuffle=False, num_workers=8,
                                      seed_nodes=seeds)))
    nf.copy_from_parent()
    of = graph_oracle.full_neighbor_flow(gg, seeds, n_layers)
    assert nf.num_layers == n_layers + 1
    for i in range(n_layers + 1):
        assert torch.equal(nf.layer_parent_nid(i), of.layer_nid[i])
        assert torch.equal(nf.layers[i].data["id"], of.layer_id[i])
    assert torch.equal(nf.layers[0].data["features"], of.features)
    for i in range(n_layers):
        b, ob = nf.blocks[i], of.blocks[i]
        assert (b.n_src, b.n_dst) == (ob.n_src, ob.n_dst)
        dst = torch.repeat_interleave(torch.arange(b.n_dst), b.rowptr[1:] - b.rowptr[:-1])
        k1 = dst * b.n_src + b.col.long(); k2 = ob.dst * ob.n_src + ob.src
        o1, o2 = torch.argsort(k1), torch.argsort(k2)
        assert torch.equal(k1[o1], k2[o2]) and torch.equal(b.weight[o1], ob.weight[o2])


def test_sampler_batches_shuffle_and_fanout(golden_train):
    z = golden_train
    gg = golden_graph(z)
    g = sd.DeepSortGraph.from_edges(gg.src, gg.dst, gg.weight, gg.node_id, gg.features, gg.num_genes)
    seeds = torch.from_numpy(z["train_ids"])
    gen = torch.Generator().manual_seed(3)
    sampler = sd.NeighborSampler(g, 50, 7, 2, 'in', shuffle=True, seed_nodes=seeds, generator=gen)
    seen = []
    for nf in sampler:
        seen.append(nf.layer_parent_nid(-1))
        for i, b in enumerate(nf.blocks):
            deg = b.rowptr[1:] - b.rowptr[:-1]
            full = (g.in_rowptr[1:] - g.in_rowptr[:-1])[nf.layer_parent_nid(i + 1)]
            assert torch.equal(deg, torch.clamp(full, max=7))            # ≤ fanout, without replacement
            eid = nf.block_parent_eid(i)
            assert eid.unique().shape[0] == eid.shape[0]
            # every sampled edge really is an in-edge of its destination, weight untouched
            dst = torch.repeat_interleave(nf.layer_parent_nid(i + 1), deg)
            assert torch.all((eid >= g.in_rowptr[dst]) & (eid < g.in_rowptr[dst + 1]))
            assert torch.equal(b.weight, g.in_weight[eid])
            assert torch.equal(nf.layer_parent_nid(i)[b.col.long()], g.in_src[eid])
    assert len(seen) == len(sampler) == (len(seeds) + 49) // 50
    allseen = torch.cat(seen)
    assert torch.equal(torch.sort(allseen).values, torch.sort(seeds).values) and not torch.equal(allseen, seeds)


def test_missing_extension_fails_loudly(monkeypatch, tmp_path):
    """No CPU fallback: without libwsage.so the binding raises instead of computing something else."""
    monkeypatch.setattr(sd._lib, "_lib", None)
    monkeypatch.setattr(sd._lib, "LIB_PATH", tmp_path / "libwsage.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        sd._lib.load()


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under scdeepsort_b200/ (nor bench.py's GPU arm) may route through it."""
    for path in (ROOT / "scdeepsort_b200").rglob("*.py"):
        text = path.read_text()
        assert "import oracle" not in text and "from oracle" not in text, path
    bench = (ROOT / "bench.py").read_text()
    gpu_arm = bench[bench.index("def run_ours"):bench.index("def run_sampled")]
    # the only oracle use inside run_ours is the bounded cpu_baseline leg (cpu_reference_steps)
    assert "from oracle" not in gpu_arm and gpu_arm.count("cpu_reference_steps(") == 1


def _csr_matrix(csr):
    import scipy.sparse as sp
    col = csr.col.to(torch.int64)
    if csr.col_bits == 16:
        col = col & 0xFFFF
    return sp.csr_matrix((csr.x.numpy(), col.numpy(), csr.rowptr.numpy()), shape=(csr.n_dst, csr.n_src)).toarray()


def test_densify_splits_expression_matrix_exactly():
    """BipartiteGraph.densify: CSR remainder + the 16-bit block reproduce the expression matrix in both directions
    (fp16 hi + lo of x·2^k carries 22 bits: ≤ 2^-22 relative; host tensors — the kernel that consumes the block is
    covered by tests/test_gpu_dense_block.py)."""
    import scipy.sparse as sp
    from scds_helpers import csr_dense_matrix
    from scdeepsort_b200.graph import BipartiteGraph, DeepSortGraph
    rng = np.random.RandomState(0)
    c, g = 300, 150
    pop = np.minimum(1, 3 * np.arange(1, g + 1) ** -0.8)[rng.permutation(g)]
    mask = rng.rand(c, g) < pop[None, :]
    x = sp.csr_matrix(np.where(mask, rng.rand(c, g) + 0.1, 0).astype(np.float32))
    full = x.toarray()

    bg = BipartiteGraph.from_expression(x).densify(0.2)
    assert bg.densified and 0 < len(bg.dense_genes) < g
    d = bg.cell_csr.dense
    assert d is bg.gene_csr.dense and bg.cell_csr.dense_side == 0 and bg.gene_csr.dense_side == 1      # ONE copy, both directions
    assert d.slots_pad % 256 == 0 and d.hi.numel() == ((c + 127) // 128) * d.slots_pad * 128 == d.lo.numel()
    assert bg.cell_csr.nnz + d.nnz == x.nnz == bg.nnz and bg.gene_csr.nnz == bg.cell_csr.nnz
    np.testing.assert_allclose(_csr_matrix(bg.cell_csr) + csr_dense_matrix(bg.cell_csr).numpy(), full, rtol=2.0 ** -21, atol=0)
    np.testing.assert_allclose(_csr_matrix(bg.gene_csr) + csr_dense_matrix(bg.gene_csr).numpy(), full.T, rtol=2.0 ** -21, atol=0)
    # the dense genes are gone from the CSRs, the others untouched; columns stay sorted inside every remaining row
    dense = np.zeros(g, bool); dense[bg.dense_genes.numpy()] = True
    assert np.array_equal(_csr_matrix(bg.cell_csr)[:, ~dense], full[:, ~dense]) and not _csr_matrix(bg.cell_csr)[:, dense].any()
    rp, col = bg.gene_csr.rowptr.numpy(), (bg.gene_csr.col.to(torch.int64) & 0xFFFF).numpy()
    assert all(np.all(np.diff(col[rp[i]:rp[i + 1]]) > 0) for i in range(g))
    # bf16 planes: one plane, values rounded to 8 bits
    bgb = BipartiteGraph.from_expression(x).densify(0.2, fmt="bf16")
    assert bgb.cell_csr.dense.lo is None and bgb.cell_csr.dense.x_scale == 1.0
    np.testing.assert_allclose(_csr_matrix(bgb.cell_csr) + csr_dense_matrix(bgb.cell_csr).numpy(), full, rtol=2.0 ** -8, atol=0)
    # every gene with an edge in the block: the CSRs are empty and wsage_spmm only reduces
    bga = BipartiteGraph.from_expression(x).densify(0.0)
    assert bga.cell_csr.nnz == 0 == bga.gene_csr.nnz and bga.nnz == x.nnz
    # nothing popular enough: unchanged
    bg3 = BipartiteGraph.from_expression(x).densify(1.1)
    assert not getattr(bg3, "densified", False) and bg3.gene_csr.dense is None
    # the mini-batch graph needs the complete CSRs: refused after densify
    with pytest.raises(ValueError, match="complete CSRs"):
        DeepSortGraph.from_bipartite(bg)


def test_bench_reference_arm_prints_exactly_one_json_line():
    """Driver contract: stdout of bench.py is ONE JSON line (library banners go to stderr); the reference arm
    runs the oracle port on the host cores and needs no GPU."""
    import json
    import subprocess
    import sys
    proc = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                           "--cells", "4000", "--genes", "500", "--deg", "40", "--dim", "16", "--hidden", "16",
                           "--cpu-sample-cells", "48"], capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [l for l in proc.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "cells/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]


def test_new_entry_points_validate_arguments_without_a_gpu():
    """Argument checks of the ABI-1001 additions run before any CUDA call, so they are testable here."""
    import ctypes
    lib = sd._lib.load()
    assert lib.wsage_version() >= 1001
    # accumulation chains of the weight gradient are capped at 1024 rows (64 k-blocks of 16)
    assert lib.wsage_grad_w_splits(780_000, 400) == -(-(-(-780_000 // 16)) // 64) == 762
    assert lib.wsage_grad_w_splits(1000, 400) == 1 and lib.wsage_grad_w_splits(0, 400) == 0
    one = ctypes.c_void_p(16)          # non-null, 16-byte aligned dummy: never dereferenced on these paths
    rc = lib.wsage_grad_w_tc(one, one, 400, one, one, 400, 1000, 398, 400, one, 1, one, 400, None)
    assert rc == sd._lib.EINVAL and b"multiples of 4" in lib.wsage_last_error()
    rc = lib.wsage_grad_w_tc(one, one, 400, one, one, 600, 1000, 400, 600, one, 1, one, 600, None)
    assert rc == sd._lib.EINVAL and b"512" in lib.wsage_last_error()
    rc = lib.wsage_grad_w_tc(one, one, 400, one, one, 400, 100_000, 400, 400, one, 3, one, 400, None)
    assert rc == sd._lib.EINVAL and b"wsage_grad_w_splits" in lib.wsage_last_error()
    a = sd._lib.SpmmArgs()
    a.n_dst, a.n_src, a.dim, a.col_bits, a.nnz = 8, 8, 400, 32, 4
    a.rowptr = a.hs = a.out = one
    a.ld_hs = a.ld_out = 400
    a.init, a.init_slabs, a.init_rows = one, 0, 8
    assert lib.wsage_spmm(ctypes.byref(a), None) == sd._lib.EINVAL and b"init_slabs" in lib.wsage_last_error()
    a.init_slabs, a.init_rows = 2, 5                # init_map == NULL needs init_rows >= n_dst
    assert lib.wsage_spmm(ctypes.byref(a), None) == sd._lib.EINVAL and b"init_map" in lib.wsage_last_error()
    a.init_rows, a.algo = 8, 1                      # the gather kernel cannot take seeded accumulators
    assert lib.wsage_spmm(ctypes.byref(a), None) == sd._lib.EINVAL and b"tiled" in lib.wsage_last_error()
    assert lib.wsage_spmm_algo(ctypes.byref(a)) == 0
    a.algo = 0
    assert lib.wsage_spmm_algo(ctypes.byref(a)) == 2                  # seeded accumulators always take the tiled kernel
    a.nnz = 0                                        # empty CSR + init: a plain reduction, no workspace
    assert lib.wsage_spmm_workspace_bytes(ctypes.byref(a)) == 0


def test_dense16_entry_points_validate_arguments_without_a_gpu():
    """ABI 2000: wsage_amax / wsage_split16 / wsage_dense16 reject bad arguments before any CUDA call."""
    import ctypes
    lib = sd._lib.load()
    assert lib.wsage_version() >= 2000
    assert lib.wsage_dense16_slots_pad(1) == 256 and lib.wsage_dense16_slots_pad(257) == 512 and lib.wsage_dense16_slots_pad(0) == 0
    one = ctypes.c_void_p(16)
    assert lib.wsage_amax(one, 400, None, None, 10, 398, one, None) == sd._lib.EINVAL and b"multiple of 4" in lib.wsage_last_error()
    assert lib.wsage_amax(one, 400, None, None, 10, 400, None, None) == sd._lib.EINVAL
    rc = lib.wsage_split16(one, 400, None, None, 10, 400, one, sd._lib.D16_F16X2, 0, one, one, 404, None)
    assert rc == sd._lib.EINVAL and b"ld_out % 8" in lib.wsage_last_error()
    rc = lib.wsage_split16(one, 400, one, None, 10, 400, one, sd._lib.D16_F16X2, 0, one, one, 400, None)
    assert rc == sd._lib.EINVAL and b"row_ids needs a transposed layout" in lib.wsage_last_error()
    rc = lib.wsage_split16(one, 400, None, None, 10, 400, one, sd._lib.D16_F16X2, sd._lib.SPLIT_COLBLOCKS, one, one, 9, None)
    assert rc == sd._lib.EINVAL and b"ld_out too small" in lib.wsage_last_error()
    rc = lib.wsage_split16(one, 400, None, None, 10, 400, one, 7, 0, one, one, 400, None)
    assert rc == sd._lib.EINVAL and b"fmt" in lib.wsage_last_error()
    a = sd._lib.Dense16Args()
    a.x_hi = a.x_lo = a.h_hi = a.h_lo = a.out = one
    a.fmt, a.cells, a.gene_slots, a.x_scale, a.dim, a.ld_h = sd._lib.D16_F16X2, 1000, 200, 1024.0, 400, 400
    a.side, a.n_src_cells = 1, 1000
    n = lib.wsage_dense16_splits(ctypes.byref(a))
    assert n == 1                                   # 1000 cells = one chain of 2048 rows
    a.cells = a.n_src_cells = 760_000
    a.gene_slots = 20_000
    n = lib.wsage_dense16_splits(ctypes.byref(a))
    assert 2 <= n <= 64                              # (tile, split) items fill the 148 SMs in whole rounds
    a.n_src_cells = 760_001
    assert lib.wsage_dense16_splits(ctypes.byref(a)) == 0 and b"n_src_cells" in lib.wsage_last_error()
    a.n_src_cells, a.dim = 760_000, 516
    assert lib.wsage_dense16(ctypes.byref(a), None) == sd._lib.EINVAL and b"512" in lib.wsage_last_error()
    a.dim, a.side, a.n_dst, a.ld_h, a.ld_out = 400, 0, 760_000, 384, 400
    assert lib.wsage_dense16(ctypes.byref(a), None) == sd._lib.EINVAL and b"ld_h" in lib.wsage_last_error()
    a.ld_h, a.selfcoef = 400, one
    assert lib.wsage_dense16(ctypes.byref(a), None) == sd._lib.EINVAL and b"hself" in lib.wsage_last_error()
    a.selfcoef, a.x_scale = None, 0.0
    assert lib.wsage_dense16(ctypes.byref(a), None) == sd._lib.EINVAL and b"x_scale" in lib.wsage_last_error()


def test_densify_with_test_cells_splits_all_three_csrs():
    """Inference graphs carry test cells (gene->cell only, utils/preprocess.py:185-187): gene_csr covers the support
    cells, cell_csr_t is the transpose of the full cell_csr; densify strips all three and they share one block
    that covers every cell (the gene-destination side of gene_csr reads only its first num_support cells)."""
    import scipy.sparse as sp
    from scds_helpers import csr_dense_matrix
    from scdeepsort_b200.graph import BipartiteGraph
    rng = np.random.RandomState(5)
    cs, ct, g = 120, 30, 90
    pop = np.minimum(1, 2.5 * np.arange(1, g + 1) ** -0.8)[rng.permutation(g)]
    xs = sp.csr_matrix(np.where(rng.rand(cs, g) < pop[None, :], rng.rand(cs, g) + 0.1, 0).astype(np.float32))
    xt = sp.csr_matrix(np.where(rng.rand(ct, g) < pop[None, :], rng.rand(ct, g) + 0.1, 0).astype(np.float32))
    bg = BipartiteGraph.from_expression(xs, xt).densify(0.25)
    assert bg.densified and bg.cell_csr.dense is bg.gene_csr.dense is bg.cell_csr_t.dense
    assert bg.cell_csr.dense.cells == cs + ct
    full = sp.vstack([xs, xt]).toarray()
    tol = dict(rtol=2.0 ** -21, atol=0)
    np.testing.assert_allclose(_csr_matrix(bg.gene_csr) + csr_dense_matrix(bg.gene_csr).numpy(), xs.toarray().T, **tol)
    np.testing.assert_allclose(_csr_matrix(bg.cell_csr_t) + csr_dense_matrix(bg.cell_csr_t).numpy(), full.T, **tol)
    np.testing.assert_allclose(_csr_matrix(bg.cell_csr) + csr_dense_matrix(bg.cell_csr).numpy(), full, **tol)
    assert bg.transpose_of_cell_csr() is bg.cell_csr_t
    # the support rows of cell_csr (backward of the gene aggregation) keep the block and a consistent nnz
    sup = bg.support_cell_csr()
    assert sup.n_dst == cs and sup.dense is bg.cell_csr.dense and sup.nnz == int(bg.cell_csr.rowptr[cs])
    np.testing.assert_allclose(_csr_matrix(sup) + csr_dense_matrix(sup).numpy(), xs.toarray(), **tol)
    assert bg.support_cell_csr() is sup


def test_peer_reduce_rejects_bad_arguments_before_touching_the_device():
    """ABI 2001: wsage_peer_reduce validates on the host (no GPU here: every call must fail with EINVAL, not crash)."""
    import ctypes
    lib = sd._lib.load()
    assert lib.wsage_version() >= 2001
    assert lib.wsage_peer_bytes(0) == 0 and lib.wsage_peer_bytes(1000) == 4096 + 4 * 4096
    a = sd._lib.PeerReduceArgs()
    assert lib.wsage_peer_reduce(None, None) == sd._lib.EINVAL
    a.rank, a.world = 0, 9
    assert lib.wsage_peer_reduce(ctypes.byref(a), None) == sd._lib.EINVAL and b"world" in lib.wsage_last_error()
    a.world, a.rows, a.dim, a.max_elems, a.epoch = 2, 10, 6, 100, 1
    assert lib.wsage_peer_reduce(ctypes.byref(a), None) == sd._lib.EINVAL and b"dim % 4" in lib.wsage_last_error()
    a.dim, a.epoch = 8, 2
    assert lib.wsage_peer_reduce(ctypes.byref(a), None) == sd._lib.EINVAL and b"epoch" in lib.wsage_last_error()
    a.epoch = 3
    assert lib.wsage_peer_reduce(ctypes.byref(a), None) == sd._lib.EINVAL and b"null" in lib.wsage_last_error()


def test_peer_exchange_is_off_without_a_process_group_or_a_gpu():
    """parallel.enable_peer_exchange / peer.enable are collective set-up calls: outside torch.distributed, or for a graph that
    is not on a GPU, they leave the NCCL / gloo all-reduce in place and say so by returning None."""
    from scdeepsort_b200 import parallel, peer
    from scdeepsort_b200.synthetic import synthetic_bipartite
    bg = synthetic_bipartite(60, 40, 8, device="cpu")
    assert parallel.enable_peer_exchange(bg, 64) is None and bg.peer_group is None
    assert peer.enable(1000) is None and peer.active() is None
    peer.disable()                                    # nothing to tear down: a no-op
    header = (ROOT / "include" / "wsage.h").read_text()
    assert f"#define WSAGE_PEER_MAX {sd._lib.PEER_MAX}" in header
