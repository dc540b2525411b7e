"""Diagnostic for wsage_dense16 (run on a GPU box): both sides on small random blocks against fp64, one line per case,
never asserts — a first look at a new build before the pytest suite.  `--time` adds c3/c4-shaped timings."""
import argparse
import sys
import time
import traceback
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

import numpy as np  # noqa: E402
import scipy.sparse as sp  # noqa: E402
import torch  # noqa: E402

import scdeepsort_b200 as sd  # noqa: E402
from scdeepsort_b200 import ops  # noqa: E402
from scds_helpers import dense_block_matrix, rel_err  # noqa: E402

DEV = "cuda:0"


def expr(n_cells, n_genes, seed, dens=0.3):
    rng = np.random.RandomState(seed)
    m = rng.rand(n_cells, n_genes) < dens
    return sp.csr_matrix(np.where(m, rng.uniform(0.05, 9, (n_cells, n_genes)), 0).astype(np.float32))


def case(n_cells, n_genes, dim, chunk, fmt):
    x = expr(n_cells, n_genes, n_cells + dim)
    bg = sd.BipartiteGraph.from_expression(x, device=DEV).densify(0.0, fmt=fmt)
    d = bg.cell_csr.dense
    xd = dense_block_matrix(d).double()
    g = torch.Generator(device=DEV).manual_seed(3)
    ids = d.gene_ids.cpu().long()
    hg = torch.randn(n_genes, dim, device=DEV, generator=g) * 3.0
    hc = torch.randn(n_cells, dim, device=DEV, generator=g)
    res = {}
    for name, fn in (
            ("side0", lambda: rel_err(ops.dense16(d, 0, hg, n_dst=n_cells, chunk_rows=chunk).cpu(), xd @ hg.double().cpu()[ids])),
            ("side1", lambda: rel_err(ops.dense16(d, 1, hc, n_src_cells=n_cells, chunk_rows=chunk).sum(0)[:d.gene_slots].cpu(),
                                      xd.t() @ hc.double().cpu()))):
        try:
            res[name] = f"{fn():.2e}"
            torch.cuda.synchronize()
        except Exception as e:       # noqa: BLE001
            res[name] = f"EXC {type(e).__name__}: {str(e)[:200]}"
            traceback.print_exc()
    print(f"cells {n_cells:6d} genes {n_genes:5d} dim {dim:3d} chunk {chunk:4d} {fmt:6s} slots {d.gene_slots:5d} -> {res}", flush=True)


def timing(n_cells, n_genes, deg, dim, thr, fmt, chunk):
    from scdeepsort_b200.synthetic import synthetic_bipartite
    t0 = time.time()
    bg = synthetic_bipartite(n_cells, n_genes, deg, device=DEV)
    nnz = bg.cell_csr.nnz
    bg.densify(thr, fmt=fmt)
    torch.cuda.synchronize()
    d = bg.cell_csr.dense
    print(f"[{n_cells}x{n_genes} thr {thr} {fmt}] build {time.time() - t0:.1f}s, {d.gene_slots} dense genes hold {d.nnz / nnz:.3f} of the edges, "
          f"csr left {bg.cell_csr.nnz}", flush=True)
    g = torch.Generator(device=DEV).manual_seed(1)
    hg = torch.randn(n_genes, dim, device=DEV, generator=g)
    hc = torch.randn(n_cells, dim, device=DEV, generator=g)
    for name, fn in (("dense16 side0", lambda: ops.dense16(d, 0, hg, n_dst=n_cells, chunk_rows=chunk)),
                     ("dense16 side1", lambda: ops.dense16(d, 1, hc, n_src_cells=n_cells, chunk_rows=chunk)),
                     ("spmm cell<-gene", lambda: sd.spmm(bg.cell_csr, hg)),
                     ("spmm gene<-cell", lambda: sd.spmm(bg.gene_csr, hc))):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        terms = 3 if fmt == "f16x2" else 1
        tf = terms * 2 * n_cells * d.slots_pad * dim / (min(ts) * 1e-3) / 1e12
        print(f"   {name:16s} best {min(ts):8.3f} ms  mean {sum(ts) / len(ts):8.3f} ms   (block MMA rate if dense16: {tf:7.1f} TFLOP/s)", flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--time", action="store_true")
    ap.add_argument("--cells", type=int, default=100_000)
    ap.add_argument("--thr", default="0.0")
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--fmt", default="f16x2")
    a = ap.parse_args()
    if not a.time:
        for c in [(128, 32, 64, 0, "f16x2"), (256, 128, 128, 0, "f16x2"), (1000, 700, 400, 0, "f16x2"), (333, 257, 400, 64, "f16x2"),
                  (5000, 300, 200, 512, "f16x2"), (640, 500, 132, 96, "f16x2"), (150, 90, 512, 32, "f16x2"),
                  (1000, 700, 400, 0, "bf16"), (40000, 160, 400, 0, "f16x2"), (40000, 160, 400, 4096, "f16x2"), (40000, 160, 400, 8192, "f16x2")]:
            case(*c)
    else:
        for thr in [float(t) for t in a.thr.split(",")]:
            timing(a.cells, 20_000, 2000, 400, thr, a.fmt, a.chunk)
