#!/usr/bin/env python
"""HBM-bound regime of the aggregation: sampled blocks whose source rows are (almost) all distinct and come
from a table far larger than L2 (gene destinations sampling `fanout` cells each from the 760k-cell table).
Reports achieved HBM GB/s of wsage_block_agg_fwd / _bwd against the measured copy peak."""
import argparse
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import scdeepsort_b200 as sd  # noqa: E402
from scdeepsort_b200.nodeflow import _sample_edges_cuda  # noqa: E402
from scdeepsort_b200.synthetic import synthetic_bipartite  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=760000)
    ap.add_argument("--genes", type=int, default=20000)
    ap.add_argument("--deg", type=float, default=2000)
    ap.add_argument("--dim", type=int, default=400)
    ap.add_argument("--fanouts", default="10,32")
    ap.add_argument("--reps", type=int, default=8, help="independent sampled blocks aggregated back to back per timing")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    peaks = json.loads((Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").read_text()) \
        if (Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").exists() else {"hbm_gbs": 6650.0}
    bg = synthetic_bipartite(args.cells, args.genes, args.deg, device=dev)
    g = bg.num_genes
    # parent "graph" = the gene-major CSR (in-edges of genes come from cells)
    parent = sd.DeepSortGraph(g, 0, bg.gene_csr.rowptr, bg.gene_csr.col, bg.gene_csr.x, {})
    hc = torch.randn(bg.num_cells, args.dim, device=dev)
    alpha = (torch.rand(g + 2, 1, device=dev) + 0.5).requires_grad_(True)
    src_id = torch.full((bg.num_cells,), -1, dtype=torch.int32, device=dev)
    dst_id = torch.arange(g, dtype=torch.int32, device=dev)
    nodes = torch.arange(g, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out = {}
    for f in [int(v) for v in args.fanouts.split(",")]:
        blocks = []
        for rep in range(args.reps):
            eid, deg = _sample_edges_cuda(parent, nodes, f, seed=100 + rep)
            rowptr = torch.zeros(g + 1, dtype=torch.int64, device=dev)
            rowptr[1:] = torch.cumsum(deg, 0)
            blocks.append(sd.Block(rowptr, bg.gene_csr.col[eid].to(torch.int32), bg.gene_csr.x[eid], bg.num_cells, g))
        e = sum(int(b.col.shape[0]) for b in blocks)
        distinct = sum(int(torch.unique(b.col).shape[0]) for b in blocks)
        h = hc.clone().requires_grad_(True)
        for mode in ("fwd", "bwd"):
            best = 1e9
            for _ in range(4):
                outs = [sd.block_aggregate(h, alpha, b, src_id, dst_id, g) for b in blocks] if mode == "bwd" else None
                gout = [torch.ones_like(o) for o in outs] if outs else None
                flush.zero_()
                a, bb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                if mode == "fwd":
                    with torch.no_grad():
                        for b in blocks:
                            sd.block_aggregate(h, alpha, b, src_id, dst_id, g)
                else:
                    torch.autograd.backward(outs, gout)
                bb.record(); torch.cuda.synchronize()
                best = min(best, a.elapsed_time(bb))
                h.grad = None; alpha.grad = None
            # algorithmic bytes: edges (col 4 + w 4), distinct source rows once, destination rows written (fwd) / read (bwd);
            # bwd additionally read-modify-writes dH for every edge row (atomics) and re-reads H for d_alpha
            row = args.dim * 4
            bytes_ = e * 8 + distinct * row + args.reps * g * row + (e * row * 2 if mode == "bwd" else 0)
            out[f"fanout{f}/{mode}"] = dict(ms=round(best, 4), edges=e, distinct_src_rows=distinct,
                                            algorithmic_GBs=round(bytes_ / best / 1e6, 1),
                                            frac_of_hbm_peak=round(bytes_ / best / 1e6 / peaks["hbm_gbs"], 3))
            print(f"fanout {f} {mode}", out[f"fanout{f}/{mode}"], flush=True)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
