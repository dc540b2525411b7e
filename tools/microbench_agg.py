#!/usr/bin/env python
"""Kernel-level micro-benchmark of wsage_spmm (gather vs tiled) on a synthetic atlas.
Run on the GPU box:  python tools/microbench_agg.py --cells 100000 --genes 20000 --deg 2000 --dim 400"""
import argparse
import json
import time

import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import scdeepsort_b200 as sd
from scdeepsort_b200.synthetic import synthetic_bipartite


def timed(fn, iters, flush):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()                       # evict L2 between timed launches
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), sum(ts) / len(ts)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=100000)
    ap.add_argument("--genes", type=int, default=20000)
    ap.add_argument("--deg", type=float, default=2000)
    ap.add_argument("--dim", type=int, default=400)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--algos", default="1,2")
    ap.add_argument("--dense", type=float, default=0.0, help="> 0: also time the densified graph (popular genes as a dense block)")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    t0 = time.time()
    bg = synthetic_bipartite(args.cells, args.genes, args.deg, device=dev)
    torch.cuda.synchronize()
    print(f"graph: C={bg.num_cells} G={bg.num_genes} nnz={bg.nnz} built in {time.time()-t0:.1f}s", flush=True)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    hg = torch.randn(bg.num_genes, args.dim, device=dev)
    hc = torch.randn(bg.num_cells, args.dim, device=dev)
    res = {}
    for name, csr, hs, hself in (("cell<-gene", bg.cell_csr, hg, hc), ("gene<-cell", bg.gene_csr, hc, hg)):
        outs = {}
        for algo in [int(a) for a in args.algos.split(",")]:
            dscale = torch.rand(csr.n_dst, device=dev) + 0.5
            selfc = torch.rand(csr.n_dst, device=dev)
            torch.manual_seed(0)
            fn = lambda: sd.spmm(csr, hs, dscale=dscale, selfcoef=selfc, hself=hself, algo=algo)[0]  # noqa: E731
            best, mean = timed(fn, args.iters, flush)
            outs[algo] = (fn(), dscale, selfc)
            i_bytes = 2 if csr.col_bits == 16 else 4
            alg_bytes = csr.nnz * (i_bytes + 4) + (csr.n_dst + 1) * 8 + csr.n_src * args.dim * 4 + 2 * csr.n_dst * args.dim * 4
            gather_bytes = csr.nnz * args.dim * 4
            res[f"{name}/algo{algo}"] = dict(ms_best=round(best, 3), ms_mean=round(mean, 3),
                                             gedges_per_s=round(csr.nnz / best / 1e6, 2),
                                             algorithmic_GBs=round(alg_bytes / best / 1e6, 1),
                                             gather_TBs=round(gather_bytes / best / 1e9, 2))
            print(name, "algo", algo, res[f"{name}/algo{algo}"], flush=True)
        if len(outs) == 2:
            (o1, d1, s1), (o2, d2, s2) = outs[1], outs[2]
            # same epilogue inputs? (they are re-drawn per algo) -> compare raw sums instead
            r1 = sd.spmm(csr, hs, algo=1)[0]; r2 = sd.spmm(csr, hs, algo=2)[0]
            err = float((r1 - r2).abs().max() / r1.abs().max())
            print(name, "gather vs tiled rel err", err, flush=True)
            res[f"{name}/agree"] = err
    if args.dense > 0:
        plain = {"cell<-gene": sd.spmm(bg.cell_csr, hg)[0], "gene<-cell": sd.spmm(bg.gene_csr, hc)[0]}
        t0 = time.time()
        bg.densify(args.dense)
        torch.cuda.synchronize()
        print(f"densify({args.dense}): {len(bg.dense_genes)} genes, {bg.cell_csr.dense.nnz / max(1, bg.nnz):.3f} "
              f"of the edges, {time.time()-t0:.1f}s", flush=True)
        for name, csr, hs, hself in (("cell<-gene", bg.cell_csr, hg, hc), ("gene<-cell", bg.gene_csr, hc, hg)):
            dscale = torch.rand(csr.n_dst, device=dev) + 0.5
            selfc = torch.rand(csr.n_dst, device=dev)
            fn = lambda: sd.spmm(csr, hs, dscale=dscale, selfcoef=selfc, hself=hself)[0]  # noqa: E731
            best, mean = timed(fn, args.iters, flush)
            err = float((sd.spmm(csr, hs)[0] - plain[name]).abs().max() / plain[name].abs().max())
            res[f"{name}/dense{args.dense}"] = dict(ms_best=round(best, 3), ms_mean=round(mean, 3), rel_err_vs_plain=err)
            print(name, "densified", res[f"{name}/dense{args.dense}"], flush=True)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
