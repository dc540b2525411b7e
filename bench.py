#!/usr/bin/env python
"""Benchmark of the weighted-GraphSAGE hot path (BASELINE.json metric: cells/sec, forward+backward,
760k-cell × 20k-gene synthetic atlas, 2-layer hidden=400).

  python bench.py [--gpus N] [--steps K] [--warmup W]            our CUDA path (one JSON line)
  python bench.py --impl reference [...]                          the CPU restatement of the reference
  torchrun --nproc-per-node N bench.py --gpus N ...               cell-sharded, NCCL

A "step" = one full-graph training step over every cell of the atlas: 2-layer forward (3 aggregation
passes + 2 Linear/ReLU + classifier), CE(sum), backward (2 transposed aggregation passes, dα row-dots,
dense backward) and Adam.  Strong scaling: the atlas is fixed, cells are sharded over ranks.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402
import torch  # noqa: E402

SEED = 10086
NUM_CLASSES = 16


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="default 5 (full-graph step), 30 (sampled mini-batches)")
    ap.add_argument("--warmup", type=int, default=None, help="default 3 (full-graph step), 10 (sampled: batch shapes vary, the allocator "
                    "needs a few batches to stop calling cudaMalloc)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells", type=int, default=760_000)
    ap.add_argument("--genes", type=int, default=20_000)
    ap.add_argument("--deg", type=float, default=2000.0)
    ap.add_argument("--dim", type=int, default=400)
    ap.add_argument("--hidden", type=int, default=400)
    ap.add_argument("--layers", type=int, default=2)
    ap.add_argument("--algo", type=int, default=0, help="wsage_spmm algo (0 auto, 1 gather, 2 tiled)")
    ap.add_argument("--dense-threshold", type=float, default=0.0,
                    help="genes expressed in at least this share of the cells leave the CSRs and run on the tensor-core "
                         "dense-block kernel (BipartiteGraph.densify); 0 = every gene, negative = CSR only")
    ap.add_argument("--dense-fmt", default="f16x2", choices=["f16x2", "bf16"],
                    help="f16x2 = fp16 hi+lo planes, three products (fp32-grade); bf16 = one product (BASELINE configs[2])")
    ap.add_argument("--dropout", type=float, default=0.0)
    ap.add_argument("--config", default="c4", choices=["c1", "c3", "c4"],
                    help="BASELINE.json configs: c4 (default, the metric's configuration) 760k x 20k fp32; c3 = 100k x 20k, bf16 planes "
                         "and single-product tensor-core Linear; c1 = 1k x 2k, hidden 64 (the reference's CPU-runnable case)")
    ap.add_argument("--cpu-sample-cells", type=int, default=512)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--mode", default="full", choices=["full", "sampled"],
                    help="full = BASELINE metric (full-graph step); sampled = configs[4]: neighbour-sampled mini-batches")
    ap.add_argument("--fanouts", default="25,10,5", help="sampled mode: per-hop fan-outs, seed hop first")
    ap.add_argument("--batch", type=int, default=1024, help="sampled mode: seed cells per GPU per step")
    a = ap.parse_args()
    if a.steps is None:
        a.steps = 30 if a.mode == "sampled" else 5
    if a.warmup is None:
        a.warmup = 10 if a.mode == "sampled" else 3
    if a.config == "c3":
        a.cells, a.genes, a.deg, a.dim, a.hidden, a.dense_fmt = 100_000, 20_000, 2000.0, 400, 400, "bf16"
    elif a.config == "c1":
        a.cells, a.genes, a.deg, a.dim, a.hidden = 1000, 2000, 200.0, 64, 64
        a.cpu_sample_cells = min(a.cpu_sample_cells, 1000)
    return a


def workload_name(a):
    return (f"synthetic {a.cells} cells x {a.genes} genes, avg-degree {int(a.deg)}, {a.layers}-layer "
            f"hidden={a.hidden} dense_dim={a.dim}, full-neighbour full-graph training step (fwd+bwd+Adam)")


# ----------------------------------------------------------------------------------------------
# CPU arm: the oracle's literal (edge-materialising) restatement of the reference on a bounded sample
# ----------------------------------------------------------------------------------------------
def cpu_reference_steps(a, n_cells, steps, warmup):
    """Full-graph fwd+bwd+Adam on the first ``n_cells`` cells of the same atlas (same generator, same
    genes / degree / widths) with the oracle (message tensor materialised as models/gnn.py:54-56 does),
    all host threads.  Returns seconds per step."""
    import scipy.sparse as sp
    from oracle import gnn_oracle, graph_oracle
    from scdeepsort_b200.synthetic import synthetic_bipartite, synthetic_features
    torch.set_num_threads(os.cpu_count())
    bg = synthetic_bipartite(a.cells, a.genes, a.deg, seed=SEED, device="cpu", cell_range=(0, n_cells))
    cs = bg.cell_csr
    col = cs.col.numpy().view(np.uint16).astype(np.int64) if cs.col_bits == 16 else cs.col.numpy()
    x = sp.csr_matrix((cs.x.numpy(), col, cs.rowptr.numpy()), shape=(n_cells, a.genes))
    og = graph_oracle.build_graph(x)
    og.features = synthetic_features(bg, a.dim, seed=SEED)
    seeds = torch.arange(og.num_genes, og.num_nodes)
    flow = graph_oracle.full_neighbor_flow(og, seeds, a.layers)
    params = gnn_oracle.init_params(a.dim, a.hidden, NUM_CLASSES, a.layers, a.genes, seed=SEED)
    params = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    opt = torch.optim.Adam(list(params.values()), lr=1e-3, weight_decay=5e-4)
    labels = torch.randint(0, NUM_CLASSES, (n_cells,), generator=torch.Generator().manual_seed(SEED))
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        logits = gnn_oracle.forward(params, flow, a.genes)
        loss = torch.nn.functional.cross_entropy(logits, labels, reduction="sum")
        opt.zero_grad()
        loss.backward()
        opt.step()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return sum(times) / len(times)


def cpu_spmm_steps(a, n_cells, steps, warmup):
    """The non-strawman CPU number of SURVEY 8(d): same step, closed form on torch sparse CSR products
    (oracle/spmm_oracle.py, no per-edge message tensor), first ``n_cells`` cells of the atlas, all host threads."""
    import warnings
    import scipy.sparse as sp
    from oracle import gnn_oracle, spmm_oracle
    from scdeepsort_b200.synthetic import synthetic_bipartite, synthetic_features
    warnings.filterwarnings("ignore", message=".*[Ss]parse.*")
    torch.set_num_threads(os.cpu_count())
    bg = synthetic_bipartite(a.cells, a.genes, a.deg, seed=SEED, device="cpu", cell_range=(0, n_cells))
    cs = bg.cell_csr
    col = cs.col.numpy().view(np.uint16).astype(np.int64) if cs.col_bits == 16 else cs.col.numpy()
    x = sp.csr_matrix((cs.x.numpy(), col, cs.rowptr.numpy()), shape=(n_cells, a.genes))
    graph = spmm_oracle.SpmmGraph(x)
    feats = synthetic_features(bg, a.dim, seed=SEED)
    params = gnn_oracle.init_params(a.dim, a.hidden, NUM_CLASSES, a.layers, a.genes, seed=SEED)
    params = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    opt = torch.optim.Adam(list(params.values()), lr=1e-3, weight_decay=5e-4)
    labels = torch.randint(0, NUM_CLASSES, (n_cells,), generator=torch.Generator().manual_seed(SEED))
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        loss = torch.nn.functional.cross_entropy(spmm_oracle.forward(params, graph, feats, a.layers), labels, reduction="sum")
        opt.zero_grad()
        loss.backward()
        opt.step()
        if it >= warmup:
            times.append(time.perf_counter() - t0)
    return sum(times) / len(times)


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = min(a.cpu_sample_cells, a.cells)
    sec = cpu_reference_steps(a, n, a.steps, a.warmup)
    value = n / sec
    sample = (f"first {n} cells of the atlas (same generator, {a.genes} genes, avg-degree {int(a.deg)}), one "
              f"full-graph fwd+bwd+Adam step per timed step, oracle port of models/gnn.py (DGL 0.4.3 not installable)")
    # the non-strawman CPU number travels with every reference line: same step in closed form on sparse-CSR products
    n2 = min(8 * n, a.cells)
    sec2 = cpu_spmm_steps(a, n2, 1, 1)
    optimised = {"value": n2 / sec2, "unit": "cells/s", "cores": os.cpu_count(), "kind": "port, closed form",
                 "sample": f"first {n2} cells, same step without the per-edge message tensor (oracle/spmm_oracle.py), 1 warm-up + 1 timed step"}
    print(json.dumps({
        "impl": "reference", "metric": "cells/sec (forward+backward)", "value": value, "unit": "cells/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a)},
        "cpu_baseline": {"value": value, "unit": "cells/s", "cores": os.cpu_count(), "kind": "port", "sample": sample,
                         "optimised": optimised},
        "e2e": {"value": value, "unit": "cells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([f.strip() for f in out.split(",")])
            except Exception:
                pass
            self._stop_evt.wait(0.2)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=5)
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples for n, v in zip(names, s[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": float(self.samples[0][1]) if self.samples else None,
                "reasons": reasons, "samples": len(self.samples)}


def algorithmic_bytes(t):
    """E·(i+w) + (N_dst+1)·8 + N_src·D·s + N_dst·D·s per [N_dst, D] row stream read or
    written (self rows, q rows, out, raw); distinct source rows counted once, gather re-reads not counted."""
    s = 4
    idx = 2 if t["col_bits"] == 16 else 4
    streams = t["n_out"] + (1 if t["self"] else 0) + (1 if t["dot"] else 0)
    return t["nnz"] * (idx + 4) + (t["n_dst"] + 1) * 8 + t["n_src"] * t["dim"] * s + streams * t["n_dst"] * t["dim"] * s


def dense16_work(t):
    """What one wsage_dense16 launch does.  flops: the MMAs it issues (3 products for fp16 hi+lo, zeros included);
    useful_flops: 2·D per expression entry the block stands for; plane_bytes: the X planes it streams from HBM once
    (hi + lo, 2 bytes each per (cell, slot)); algorithmic_bytes: SURVEY §8(d) for the same entries as a CSR
    (6 B per entry: uint16 index + fp32 value) + the source and destination rows once."""
    terms = 3 if t["fmt"] == 0 else 1
    cells_pad = -(-t["cells"] // 128) * 128
    n_pad = -(-t["dim"] // 16) * 16
    k_side0 = -(-t["gene_slots"] // 32) * 32
    pairs = cells_pad * (k_side0 if t["side"] == 0 else t["slots_pad"])
    n_src, n_dst = (t["gene_slots"], t["cells"]) if t["side"] == 0 else (t["cells"], t["gene_slots"])
    return dict(flops=terms * 2 * pairs * n_pad, useful_flops=2 * t["dense_nnz"] * t["dim"],
                plane_bytes=(2 if terms == 3 else 1) * 2 * t["cells"] * t["slots_pad"],
                algorithmic_bytes=t["dense_nnz"] * 6 + (n_src + n_dst) * t["dim"] * 4)


def run_ours(a):
    import torch.distributed as dist
    import scdeepsort_b200 as sd
    from scdeepsort_b200 import ops, parallel
    from scdeepsort_b200.synthetic import synthetic_bipartite, synthetic_features
    from scdeepsort_b200.trainer import FullGraphTrainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus:
        raise SystemExit(f"--gpus {a.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {a.gpus}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    sd._lib.load()      # fail loudly if the CUDA extension is missing
    if a.dense_fmt == "bf16":
        sd.dense.single_product = True      # configs[2]: bf16 aggregation planes, plain-tf32 Linear layers (fp32 accumulation)

    lo, hi = parallel.cell_ranges(a.cells, world)[rank]
    t0 = time.time()
    graph = synthetic_bipartite(a.cells, a.genes, a.deg, seed=SEED, device=dev, cell_range=(lo, hi))
    parallel.globalize_gene_normalisers(graph)
    feats = synthetic_features(graph, a.dim, seed=SEED)
    labels = torch.randint(0, NUM_CLASSES, (a.cells,), generator=torch.Generator().manual_seed(SEED))[lo:hi].to(dev)
    if a.dense_threshold >= 0:
        graph.densify(a.dense_threshold, fmt=a.dense_fmt)
    torch.cuda.synchronize()
    build_s = time.time() - t0

    trainer = FullGraphTrainer(graph, NUM_CLASSES, dense_dim=a.dim, hidden_dim=a.hidden, n_layers=a.layers,
                               dropout=a.dropout, seed=SEED, sharded=world > 1, spmm_algo=a.algo)
    parallel.broadcast_params(trainer.model)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t)

    # ---- device-resident run: `value` --------------------------------------------------------
    for _ in range(a.warmup):
        trainer.step(feats, labels, return_loss=False)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    ops.TIMING = []
    sd._lib.launch_count(reset=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.profiler.start()          # ncu --profile-from-start off captures exactly the timed region
    e0.record()
    loss = None
    for _ in range(a.steps):
        loss = trainer.step(feats, labels, return_loss=False)
    e1.record()
    barrier()
    torch.cuda.profiler.stop()
    ms_step = max_over_ranks(e0.elapsed_time(e1) / a.steps)
    launches = sd._lib.launch_count()
    timing, ops.TIMING = ops.TIMING, None
    clocks = sampler.stop() if sampler else None
    final_loss = float(loss)

    # ---- roofline of the dominant aggregation kernel (per-launch CUDA events, timed region) ----
    peaks = {}
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peaks = json.loads(pk.read_text())
    hbm_peak, peak_src = (peaks["hbm_gbs"], "measured") if "hbm_gbs" in peaks else (6650.0, "fallback")
    # a kernel timed inside a long step: the sustained tensor figure
    tc_peak = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 1400.0))
    groups = {}
    for t in timing:
        ms = t["events"][0].elapsed_time(t["events"][1])
        if t["kind"] == "dense16":
            key = ("dense16", "cell<-gene" if t["side"] == 0 else "gene<-cell")
            w = dense16_work(t)
            g = groups.setdefault(key, dict(ms=0.0, n=0, bytes=0, flops=0, useful_flops=0, plane_bytes=0, nnz=0))
            g["flops"] += w["flops"]; g["useful_flops"] += w["useful_flops"]; g["plane_bytes"] += w["plane_bytes"]
            g["bytes"] += w["algorithmic_bytes"]; g["nnz"] += t["dense_nnz"]
        else:
            key = ({2: "tiled", 1: "gather", 0: "reduce"}[t["algo"]], "gene<-cell" if t["n_dst"] == a.genes else "cell<-gene")
            g = groups.setdefault(key, dict(ms=0.0, n=0, bytes=0, nnz=0))
            g["bytes"] += algorithmic_bytes(t); g["nnz"] += t["nnz"]
        g["ms"] += ms
        g["n"] += 1
    roofline, kernels = None, {}
    if groups:
        for key, g in groups.items():
            k = {"launches": g["n"], "ms_per_launch": g["ms"] / g["n"], "share_of_step": g["ms"] / a.steps / ms_step,
                 "algorithmic_gbs": g["bytes"] / g["ms"] / 1e6, "entries_per_launch": g["nnz"] / g["n"]}
            if key[0] == "dense16":
                k.update(mma_tflops=g["flops"] / g["ms"] / 1e9, useful_tflops=g["useful_flops"] / g["ms"] / 1e9,
                         hbm_plane_stream_gbs=g["plane_bytes"] / g["ms"] / 1e6)
            kernels[f"{key[0]}:{key[1]}"] = k
        key, g = max(groups.items(), key=lambda kv: kv[1]["ms"])
        traffic = None          # ncu dram bytes per launch, recorded under profiles/ for the c4 shape
        tj = ROOT / "profiles" / "r02_traffic.json"
        if tj.exists() and (a.cells, a.genes, int(a.deg), a.dim, world) == (760_000, 20_000, 2000, 400, 1):
            traffic = json.loads(tj.read_text()).get("c4", {}).get(f"{key[0]}:{key[1]}", {}).get("dram_bytes_per_launch")
        if key[0] == "dense16":
            achieved = g["flops"] / g["ms"] / 1e9
            roofline = {"kernel": f"dense16_kernel ({key[1]})", "bound": "tensor", "achieved": achieved, "peak": tc_peak,
                        "peak_source": "measured (MEASURED_PEAKS.json bf16_tflops_sustained; kind::f16 MMAs run at the bf16 rate)"
                        if peaks else "fallback", "unit": "TFLOP/s", "frac": achieved / tc_peak, "traffic": traffic,
                        "mma_flops_per_launch": g["flops"] / g["n"], "ms_per_launch": g["ms"] / g["n"],
                        "useful_tflops": g["useful_flops"] / g["ms"] / 1e9,
                        "hbm": {"algorithmic_gbs": g["bytes"] / g["ms"] / 1e6, "plane_stream_gbs": g["plane_bytes"] / g["ms"] / 1e6,
                                "peak_gbs": hbm_peak, "algorithmic_frac": g["bytes"] / g["ms"] / 1e6 / hbm_peak,
                                "plane_stream_frac": g["plane_bytes"] / g["ms"] / 1e6 / hbm_peak,
                                "algorithmic_bytes_per_launch": g["bytes"] / g["n"]},
                        "note": "achieved = MMA flops the kernel issues (fp16 hi*hi + lo*hi + hi*lo over the zero-filled block) / its "
                                "CUDA-event time; useful_tflops = 2*D per expression entry; hbm.algorithmic = SURVEY 8d bytes of the "
                                "same entries as a CSR, hbm.plane_stream = the 16-bit planes the kernel actually streams"}
        else:
            achieved = g["bytes"] / g["ms"] / 1e6
            roofline = {"kernel": f"agg_{key[0]} ({key[1]})", "bound": "hbm", "achieved": achieved, "peak": hbm_peak,
                        "peak_source": peak_src, "unit": "GB/s", "frac": achieved / hbm_peak, "traffic": traffic,
                        "algorithmic_bytes_per_launch": g["bytes"] / g["n"], "ms_per_launch": g["ms"] / g["n"],
                        "gather_side_tbs": g["nnz"] * a.dim * 4 / g["ms"] / 1e9,
                        "note": "algorithmic bytes count distinct rows once (SURVEY 8d); the E*D gather is served "
                                "on-chip (L2/L1/smem), reported as gather_side_tbs"}

    # ---- end-to-end through the public API with HOST buffers: `e2e` --------------------------
    e2e = None
    if not a.no_e2e:
        hfeat = feats.cpu().pin_memory()
        hlab = labels.cpu().pin_memory()
        trainer.step(hfeat, hlab)
        barrier()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            trainer.step(hfeat, hlab)          # H2D features+labels, fwd+bwd+Adam, D2H loss
        torch.cuda.synchronize()
        e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / a.steps)
        h2d = hfeat.numel() * 4 + hlab.numel() * 8
        e2e = {"value": a.cells / (e2e_ms / 1e3), "unit": "cells/s", "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": 4 * world,
               "api": "FullGraphTrainer.step(features_host_pinned, labels_host_pinned) -> float loss"}

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        n = min(a.cpu_sample_cells, a.cells)
        sec = cpu_reference_steps(a, n, 1, 1)
        cpu = {"value": n / sec, "unit": "cells/s", "cores": os.cpu_count(), "kind": "port",
               "sample": f"first {n} cells of the same atlas, 1 warm-up + 1 timed full-graph fwd+bwd+Adam step, "
                         f"oracle port (edge-materialising, as models/gnn.py:54-56), torch CPU fp32"}
        n2 = min(8 * n, a.cells)
        sec2 = cpu_spmm_steps(a, n2, 1, 1)
        cpu["optimised"] = {"value": n2 / sec2, "unit": "cells/s", "cores": os.cpu_count(), "kind": "port, closed form",
                            "sample": f"first {n2} cells, same step without the per-edge message tensor: torch sparse-CSR x "
                                      f"dense products (oracle/spmm_oracle.py), 1 warm-up + 1 timed step"}

    if trainer.peer_group is not None:
        trainer.peer_group.check()              # a timed-out cross-GPU barrier is an error, not a slow step
    if rank == 0:
        print(json.dumps({
            "metric": "cells/sec (forward+backward)", "value": a.cells / (ms_step / 1e3), "unit": "cells/s",
            "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32" if a.dense_fmt == "f16x2" else "bf16", "data": "synthetic",
            "config": {"workload": workload_name(a), "cells": a.cells, "genes": a.genes, "nnz_per_rank": graph.nnz,
                       "dense_threshold": a.dense_threshold, "dense_fmt": a.dense_fmt, "dense_genes": int(len(getattr(graph, "dense_genes", []))),
                       "csr_entries_left": graph.cell_csr.nnz, "dropout": a.dropout,
                       "parallelism": f"cell-sharded x{world}" if world > 1 else "single GPU",
                       "gene_sum_exchange": None if world == 1 else "wsage_peer_reduce (slab sum + all-reduce + epilogue, one kernel over "
                       "NVLink peer memory)" if trainer.peer_group is not None else "sum_slabs + NCCL all-reduce + element-wise",
                       "l2_policy": "inputs_exceed_l2 (graph + activations >> 126 MB; no explicit flush)",
                       "graph_build_s": build_s, "final_loss_per_cell": final_loss / (hi - lo)},
            "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": launches, "clocks": clocks,
        }))
    if world > 1:
        if trainer.peer_group is not None:
            from scdeepsort_b200 import peer
            peer.disable()
        dist.destroy_process_group()


def run_sampled(a):
    """BASELINE configs[4]: neighbour-sampled mini-batch training (graph replicated per GPU, seeds sharded,
    gradients all-reduced).  Not the driver's default line; `python bench.py --mode sampled --layers 3 --hidden 800`."""
    import torch.distributed as dist
    import scdeepsort_b200 as sd
    from scdeepsort_b200 import parallel
    from scdeepsort_b200.synthetic import synthetic_bipartite, synthetic_features

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    fanouts = [int(f) for f in a.fanouts.split(",")]
    if len(fanouts) != a.layers:
        raise SystemExit("--fanouts needs one entry per layer")
    t0 = time.time()
    bg = synthetic_bipartite(a.cells, a.genes, a.deg, seed=SEED, device=dev)
    feats = synthetic_features(bg, a.dim, seed=SEED)
    graph = sd.DeepSortGraph.from_bipartite(bg, feats)
    del bg
    torch.cuda.empty_cache()
    labels = torch.cat([torch.full((a.genes,), -1, dtype=torch.int64),
                        torch.randint(0, NUM_CLASSES, (a.cells,), generator=torch.Generator().manual_seed(SEED))]).to(dev)
    torch.cuda.synchronize()
    build_s = time.time() - t0
    torch.manual_seed(SEED)
    model = sd.GNN(a.dim, a.hidden, NUM_CLASSES, a.layers, a.genes, activation=torch.relu, dropout=0.0).to(dev)
    parallel.broadcast_params(model)
    opt = sd.optim.Adam(model.parameters(), lr=1e-3, weight_decay=5e-4)
    lo, hi = parallel.cell_ranges(a.cells, world)[rank]
    seeds = torch.arange(a.genes + lo, a.genes + hi, device=dev)
    sampler = sd.NeighborSampler(graph, a.batch, num_hops=a.layers, neighbor_type='in', shuffle=True, seed_nodes=seeds,
                                 fanouts=fanouts, seed=SEED + rank, generator=torch.Generator(device=dev).manual_seed(SEED + rank))
    it = iter(sampler)
    sizes = []

    def step():
        nf = next(it)
        nf.copy_from_parent()
        logits = model(nf)
        loss = sd.optim.cross_entropy_sum(logits, labels[nf.layer_parent_nid(-1)])
        opt.zero_grad(set_to_none=True)
        loss.backward()
        parallel.allreduce_grads(model)
        opt.step()
        sizes.append([nf.layer_size(i) for i in range(nf.num_layers)] + [nf.block_size(i) for i in range(nf.num_blocks)])
        return loss

    from scdeepsort_b200 import ops
    for _ in range(a.warmup):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    csamp = ClockSampler(local_rank) if rank == 0 else None
    if csamp:
        csamp.start()
    sd._lib.launch_count(reset=True)
    ops.TIMING = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    timing, ops.TIMING = ops.TIMING, None
    launches = sd._lib.launch_count()
    ms = e0.elapsed_time(e1) / a.steps
    # ---- end to end: seed ids and labels come from HOST memory every step, the loss goes back (train.py:79-87) ----
    h_seeds = seeds.cpu().pin_memory()
    h_lab = labels.cpu().pin_memory()
    perm = torch.randperm(h_seeds.shape[0], generator=torch.Generator().manual_seed(SEED))

    def e2e_step(i):
        idx = perm[(i * a.batch) % max(1, h_seeds.shape[0] - a.batch):][:a.batch]
        batch = h_seeds[idx].pin_memory().to(dev, non_blocking=True)                 # H2D: this step's seed ids
        lab = h_lab[h_seeds[idx]].pin_memory().to(dev, non_blocking=True)            # H2D: their labels
        nf = sampler.build(batch)
        nf.copy_from_parent()
        loss = sd.optim.cross_entropy_sum(model(nf), lab)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        parallel.allreduce_grads(model)
        opt.step()
        return float(loss.detach())                                                           # D2H: loss.item()

    e2e_step(0)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(a.steps):
        e2e_step(i + 1)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / a.steps
    clocks = csamp.stop() if csamp else None
    if world > 1:
        t = torch.tensor([ms, e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms = float(t[0]), float(t[1])
    # ---- roofline: the gather kernels of the sampled blocks are the HBM-bound regime (rows come from a 1.2 GB table) ----
    peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text()) if (ROOT / "MEASURED_PEAKS.json").exists() else {}
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    groups = {}
    for t in timing:
        if t["kind"] not in ("block_fwd", "block_bwd"):
            continue
        g = groups.setdefault(t["kind"], dict(ms=0.0, n=0, bytes=0))
        g["ms"] += t["events"][0].elapsed_time(t["events"][1])
        g["n"] += 1
        rows = t["n_src"] * t["dim"] * 4
        if t["kind"] == "block_fwd":        # edges (int32 + fp32) + rowptr + distinct source rows once + destination rows written
            g["bytes"] += t["edges"] * 8 + (t["n_dst"] + 1) * 8 + rows + t["n_dst"] * t["dim"] * 4
        else:                                # edges + dOut rows read + dH rows accumulated (read + write) + H rows read for d-alpha
            g["bytes"] += t["edges"] * 8 + (t["n_dst"] + 1) * 8 + t["n_dst"] * t["dim"] * 4 + (2 * rows if t["need_h"] else 0) + (rows if t["need_a"] else 0)
    kernels = {k: {"launches": g["n"], "ms_per_launch": g["ms"] / g["n"], "share_of_step": g["ms"] / a.steps / ms,
                   "algorithmic_gbs": g["bytes"] / g["ms"] / 1e6, "frac_of_hbm_peak": g["bytes"] / g["ms"] / 1e6 / hbm_peak}
               for k, g in groups.items()}
    roofline = None
    if groups:
        key, g = max(groups.items(), key=lambda kv: kv[1]["ms"])
        ach = g["bytes"] / g["ms"] / 1e6
        roofline = {"kernel": "agg_gather_fwd_kernel" if key == "block_fwd" else "agg_gather_bwd_kernel", "bound": "hbm", "achieved": ach,
                    "peak": hbm_peak, "peak_source": "measured" if peaks else "fallback", "unit": "GB/s", "frac": ach / hbm_peak,
                    "traffic": None, "algorithmic_bytes_per_launch": g["bytes"] / g["n"], "ms_per_launch": g["ms"] / g["n"],
                    "note": "sum over the blocks of a step (the outermost block dominates); the step itself is launch-bound: "
                            "see share_of_step in kernels"}
    # ---- CPU baseline: the oracle's literal restatement of models/gnn.py on the SAME sampled blocks, host cores ----
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        from oracle import gnn_oracle
        from oracle.gnn_oracle import OracleBlock, OracleFlow
        torch.set_num_threads(os.cpu_count())
        nb = max(8, min(64, a.batch))
        nf = sampler.build(seeds[:nb])
        nf.copy_from_parent()
        blocks = []
        for b in nf.blocks:
            deg = (b.rowptr[1:] - b.rowptr[:-1]).cpu()
            blocks.append(OracleBlock(b.col.cpu().long(), torch.repeat_interleave(torch.arange(b.n_dst), deg), b.weight.cpu(), b.n_src, b.n_dst))
        flow = OracleFlow([nf.layer_parent_nid(i).cpu() for i in range(nf.num_layers)],
                          [nf.layers[i].data["id"].reshape(-1).cpu() for i in range(nf.num_layers)],
                          nf.layers[0].data["features"].cpu(), blocks)
        params = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in model.state_dict().items()}
        copt = torch.optim.Adam(list(params.values()), lr=1e-3, weight_decay=5e-4)
        lab = labels[nf.layer_parent_nid(-1)].cpu()
        ts = []
        for i in range(3):
            t0 = time.perf_counter()
            closs = torch.nn.functional.cross_entropy(gnn_oracle.forward(params, flow, a.genes), lab, reduction="sum")
            copt.zero_grad()
            closs.backward()
            copt.step()
            ts.append(time.perf_counter() - t0)
        cpu = {"value": nb / min(ts[1:]), "unit": "cells/s", "cores": os.cpu_count(), "kind": "port",
               "sample": f"one sampled mini-batch of {nb} seed cells (same sampler, same fan-outs, blocks copied to the host), fwd+bwd+Adam with "
                         f"the oracle port of models/gnn.py (edge-materialising), best of 2 after 1 warm-up; sampling itself not timed"}
    if rank == 0:
        print(json.dumps({
            "metric": "cells/sec (forward+backward), neighbour-sampled mini-batches", "value": world * a.batch / (ms / 1e3),
            "unit": "cells/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"synthetic {a.cells} cells x {a.genes} genes, avg-degree {int(a.deg)}, {a.layers}-layer "
                                   f"hidden={a.hidden}, fan-outs {fanouts}, {a.batch} seed cells per GPU per step, graph replicated",
                       "graph_build_s": build_s, "graph_edges": graph.number_of_edges(),
                       "l2_policy": "inputs_exceed_l2 (rows are gathered from a 1.2 GB feature table; no explicit flush)",
                       "last_flow_layers_then_blocks": sizes[-1], "final_loss_per_cell": float(loss) / a.batch},
            "gpu_launches": launches, "roofline": roofline, "kernels": kernels, "cpu_baseline": cpu,
            "e2e": {"value": world * a.batch / (e2e_ms / 1e3), "unit": "cells/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": world * a.batch * 16, "d2h_bytes_per_step": world * 4,
                    "api": "NeighborSampler.build(seed ids from host) -> GNN(nf) -> cross_entropy_sum -> backward -> Adam.step -> float(loss)"},
            "clocks": clocks}))
    if world > 1:
        dist.destroy_process_group()


def _json_only_stdout():
    """Everything the libraries print to fd 1 (e.g. c10d's "NCCL version ..." banner) goes to stderr; the
    returned function writes the ONE JSON line to the real stdout."""
    sys.stdout.flush()
    real = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.write(real, (line + "\n").encode())
    return emit


if __name__ == "__main__":
    args = parse()
    _emit = _json_only_stdout()
    import builtins
    _print = builtins.print
    builtins.print = lambda *a, **k: _emit(" ".join(str(x) for x in a)) if not k.get("file") else _print(*a, **k)
    if args.impl == "reference":
        run_reference(args)
    elif args.mode == "sampled":
        run_sampled(args)
    else:
        run_ours(args)
