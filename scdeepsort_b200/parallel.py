"""Cell-sharded data parallelism for the full-graph path.  One process per GPU.

Cells (rows of X, their CSR edges, feature rows, labels) are partitioned contiguously over
ranks; gene-side state (gene features/activations, α, all weights) is replicated.  The only
data-path exchange is where genes aggregate over ALL cells:

  forward   S_g = Σ_ranks Σ_{c in shard} x_cg·h_c         → all-reduce(sum) of [G, D] per gene layer
  backward  of that all-reduce is an all-reduce of the incoming gradient — only reached when the
            cell features feeding it require grad (layers ≥ 2)
  grads     loss is a SUM over cells (train.py:36), and back-propagation is linear in the
            incoming gradient, so each rank back-propagates its own cells' partial gradient
            through the replicated gene-side ops and ONE all-reduce(sum) of the flattened
            parameter gradients gives the exact full-batch gradient.

The reference has no distributed code at all (no NCCL/Gloo call sites); this is new surface.
Communication helpers here are backend-agnostic (NCCL on GPUs, gloo in the CPU tests).
"""
from typing import List, Tuple

import torch
import torch.distributed as dist

from .gnn import GNN
from .graph import BipartiteGraph
from .nodeflow import FullGraphFlow


def is_dist():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def cell_ranges(num_cells: int, world_size: int) -> List[Tuple[int, int]]:
    """Contiguous, balanced partition of cell indices: rank r owns [lo, hi)."""
    base, rem = divmod(num_cells, world_size)
    out, lo = [], 0
    for r in range(world_size):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


def globalize_gene_normalisers(graph: BipartiteGraph):
    """deg_g and Σ_c x_cg are sums over ALL cells: all-reduce the shard-local partials once and
    rebuild ``norm_g`` / ``mean_g`` (utils/preprocess_internal.py:15-23 applied to the whole atlas)."""
    deg = graph.local_deg_g.to(torch.float64).clone()
    colsum = graph.local_colsum_g.to(torch.float64).clone()
    if is_dist():
        dist.all_reduce(deg)
        dist.all_reduce(colsum)
    graph.norm_g = torch.where(deg > 0, deg / colsum.clamp(min=1e-30), torch.zeros_like(deg)).float()
    graph.mean_g = (1.0 / (deg + 1)).float()
    return graph


class AllReduceSum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        y = x.contiguous().clone()
        if is_dist():
            dist.all_reduce(y)
        return y

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous().clone()
        if is_dist():
            dist.all_reduce(g)
        return g


def sharded_forward(model: GNN, graph: BipartiteGraph, features: torch.Tensor, cells_ready=None) -> torch.Tensor:
    """GNN.forward on this rank's shard: features = cat[gene rows (replicated); local cell rows].  The layer loop is
    the model's own (``GNN._forward_full``); the flow's ``sharded`` flag makes every gene-producing layer all-reduce
    its raw gene sums and draws the replicated gene rows' dropout mask from a generator seeded alike on every rank.
    ``cells_ready``: event after which the cell rows of ``features`` are valid (see FullGraphFlow)."""
    flow = FullGraphFlow(graph, features, cells_ready=cells_ready)
    flow.sharded = True
    return model._forward_full(flow)


def enable_peer_exchange(graph: BipartiteGraph, max_dim: int):
    """Collective.  Routes the all-reduce of the raw gene sums of ``graph`` through ``wsage_peer_reduce`` (one kernel over
    NVLink peer memory, scdeepsort_b200/peer.py) when every rank runs its gene passes entirely on the dense block and the
    GPUs can map each other's memory; otherwise the NCCL all-reduce stays.  Sets and returns ``graph.peer_group``."""
    from . import peer
    graph.peer_group = None
    if not is_dist() or graph.device.type != "cuda":
        return None
    mine = graph.gene_csr.dense is not None and graph.gene_csr.dense_side == 1 and graph.gene_csr.nnz == 0
    flag = torch.tensor([1 if mine else 0], device=graph.device, dtype=torch.int32)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if int(flag) == 1:
        graph.peer_group = peer.enable(graph.num_genes * int(max_dim))
    return graph.peer_group


def allreduce_grads(model: torch.nn.Module):
    """One flattened all-reduce(sum) of every parameter gradient (≈1.4 MB at H=400)."""
    if not is_dist():
        return
    grads = [p.grad for p in model.parameters() if p.grad is not None]
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n


def broadcast_params(model: torch.nn.Module, src=0):
    if not is_dist():
        return
    for p in model.parameters():
        dist.broadcast(p.data, src)
