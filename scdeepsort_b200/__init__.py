"""scdeepsort_b200 — B200-native weighted-GraphSAGE hot path of scDeepSort.

Host side in Python/PyTorch (mirroring /root/reference/models/gnn.py and the DGL NodeFlow
surface), arithmetic in hand-written sm_100a CUDA behind the C ABI of ``include/wsage.h``.
"""
from . import _lib, dense, optim, parallel
from .gnn import GNN, NodeUpdate, predict_labels
from .graph import BipartiteGraph, DeepSortGraph
from .nodeflow import FullGraphFlow, NeighborSampler, NodeFlow
from .ops import Block, Csr, block_aggregate, spmm

__all__ = ["GNN", "NodeUpdate", "predict_labels", "BipartiteGraph", "DeepSortGraph", "FullGraphFlow",
           "NeighborSampler", "NodeFlow", "Block", "Csr", "block_aggregate", "spmm", "_lib"]
