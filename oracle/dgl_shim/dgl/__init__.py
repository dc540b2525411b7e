"""Minimal stand-in for ``dgl==0.4.3.post2`` (TEST INFRASTRUCTURE ONLY).

The reference pins DGL 0.4.3 (``/root/reference/requirements.txt:5``), which has no
wheel for Python 3.12 / torch 2.11 and cannot be installed here (no network).  This
package restates, in plain CPU PyTorch, exactly the slice of the DGL 0.4 API that the
reference touches, so the *unmodified* reference sources (``models/gnn.py``,
``utils/preprocess_internal.py``, ``utils/preprocess.py``) can be imported and executed
in this container to mint golden vectors (``oracle/gen_golden.py``).

API surface restated (call sites in the reference):
  * ``DGLGraph()``: ``add_nodes`` / ``add_edges`` / ``ndata`` / ``edata`` / ``in_degrees`` /
    ``in_edges(v, form='all')`` / ``number_of_nodes`` / ``number_of_edges`` / ``nodes`` /
    ``readonly``  (``utils/preprocess_internal.py:15-23,107-110,168-173,202,211-215``)
  * ``dgl.function.mean``                               (``models/gnn.py:65``)
  * ``dgl.contrib.sampling.NeighborSampler`` → ``NodeFlow`` with ``layers[i].data``,
    ``copy_from_parent``, ``block_compute``, ``layer_parent_nid``
    (``train.py:71-81``, ``predict.py:64-75``, ``models/gnn.py:58-66``)
  * ``EdgeBatch.src / .dst / .data`` and ``NodeBatch.data``   (``models/gnn.py:18-25,47-56``)

Semantics follow DGL 0.4's published behaviour: edge ids are insertion order; ``fn.mean``
is sum over in-edges of the block divided by the in-degree *inside the block*; a
NodeFlow's layer ``i+1`` nodes receive from layer ``i`` through block ``i``; full-neighbour
sampling (``expand_factor`` ≥ in-degree) keeps every in-edge; sampling is uniform without
replacement otherwise.  Nothing under ``scdeepsort_b200/`` imports this.
"""
from . import function  # noqa: F401
from .graph import DGLGraph, EdgeBatch, NodeBatch  # noqa: F401
from .nodeflow import NodeFlow  # noqa: F401
from . import contrib  # noqa: F401

__version__ = "0.4.3.post2+shim"
