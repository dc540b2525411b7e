"""CPU oracle for scDeepSort's weighted-GraphSAGE hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl
reference`` legs may import this module, and only as the checker or the timed CPU
baseline — never from the product path under ``scdeepsort_b200/``.

This is a literal restatement (plain CPU PyTorch, fp32 or fp64) of
  * ``GNN.message_func``      /root/reference/models/gnn.py:47-56   (α-index cascade + h·α·w)
  * ``fn.mean('m','neigh')``  /root/reference/models/gnn.py:65      (Σ in-edges ÷ in-degree in block)
  * ``NodeUpdate.forward``    /root/reference/models/gnn.py:18-25   (Linear → activation)
  * ``GNN.forward``           /root/reference/models/gnn.py:58-68   (layer loop, dropout placement, classifier)
  * loss                      /root/reference/train.py:36,82        (CrossEntropyLoss(reduction='sum'))
  * unsure rule               /root/reference/train.py:108-113, predict.py:79-87
The per-edge message tensor [E, D] IS materialised, as the reference does.

Parity pin: the reference has no tests/golden vectors of its own, and
``dgl==0.4.3.post2`` cannot be installed here.  The pin is therefore "outputs of the
reference itself run here": ``oracle/gen_golden.py`` imports the UNMODIFIED
``/root/reference/models/gnn.py`` (and ``utils/preprocess_internal.py``) on top of the DGL
API shim in ``oracle/dgl_shim`` and stores its logits / gradients under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this oracle against those files.
"""
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np
import torch
import torch.nn.functional as F


@dataclass
class OracleBlock:
    """Edges from layer i (src, local index) to layer i+1 (dst, local index)."""
    src: torch.Tensor      # int64 [E]
    dst: torch.Tensor      # int64 [E]
    weight: torch.Tensor   # float [E]   edata['weight'] (normalised; self-loop = 1)
    n_src: int
    n_dst: int


@dataclass
class OracleFlow:
    """What the reference's NodeFlow carries after ``copy_from_parent`` (train.py:79)."""
    layer_nid: List[torch.Tensor]      # parent node ids per layer, int64
    layer_id: List[torch.Tensor]       # ndata['id'] per layer: gene → gene index, cell → -1
    features: torch.Tensor             # ndata['features'] of layer 0, [N0, D0]
    blocks: List[OracleBlock] = field(default_factory=list)


def alpha_index(src_id: np.ndarray, dst_id: np.ndarray, gene_num: int) -> np.ndarray:
    """The host ``np.where`` cascade of models/gnn.py:49-53, verbatim semantics."""
    n = src_id.shape[0]
    idx = np.full(n, gene_num + 1, dtype=np.int64)                        # default: cell-cell self loop
    idx = np.where((src_id >= 0) & (dst_id < 0), src_id, idx)             # gene -> cell
    idx = np.where((dst_id >= 0) & (src_id < 0), dst_id, idx)             # cell -> gene
    idx = np.where((dst_id >= 0) & (src_id >= 0), gene_num, idx)          # gene - gene
    return idx


def block_aggregate(h, alpha, block: OracleBlock, src_id, dst_id, gene_num):
    """message (gnn.py:54-56) then mean (gnn.py:65).  h [n_src, D]; alpha [G+2, 1]."""
    idx = alpha_index(src_id[block.src].numpy(), dst_id[block.dst].numpy(), gene_num)
    m = h[block.src] * alpha[torch.from_numpy(idx)]          # [E, D] * [E, 1]
    m = m * block.weight.to(h.dtype).unsqueeze(-1)           # [E, D]
    neigh = torch.zeros(block.n_dst, h.shape[1], dtype=h.dtype).index_add(0, block.dst, m)
    deg = torch.bincount(block.dst, minlength=block.n_dst).clamp(min=1).to(h.dtype)
    return neigh / deg.unsqueeze(-1)


def forward(params: dict, flow: OracleFlow, gene_num: int, dtype=torch.float32,
            dropout_masks: Optional[List[torch.Tensor]] = None, activation=F.relu,
            relu_masks: Optional[List[torch.Tensor]] = None, hidden_out: Optional[list] = None):
    """GNN.forward (gnn.py:58-68).  ``params`` uses the reference's state_dict keys.

    dropout_masks[i], if given, is the already-scaled inverted-dropout multiplier applied to
    layer i's node features before aggregation (gnn.py:62-64); None = eval mode.
    relu_masks[i], if given, replaces layer i's ReLU by a multiplication with that 0/1 mask: gradients are
    discontinuous where a pre-activation crosses zero, so a gradient comparison between two implementations is only
    meaningful under the same kink decisions (the tests take the masks from the implementation under test and check
    separately that they differ from the oracle's own in a negligible share of entries).
    """
    n_layers = len(flow.blocks)
    alpha = params["alpha"].to(dtype)
    h = flow.features.to(dtype)
    for i in range(n_layers):
        if dropout_masks is not None:
            h = h * dropout_masks[i].to(dtype)
        neigh = block_aggregate(h, alpha, flow.blocks[i], flow.layer_id[i], flow.layer_id[i + 1], gene_num)
        w = params[f"layers.{i}.fc_neigh.weight"].to(dtype)
        b = params[f"layers.{i}.fc_neigh.bias"].to(dtype)
        h = neigh @ w.t() + b
        if relu_masks is not None:
            h = h * relu_masks[i].to(dtype)
        elif activation is not None:
            h = activation(h)
        if hidden_out is not None:
            hidden_out.append(h.detach())
    return h @ params["linear.weight"].to(dtype).t() + params["linear.bias"].to(dtype)


def loss_and_grads(params: dict, flow: OracleFlow, labels: torch.Tensor, gene_num: int,
                   dtype=torch.float32, feature_grad=False):
    """CE(sum) (train.py:36,82) and autograd gradients w.r.t. every parameter."""
    p = {k: v.detach().clone().to(dtype).requires_grad_(True) for k, v in params.items()}
    feats = flow.features.detach().clone().to(dtype).requires_grad_(feature_grad)
    fl = OracleFlow(flow.layer_nid, flow.layer_id, feats, flow.blocks)
    logits = forward(p, fl, gene_num, dtype)
    loss = F.cross_entropy(logits, labels, reduction="sum")
    loss.backward()
    grads = {k: v.grad.detach() for k, v in p.items()}
    if feature_grad:
        grads["features"] = feats.grad.detach()
    return loss.detach(), logits.detach(), grads


def predict_labels(logits: torch.Tensor, unsure_rate: float):
    """softmax → argmax, 'unsure' (-1) when max prob < unsure_rate / K (predict.py:79-87)."""
    prob = F.softmax(logits.float(), dim=1)
    max_prob, arg = prob.max(dim=1)
    k = logits.shape[1]
    return torch.where(max_prob < unsure_rate / k, torch.full_like(arg, -1), arg)


def init_params(in_feats, n_hidden, n_classes, n_layers, gene_num, seed=10086, perturb_alpha=False):
    """Reference initialisation (gnn.py:13,16,42-45): default nn.Linear init with the weight
    overwritten by xavier_uniform(gain=√2); α = 1.  ``perturb_alpha`` draws α ~ U(0.5, 1.5) so
    α-indexing bugs are visible in parity runs."""
    g = torch.Generator().manual_seed(seed)
    params = {}
    gain = torch.nn.init.calculate_gain("relu")

    def linear(prefix, fin, fout):
        bound_w = gain * (6.0 / (fin + fout)) ** 0.5
        params[prefix + ".weight"] = (torch.rand(fout, fin, generator=g) * 2 - 1) * bound_w
        bound_b = 1.0 / fin ** 0.5
        params[prefix + ".bias"] = (torch.rand(fout, generator=g) * 2 - 1) * bound_b

    dims = [in_feats] + [n_hidden] * n_layers
    for i in range(n_layers):
        linear(f"layers.{i}.fc_neigh", dims[i], dims[i + 1])
    linear("linear", n_hidden, n_classes)
    alpha = torch.ones(gene_num + 2, 1)
    if perturb_alpha:
        alpha = 0.5 + torch.rand(gene_num + 2, 1, generator=g)
    params["alpha"] = alpha
    return params
