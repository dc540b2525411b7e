"""Inner loops of the reference drivers, rebuilt on the CUDA hot path.

* ``Trainer``      mirrors /root/reference/train.py:16-123 (``train`` / ``evaluate`` per 500-seed
                   NodeFlow, CE(sum), Adam with weight decay on every parameter, best-checkpoint
                   save with the reference's ``{'model', 'optimizer'}`` layout).
* ``Runner``       mirrors /root/reference/predict.py:61-88 (batched full-neighbour inference,
                   softmax, 'unsure' rule) — post-processing vectorised on device.
* ``FullGraphTrainer`` is the throughput form: one layer-wise pass over the whole bipartite
                   graph per step (every cell is a seed), optionally cell-sharded over ranks.

Graph construction from files stays with the caller: these classes take a
``DeepSortGraph`` / ``BipartiteGraph``.
"""
from pathlib import Path
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .gnn import GNN, predict_labels
from .graph import BipartiteGraph, DeepSortGraph
from .nodeflow import FullGraphFlow, NeighborSampler
from .optim import Adam, cross_entropy_sum


class Trainer:
    def __init__(self, graph: DeepSortGraph, labels: torch.Tensor, train_ids, test_ids, num_labels: int, *,
                 dense_dim=400, hidden_dim=200, n_layers=1, dropout=0.1, lr=1e-3, weight_decay=5e-4,
                 batch_size=500, num_neighbors=0, unsure_rate=2.0, device="cuda:0", save_path=None):
        self.device = torch.device(device)
        self.graph = graph.to(self.device)
        self.labels = labels.to(self.device)
        self.train_ids = torch.as_tensor(train_ids, dtype=torch.int64)
        self.test_ids = torch.as_tensor(test_ids, dtype=torch.int64)
        self.num_labels, self.n_layers, self.batch_size, self.unsure_rate = num_labels, n_layers, batch_size, unsure_rate
        self.model = GNN(in_feats=dense_dim, n_hidden=hidden_dim, n_classes=num_labels, n_layers=n_layers,
                         gene_num=graph.num_genes, activation=F.relu, dropout=dropout).to(self.device)
        self.optimizer = Adam(self.model.parameters(), lr=lr, weight_decay=weight_decay)     # train.py:34-35
        self.loss_fn = cross_entropy_sum                                                     # train.py:36
        n = graph.number_of_nodes()
        self.num_neighbors = n if num_neighbors == 0 else num_neighbors        # train.py:37-40
        self.save_path = Path(save_path) if save_path else None
        # the reference draws fresh neighbour samples every epoch from the process-wide RNG: the device sampler is keyed
        # by (seed, batch, hop, node), so every epoch gets its own seed, derived from torch's seed (random_seed)
        self._sample_seed = int(torch.initial_seed()) & 0x7FFFFFFF
        self._epoch = 0

    def train(self):
        """train.py:68-89."""
        self.model.train()
        total = torch.zeros((), device=self.device)
        self._epoch += 1
        for nf in NeighborSampler(g=self.graph, batch_size=self.batch_size, expand_factor=self.num_neighbors,
                                  num_hops=self.n_layers, neighbor_type='in', shuffle=True, num_workers=8,
                                  seed_nodes=self.train_ids, seed=self._sample_seed + 7919 * self._epoch):
            nf.copy_from_parent()
            logits = self.model(nf)
            batch_nids = nf.layer_parent_nid(-1)
            loss = self.loss_fn(logits, self.labels[batch_nids])
            self.optimizer.zero_grad()
            loss.backward()
            self.optimizer.step()
            total += loss.detach()          # no per-batch D2H sync (the reference's loss.item(), train.py:87)
        return float(total)

    @torch.no_grad()
    def evaluate(self, ids):
        """train.py:91-115, softmax/argmax/unsure vectorised on device."""
        self.model.eval()
        correct = unsure = 0
        for nf in NeighborSampler(g=self.graph, batch_size=self.batch_size, expand_factor=self.graph.number_of_nodes(),
                                  num_hops=self.n_layers, neighbor_type='in', shuffle=True, num_workers=8, seed_nodes=ids):
            nf.copy_from_parent()
            pred = predict_labels(self.model(nf), self.unsure_rate)
            lab = self.labels[nf.layer_parent_nid(-1)]
            unsure += int((pred < 0).sum())
            correct += int(((pred == lab) & (pred >= 0)).sum())
        return correct, unsure

    def fit(self, n_epochs=300, verbose=True):
        """train.py:44-66."""
        best = dict(test_acc=0.0, epoch=0, train_acc=0.0)
        for epoch in range(n_epochs):
            loss = self.train()
            train_correct, _ = self.evaluate(self.train_ids)
            train_acc = train_correct / len(self.train_ids)
            test_correct, test_unsure = self.evaluate(self.test_ids)
            test_acc = test_correct / max(1, len(self.test_ids))
            if best["test_acc"] <= test_acc:
                best = dict(test_acc=test_acc, epoch=epoch, train_acc=train_acc, test_correct=test_correct,
                            test_unsure=test_unsure)
                self.save_model()
            if verbose:
                print(f">>>>Epoch {epoch:04d}: Train Acc {train_acc:.4f}, Loss {loss / len(self.train_ids):.4f}, "
                      f"Test correct {test_correct}, Test unsure {test_unsure}, Test Acc {test_acc:.4f}")
            if train_acc == 1:
                break
        return best

    def save_model(self):
        """train.py:117-123: same dict layout, so reference tooling can read it."""
        if self.save_path is None:
            return
        self.save_path.parent.mkdir(parents=True, exist_ok=True)
        torch.save({'model': self.model.state_dict(), 'optimizer': self.optimizer.state_dict()}, self.save_path)


class Runner:
    """predict.py:17-88 for one test graph: load ``state['model']``, batched inference, labels."""

    def __init__(self, graph: DeepSortGraph, test_nid, num_classes: int, *, dense_dim=400, hidden_dim=200,
                 n_layers=1, batch_size=500, unsure_rate=2.0, device="cuda:0", model_path=None, state_dict=None,
                 bipartite: Optional[BipartiteGraph] = None):
        self.device = torch.device(device)
        self.graph = graph.to(self.device)
        # the same graph factored for the full-graph form: every test cell is a seed of ONE layer-wise pass instead of
        # 500-seed NodeFlows that each re-derive the shared lower layers (same logits: tests/test_gpu_c2_parity.py)
        self.bipartite = bipartite
        self.test_nid = torch.as_tensor(test_nid, dtype=torch.int64)
        self.batch_size, self.unsure_rate, self.n_layers = batch_size, unsure_rate, n_layers
        self.model = GNN(in_feats=dense_dim, n_hidden=hidden_dim, n_classes=num_classes, n_layers=n_layers,
                         gene_num=graph.num_genes, activation=F.relu, dropout=0.1)
        if model_path is not None:
            state_dict = torch.load(model_path, map_location="cpu")['model']        # predict.py:58-59
        if state_dict is not None:
            self.model.load_state_dict(state_dict)
        self.model.to(self.device)

    @torch.no_grad()
    def inference(self):
        """Returns (pred, logits): pred[i] = class index or -1 ('unsure') for test cell i."""
        self.model.eval()
        n = self.graph.number_of_nodes()
        if self.bipartite is not None:
            cells = (self.test_nid - self.graph.num_genes).to(self.device)
            logits = self.model(FullGraphFlow(self.bipartite, self.graph.ndata["features"], seeds=cells))
            return predict_labels(logits, self.unsure_rate), logits
        new_logits = torch.zeros(n, self.model.linear.out_features, device=self.device)
        for nf in NeighborSampler(g=self.graph, batch_size=self.batch_size, expand_factor=n, num_hops=self.n_layers,
                                  neighbor_type='in', shuffle=False, num_workers=8, seed_nodes=self.test_nid):
            nf.copy_from_parent()
            new_logits[nf.layer_parent_nid(-1)] = self.model(nf)
        logits = new_logits[self.test_nid.to(self.device)]
        return predict_labels(logits, self.unsure_rate), logits


class FullGraphTrainer:
    """One optimisation step per pass over the whole (local shard of the) atlas.

    ``step(features, labels)`` takes HOST or device tensors: host tensors are staged through pinned
    memory and copied H2D inside the call (that copy is what ``copy_from_parent`` does for a
    CPU-resident DGL graph, train.py:79) and the scalar loss comes back D2H, like
    ``loss.item()`` (train.py:87).  With ``torch.distributed`` initialised and ``sharded=True`` the
    graph is this rank's cell shard; gene partial sums and gradients are all-reduced."""

    def __init__(self, graph: BipartiteGraph, num_labels: int, *, dense_dim=400, hidden_dim=400, n_layers=2,
                 dropout=0.0, lr=1e-3, weight_decay=5e-4, seed=10086, sharded=False, spmm_algo=0):
        self.graph = graph
        self.device = graph.device
        torch.manual_seed(seed)
        self.model = GNN(in_feats=dense_dim, n_hidden=hidden_dim, n_classes=num_labels, n_layers=n_layers,
                         gene_num=graph.num_genes, activation=F.relu, dropout=dropout).to(self.device)
        self.model.spmm_algo = spmm_algo
        self.optimizer = Adam(self.model.parameters(), lr=lr, weight_decay=weight_decay)
        self.sharded = sharded
        self.peer_group = None
        if sharded:
            from .parallel import enable_peer_exchange
            self.peer_group = enable_peer_exchange(graph, max(dense_dim, hidden_dim))
        self._dev_feat = self._dev_lab = None
        self._copy_stream = None

    def _stage(self, features, labels):
        """Returns (features, labels, cells_ready) on the device.  Host inputs: the gene rows (small) and the
        labels are copied on the compute stream; the cell rows (1.2 GB at atlas scale) go H2D on a side
        stream and ``cells_ready`` marks their arrival, so the first cell<-gene pass — which reads only the
        gene table — runs under the copy (pinned host memory makes it asynchronous)."""
        if features.is_cuda:
            return features, labels.to(self.device), None
        if self._dev_feat is None or self._dev_feat.shape != features.shape:
            self._dev_feat = torch.empty(features.shape, dtype=torch.float32, device=self.device)
            self._dev_lab = torch.empty(labels.shape, dtype=torch.int64, device=self.device)
            self._copy_stream = torch.cuda.Stream(device=self.device)
        g = self.graph.num_genes
        main = torch.cuda.current_stream(self.device)
        self._copy_stream.wait_stream(main)             # the previous step's kernels are done with the buffer
        # small copies first: the H2D engine serves requests in issue order, and the compute stream must not
        # queue behind the 1.2 GB transfer
        self._dev_feat[:g].copy_(features[:g], non_blocking=True)
        self._dev_lab.copy_(labels, non_blocking=True)
        with torch.cuda.stream(self._copy_stream):
            self._dev_feat[g:].copy_(features[g:], non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(self._copy_stream)
        return self._dev_feat, self._dev_lab, ready

    def forward_loss(self, features, labels):
        feats, lab, ready = self._stage(features, labels)
        self.model.train()
        if self.sharded:
            from .parallel import sharded_forward
            logits = sharded_forward(self.model, self.graph, feats, cells_ready=ready)
        else:
            logits = self.model(FullGraphFlow(self.graph, feats, cells_ready=ready))
        return cross_entropy_sum(logits, lab), logits

    def step(self, features, labels, return_loss=True):
        loss, _ = self.forward_loss(features, labels)
        self.optimizer.zero_grad(set_to_none=True)
        loss.backward()
        if self.sharded:
            from .parallel import allreduce_grads
            allreduce_grads(self.model)
        self.optimizer.step()
        return float(loss.detach()) if return_loss else loss.detach()

    @torch.no_grad()
    def predict(self, features, unsure_rate=2.0):
        self.model.eval()
        feats = features if features.is_cuda else features.to(self.device)
        return predict_labels(self.model(FullGraphFlow(self.graph, feats)), unsure_rate)
