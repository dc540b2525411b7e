#!/usr/bin/env python
"""Mint golden vectors by running the UNMODIFIED reference code.  TEST INFRASTRUCTURE ONLY.

Run in the build container (needs ``/root/reference``; the GPU box does not have it):

    python oracle/gen_golden.py            # rewrites tests/golden/*.npz

What executes from the reference, byte-for-byte (copied at run time to a scratch directory
under /tmp because its builders write ``pretrained/`` next to themselves and
``/root/reference`` is read-only; nothing is copied into this repository):
  * ``models/gnn.py``                (GNN, NodeUpdate, message_func)
  * ``utils/preprocess_internal.py`` (load_data_internal: training graph + features)
  * ``utils/preprocess.py``          (load_data: inference graph; ``evaluate=False`` because
                                      xlrd/openpyxl are absent)
on top of ``oracle/dgl_shim`` (DGL 0.4.3 cannot be installed) and with ``numpy.str`` aliased to
``str`` (removed in numpy ≥ 1.24; the reference was written for numpy 1.22).

Inputs are sub-sampled from the reference's own fixture ``train/mouse/mouse_Muscle1102``
(3500 genes × 260 training cells + 40 held-out cells written as a test file) so the stored
vectors stay small.  Outputs:
  tests/golden/muscle_train.npz   graph (edges, normalised weights, ids, features), X, PCA gene
                                  features, labels, state_dicts, eval logits for L=1 and L=2,
                                  CE(sum) loss + parameter gradients of one 64-seed batch
  tests/golden/muscle_test.npz    inference graph (support both directions, test cells
                                  gene→cell only), features, logits of the 40 test cells
  tests/golden/synthetic.npz      hand-sized (2 genes × 3 cells) and c1-like (60 cells × 120 genes,
                                  D0 = 18 not a multiple of 4) graphs pushed through the reference GNN
  tests/golden/adipose.npz        BASELINE configs[1] at its stated shape: the reference's FULL fixtures
                                  train/human/human_Adipose1372 (support) + test/human/human_Pancreas11 (test
                                  cells) through the unmodified builders at dense_dim 400, models (L=1,H=200 — the
                                  reference's native shape, predict.py:170-172), (L=2,H=400) and (L=2,H=200):
                                  logits of every training cell and of the 11 test cells, CE(sum) loss and
                                  gradients of one 64-seed batch.  To keep the file small the [N, 400] feature
                                  matrix is NOT the stored PCA output (26 MB) but a seeded recipe applied to the
                                  reference-built graphs (gene rows ~ N(0, 0.66²); cell rows by the reference's own
                                  formula, preprocess_internal.py:190-196), weights come back from their seeds
                                  (checksums stored), big gradient matrices are stored as a strided sample + norm.

The script re-executes itself with PYTHONHASHSEED=0: the reference orders labels by ``list(set(...))``
(preprocess_internal.py:54), so label ids — hence losses and gradients — depend on the hash seed.
"""
import argparse
import os
import shutil
import sys
import tempfile
from pathlib import Path

import numpy as np
import pandas as pd
import scipy.sparse as sp
import torch
import torch.nn.functional as F

REPO = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")
SEED = 10086  # reference default, train.py:128


def _stage_reference(tmp: Path):
    for sub in ("models", "utils"):
        shutil.copytree(REF / sub, tmp / sub)
    if not hasattr(np, "str"):
        np.str = str  # noqa: NPY001 - numpy<1.24 alias the reference relies on
    sys.path.insert(0, str(REPO / "oracle" / "dgl_shim"))
    sys.path.insert(0, str(tmp))


def _write_subsample(tmp: Path, n_genes=3500, n_train=260, n_test=40):
    df = pd.read_csv(REF / "train/mouse/mouse_Muscle1102_data.gz", compression="gzip", index_col=0)
    ct = pd.read_csv(REF / "train/mouse/mouse_Muscle1102_celltype.csv", index_col=0)
    rng = np.random.RandomState(SEED)
    genes = np.sort(rng.choice(df.shape[0], n_genes, replace=False))
    cells = rng.permutation(df.shape[1])[: n_train + n_test]
    tr, te = np.sort(cells[:n_train]), np.sort(cells[n_train:])
    (tmp / "train/mouse").mkdir(parents=True)
    (tmp / "test/mouse").mkdir(parents=True)
    df.iloc[genes, tr].to_csv(tmp / f"train/mouse/mouse_Muscle{n_train}_data.csv")
    ct.iloc[tr].to_csv(tmp / f"train/mouse/mouse_Muscle{n_train}_celltype.csv")
    df.iloc[genes, te].to_csv(tmp / f"test/mouse/mouse_Muscle{n_test}_data.csv")
    return n_train, n_test


def _graph_arrays(g):
    return dict(src=g._src.numpy(), dst=g._dst.numpy(), weight=g.edata["weight"].squeeze(1).numpy(),
                node_id=g.ndata["id"].squeeze(1).numpy(), features=g.ndata["features"].numpy())


def _state(model, tag):
    return {f"{tag}/{k}": v.detach().numpy().copy() for k, v in model.state_dict().items()}


def _make_model(GNN, in_feats, hidden, n_classes, n_layers, gene_num, seed):
    torch.manual_seed(seed)
    model = GNN(in_feats=in_feats, n_hidden=hidden, n_classes=n_classes, n_layers=n_layers,
                gene_num=gene_num, activation=F.relu, dropout=0.0)
    with torch.no_grad():   # α = 1 hides α-indexing bugs: perturb it
        model.alpha.copy_(0.5 + torch.rand(model.alpha.shape))
        model.linear.bias.copy_(torch.rand(model.linear.bias.shape) - 0.5)
    return model


def _eval_logits(model, graph, seeds, n_layers, batch_size, NeighborSampler):
    """The loop of train.py:94-105 / predict.py:64-76 (full neighbour, eval mode)."""
    model.eval()
    out = torch.zeros(graph.number_of_nodes(), model.linear.out_features)
    for nf in NeighborSampler(g=graph, batch_size=batch_size, expand_factor=graph.number_of_nodes(),
                              num_hops=n_layers, neighbor_type='in', shuffle=False, num_workers=8,
                              seed_nodes=seeds):
        nf.copy_from_parent()
        with torch.no_grad():
            logits = model(nf).cpu()
        out[nf.layer_parent_nid(-1).type(torch.long)] = logits
    return out[seeds].numpy()


def _train_grads(model, graph, seeds, labels, n_layers, NeighborSampler):
    """One iteration of train.py:79-84 (dropout=0 so it is deterministic)."""
    model.train()
    nf = next(iter(NeighborSampler(g=graph, batch_size=len(seeds), expand_factor=graph.number_of_nodes(),
                                   num_hops=n_layers, neighbor_type='in', shuffle=False, num_workers=8,
                                   seed_nodes=seeds)))
    nf.copy_from_parent()
    logits = model(nf)
    batch_nids = nf.layer_parent_nid(-1).type(torch.long)
    loss = torch.nn.CrossEntropyLoss(reduction='sum')(logits, labels[batch_nids])
    model.zero_grad()
    loss.backward()
    grads = {k: p.grad.detach().numpy().copy() for k, p in model.named_parameters()}
    return float(loss), logits.detach().numpy(), grads


def real_fixture(out_dir: Path):
    from argparse import Namespace
    tmp = Path(tempfile.mkdtemp(prefix="scds_ref_"))
    _stage_reference(tmp)
    n_train, n_test = _write_subsample(tmp)
    from models import GNN                                  # reference, unmodified
    from utils import load_data_internal, load_data         # reference, unmodified
    from dgl.contrib.sampling import NeighborSampler        # shim

    dense_dim, hidden = 48, 40
    params = Namespace(random_seed=SEED, dense_dim=dense_dim, species="mouse", tissue="Muscle", gpu=-1,
                       filetype="csv", exclude_rate=0.005, threshold=0, test_rate=0.2)
    np.random.seed(SEED); torch.manual_seed(SEED)
    num_cells, num_genes, num_labels, graph, train_ids, test_ids, labels = load_data_internal(params)
    x = sp.load_npz(tmp / "pretrained/mouse/graphs/mouse_Muscle_data.npz").tocsr()
    feats = graph.ndata["features"]
    store = dict(num_cells=num_cells, num_genes=num_genes, num_labels=num_labels,
                 x_data=x.data, x_indices=x.indices, x_indptr=x.indptr, x_shape=np.array(x.shape),
                 gene_feat=feats[:num_genes].numpy(), labels=labels.numpy(),
                 train_ids=train_ids.numpy(), test_ids=test_ids.numpy(), dense_dim=dense_dim, hidden=hidden)
    store.update({f"graph/{k}": v for k, v in _graph_arrays(graph).items()})
    all_cells = torch.arange(num_genes, num_genes + num_cells)
    for n_layers in (1, 2):
        model = _make_model(GNN, dense_dim, hidden, num_labels, n_layers, num_genes, SEED + n_layers)
        store.update(_state(model, f"L{n_layers}"))
        store[f"L{n_layers}/logits"] = _eval_logits(model, graph, all_cells, n_layers, 100, NeighborSampler)
        seeds = train_ids[:64]
        loss, logits, grads = _train_grads(model, graph, seeds, labels, n_layers, NeighborSampler)
        store[f"L{n_layers}/grad_seeds"] = seeds.numpy()
        store[f"L{n_layers}/loss"] = loss
        store[f"L{n_layers}/train_logits"] = logits
        store.update({f"L{n_layers}/grad/{k}": v for k, v in grads.items()})
    np.savez_compressed(out_dir / "muscle_train.npz", **store)
    print(f"muscle_train: G={num_genes} C={num_cells} K={num_labels} E={graph.number_of_edges()}")

    # ---- inference graph through the reference's predict-side builder --------------------
    p2 = Namespace(random_seed=SEED, dense_dim=dense_dim, species="mouse", tissue="Muscle", gpu=-1,
                   filetype="csv", threshold=0, test_dataset=[n_test], test_dir="test", evaluate=False)
    np.random.seed(SEED); torch.manual_seed(SEED)
    total_cell, num_genes2, num_labels2, id2label, test_dict, _ = load_data(p2)
    tg = test_dict["graph"][n_test]
    tdf = pd.read_csv(tmp / f"test/mouse/mouse_Muscle{n_test}_data.csv", index_col=0)
    id2gene = [ln.strip() for ln in open(tmp / "pretrained/mouse/statistics/Muscle_genes.txt", encoding="utf-8")]
    gene2id = {gname: i for i, gname in enumerate(id2gene)}
    dense_t = np.zeros((tdf.shape[1], num_genes2))
    dense_t[:, [gene2id[gname] for gname in tdf.index]] = tdf.to_numpy().T   # file gene order → gene ids
    xt = sp.csr_matrix(dense_t)
    store = dict(num_genes=num_genes2, num_labels=num_labels2, n_support=num_cells, n_test=n_test,
                 x_data=x.data, x_indices=x.indices, x_indptr=x.indptr, x_shape=np.array(x.shape),
                 xt_data=xt.data, xt_indices=xt.indices, xt_indptr=xt.indptr, xt_shape=np.array(xt.shape),
                 test_nid=test_dict["nid"][n_test].numpy(), mask=test_dict["mask"][n_test].numpy(),
                 gene_feat=tg.ndata["features"][:num_genes2].numpy(), dense_dim=dense_dim, hidden=hidden)
    store.update({f"graph/{k}": v for k, v in _graph_arrays(tg).items()})
    for n_layers in (1, 2):
        model = _make_model(GNN, dense_dim, hidden, num_labels2, n_layers, num_genes2, SEED + 10 + n_layers)
        store.update(_state(model, f"L{n_layers}"))
        store[f"L{n_layers}/logits"] = _eval_logits(model, tg, test_dict["nid"][n_test], n_layers, 16, NeighborSampler)
    np.savez_compressed(out_dir / "muscle_test.npz", **store)
    print(f"muscle_test: N={tg.number_of_nodes()} E={tg.number_of_edges()}")
    shutil.rmtree(tmp, ignore_errors=True)
    return GNN, NeighborSampler


def recipe_features(x_all, num_genes, dense_dim, seed):
    """The seeded stand-in for the PCA features (same function in tests/scds_helpers.py): gene rows ~ N(0, 0.66²) from a
    torch CPU generator; cell rows = (x / (rowsum + 1e-6)) · gene_feat in float64, as preprocess_internal.py:190-196."""
    gen = torch.Generator().manual_seed(seed)
    gene_feat = (torch.randn(num_genes, dense_dim, generator=gen) * 0.66).numpy().astype(np.float64)
    dense = np.asarray(x_all.todense(), dtype=np.float64)
    dense = dense / (np.sum(dense, axis=1, keepdims=True) + 1e-6)
    cell_feat = dense.dot(gene_feat)
    return torch.cat([torch.from_numpy(gene_feat), torch.from_numpy(cell_feat)], dim=0).type(torch.float)


def _sample(v, step=53):
    return v.reshape(-1)[::step].copy()


def adipose_fixture(out_dir: Path):
    """BASELINE configs[1]: Adipose1372 (+ Pancreas11) at dense_dim 400 through the unmodified reference."""
    from argparse import Namespace
    tmp = Path(tempfile.mkdtemp(prefix="scds_ref_adipose_"))
    for sub in ("models", "utils"):
        shutil.copytree(REF / sub, tmp / sub)
    (tmp / "train").mkdir(); (tmp / "test").mkdir()
    os.symlink(REF / "train/human", tmp / "train/human")
    (tmp / "test/human").mkdir()
    # the predict-side loader looks for {species}_{tissue}{num}_data.csv: the Pancreas cells are presented as test set 11
    # of the Adipose model (same bytes, staged name)
    os.symlink(REF / "test/human/human_Pancreas11_data.csv", tmp / "test/human/human_Adipose11_data.csv")
    for m in [k for k in sys.modules if k == "utils" or k.startswith("utils.") or k == "models" or k.startswith("models.")]:
        del sys.modules[m]
    sys.path.insert(0, str(tmp))
    from models import GNN                                  # reference, unmodified
    from utils import load_data_internal, load_data         # reference, unmodified
    from dgl.contrib.sampling import NeighborSampler        # shim
    dense_dim, feat_seed = 400, SEED + 77
    params = Namespace(random_seed=SEED, dense_dim=dense_dim, species="human", tissue="Adipose", gpu=-1,
                       filetype="gz", exclude_rate=0.005, threshold=0, test_rate=0.2)
    np.random.seed(SEED); torch.manual_seed(SEED)
    num_cells, num_genes, num_labels, graph, train_ids, test_ids, labels = load_data_internal(params)
    x = sp.load_npz(tmp / "pretrained/human/graphs/human_Adipose_data.npz").tocsr().astype(np.float32)
    pca_feat = graph.ndata["features"]
    graph.ndata["features"] = recipe_features(x, num_genes, dense_dim, feat_seed)
    w = graph.edata["weight"].squeeze(1).double()
    store = dict(num_cells=num_cells, num_genes=num_genes, num_labels=num_labels, dense_dim=dense_dim, feat_seed=feat_seed,
                 x_data=x.data, x_indices=x.indices.astype(np.int32), x_indptr=x.indptr.astype(np.int32), x_shape=np.array(x.shape),
                 labels=labels.numpy().astype(np.int16), train_ids=train_ids.numpy().astype(np.int32),
                 n_edges=graph.number_of_edges(), weight_sum=float(w.sum()), weight_sq_sum=float((w * w).sum()),
                 pca_feat_std=np.array([float(pca_feat[:num_genes].std()), float(pca_feat[num_genes:].std())]),
                 feat_sample=_sample(graph.ndata["features"].numpy(), 997))
    all_cells = torch.arange(num_genes, num_genes + num_cells)
    models = {}
    for tag, n_layers, hidden, seed in (("L1H200", 1, 200, SEED + 31), ("L2H400", 2, 400, SEED + 32), ("L2H200", 2, 200, SEED + 33)):
        model = _make_model(GNN, dense_dim, hidden, num_labels, n_layers, num_genes, seed)
        models[tag] = model
        store[f"{tag}/seed"], store[f"{tag}/n_layers"], store[f"{tag}/hidden"] = seed, n_layers, hidden
        store[f"{tag}/param_sums"] = np.array([float(v.double().sum()) for v in model.state_dict().values()])
        store[f"{tag}/logits"] = _eval_logits(model, graph, all_cells, n_layers, 500, NeighborSampler)
        seeds = train_ids[:64]
        loss, logits, grads = _train_grads(model, graph, seeds, labels, n_layers, NeighborSampler)
        store[f"{tag}/grad_seeds"] = seeds.numpy().astype(np.int32)
        store[f"{tag}/loss"] = loss
        for k, v in grads.items():
            if v.size > 20000:
                store[f"{tag}/grad_sample/{k}"] = _sample(v)
                store[f"{tag}/grad_norm/{k}"] = float(np.sqrt((v.astype(np.float64) ** 2).sum()))
            else:
                store[f"{tag}/grad/{k}"] = v
        print(f"adipose {tag}: loss {loss:.4f}", flush=True)
    # ---- inference graph: Pancreas11 test cells on the Adipose support, reference predict-side builder ----
    p2 = Namespace(random_seed=SEED, dense_dim=dense_dim, species="human", tissue="Adipose", gpu=-1,
                   filetype="csv", threshold=0, test_dataset=[11], test_dir="test", evaluate=False)
    np.random.seed(SEED); torch.manual_seed(SEED)
    total_cell, num_genes2, num_labels2, id2label, test_dict, _ = load_data(p2)
    tg = test_dict["graph"][11]
    tdf = pd.read_csv(REF / "test/human/human_Pancreas11_data.csv", index_col=0)
    id2gene = [ln.strip() for ln in open(tmp / "pretrained/human/statistics/Adipose_genes.txt", encoding="utf-8")]
    gene2id = {gname: i for i, gname in enumerate(id2gene)}
    keep = [gname for gname in tdf.index if gname in gene2id]
    dense_t = np.zeros((tdf.shape[1], num_genes2), dtype=np.float32)
    dense_t[:, [gene2id[gname] for gname in keep]] = tdf.loc[keep].to_numpy().T
    xt = sp.csr_matrix(dense_t)
    tg.ndata["features"] = recipe_features(sp.vstack([x, xt]).tocsr(), num_genes2, dense_dim, feat_seed)
    wt = tg.edata["weight"].squeeze(1).double()
    store.update(xt_data=xt.data, xt_indices=xt.indices.astype(np.int32), xt_indptr=xt.indptr.astype(np.int32), xt_shape=np.array(xt.shape),
                 test_nid=test_dict["nid"][11].numpy().astype(np.int32), test_n_edges=tg.number_of_edges(),
                 test_weight_sum=float(wt.sum()), n_test_genes_shared=len(keep))
    for tag, model in models.items():
        store[f"{tag}/test_logits"] = _eval_logits(model, tg, test_dict["nid"][11], int(store[f"{tag}/n_layers"]), 500, NeighborSampler)
    np.savez_compressed(out_dir / "adipose.npz", **store)
    print(f"adipose: G={num_genes} C={num_cells} K={num_labels} E={graph.number_of_edges()} / test N={tg.number_of_nodes()} "
          f"E={tg.number_of_edges()} shared genes {len(keep)} nnz_test {xt.nnz}")
    shutil.rmtree(tmp, ignore_errors=True)


def _shim_graph(og):
    """OracleGraph → shim DGLGraph with the reference's frame layout."""
    import dgl
    g = dgl.DGLGraph()
    g.add_nodes(og.num_nodes, {"id": og.node_id.unsqueeze(-1)})
    g.add_edges(og.src, og.dst, {"weight": og.weight.unsqueeze(1)})
    g.ndata["features"] = og.features
    g.readonly()
    return g


def synthetic_fixture(out_dir: Path, GNN, NeighborSampler):
    sys.path.insert(0, str(REPO))
    from oracle.graph_oracle import build_graph, make_features
    store = {}
    # (i) hand-sized graph: 2 genes × 3 cells (SURVEY §8c item i)
    x_tiny = sp.csr_matrix(np.array([[1.0, 2.0], [0.0, 3.0], [4.0, 0.0]]))
    # (ii) c1-like, small: D0 deliberately not a multiple of 4, some all-zero genes
    rng = np.random.RandomState(SEED)
    dense = (rng.rand(60, 120) < 0.15) * np.clip(rng.normal(3.0, 0.5, (60, 120)), 0.05, 9)
    dense[:, rng.choice(120, 9, replace=False)] = 0
    dense[np.arange(60), rng.randint(0, 120, 60)] += 1.0   # every cell has ≥ 1 gene
    for tag, xm, d0, hid, k in (("tiny", x_tiny, 3, 4, 2), ("c1s", sp.csr_matrix(dense), 18, 10, 5)):
        og = build_graph(xm)
        gene_feat = rng.normal(0, 0.66, (og.num_genes, d0))
        og.features = make_features(xm, gene_feat)
        g = _shim_graph(og)
        seeds = torch.arange(og.num_genes, og.num_nodes)
        labels = torch.from_numpy(rng.randint(0, k, og.num_nodes))
        store.update({f"{tag}/x": np.asarray(xm.todense()), f"{tag}/gene_feat": gene_feat,
                      f"{tag}/labels": labels.numpy()})
        store.update({f"{tag}/graph/{kk}": v for kk, v in _graph_arrays(g).items()})
        for n_layers in (1, 2, 3):
            model = _make_model(GNN, d0, hid, k, n_layers, og.num_genes, SEED + 20 + n_layers)
            store.update(_state(model, f"{tag}/L{n_layers}"))
            store[f"{tag}/L{n_layers}/logits"] = _eval_logits(model, g, seeds, n_layers, 25, NeighborSampler)
            loss, logits, grads = _train_grads(model, g, seeds[:17], labels, n_layers, NeighborSampler)
            store[f"{tag}/L{n_layers}/loss"] = loss
            store.update({f"{tag}/L{n_layers}/grad/{kk}": v for kk, v in grads.items()})
    np.savez_compressed(out_dir / "synthetic.npz", **store)
    print("synthetic: tiny + c1s written")


if __name__ == "__main__":
    if os.environ.get("PYTHONHASHSEED") != "0":             # label ids follow set order (preprocess_internal.py:54)
        os.environ["PYTHONHASHSEED"] = "0"
        os.execv(sys.executable, [sys.executable] + sys.argv)
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=str(REPO / "tests" / "golden"))
    ap.add_argument("--only", default="", help="comma list of: muscle, synthetic, adipose (default: all)")
    args = ap.parse_args()
    out = Path(args.out)
    out.mkdir(parents=True, exist_ok=True)
    only = set(filter(None, args.only.split(","))) or {"muscle", "synthetic", "adipose"}
    gnn_cls = sampler_cls = None
    if only & {"muscle", "synthetic"}:
        gnn_cls, sampler_cls = real_fixture(out if "muscle" in only else Path(tempfile.mkdtemp()))
    if "synthetic" in only:
        synthetic_fixture(out, gnn_cls, sampler_cls)
    if "adipose" in only:
        if gnn_cls is None:
            if not hasattr(np, "str"):
                np.str = str  # noqa: NPY001
            sys.path.insert(0, str(REPO / "oracle" / "dgl_shim"))
        adipose_fixture(out)
