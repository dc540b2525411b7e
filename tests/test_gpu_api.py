"""The docs-level API (docs/api.rst) end to end on synthetic csv files: fit -> artefacts -> predict."""
import numpy as np
import pandas as pd
import pytest
import torch

from scdeepsort_b200.api import DeepSortClassifier, DeepSortPredictor

pytestmark = pytest.mark.gpu


def _write_dataset(tmp, n_genes=160, n_cells=180, seed=0):
    rng = np.random.default_rng(seed)
    types = ["T cell", "B cell", "Macrophage"]
    genes = [f"G{i:03d}" for i in range(n_genes)]
    lab = rng.integers(0, 3, n_cells)
    base = rng.random((n_genes, n_cells)) < 0.08
    for t in range(3):                                   # each type switches on its own block of genes
        blk = slice(t * 40, t * 40 + 40)
        base[blk][:, lab == t] |= rng.random((40, int((lab == t).sum()))) < 0.6
    vals = np.where(base, np.clip(rng.normal(3.0, 0.5, base.shape), 0.05, 9), 0.0)
    cells = [f"C_{i}" for i in range(n_cells)]
    df = pd.DataFrame(vals, index=genes, columns=cells)
    n_tr = 140
    df.iloc[:, :n_tr].to_csv(tmp / "mouse_Demo140_data.csv")
    pd.DataFrame({"Cell": cells[:n_tr], "Cell_type": [types[i] + " " for i in lab[:n_tr]]},
                 index=range(1, n_tr + 1)).to_csv(tmp / "mouse_Demo140_celltype.csv")
    # the test file has extra unknown genes and misses some training genes (gene intersection path)
    te = df.iloc[10:, n_tr:].copy()
    te.loc["UNKNOWN1"] = 1.0
    te.to_csv(tmp / "mouse_Demo40_data.csv")
    return [types[i] for i in lab[n_tr:]]


def test_fit_save_predict_roundtrip(tmp_path):
    truth = _write_dataset(tmp_path)
    clf = DeepSortClassifier(species="mouse", tissue="Demo", dense_dim=48, hidden_dim=32, batch_size=64, dropout=0.1,
                             gpu_id=0, n_epochs=80, n_layers=1, random_seed=1, validation_fraction=0.15,
                             learning_rate=5e-3)
    best = clf.fit([(str(tmp_path / "mouse_Demo140_data.csv"), str(tmp_path / "mouse_Demo140_celltype.csv"))],
                   save_path=str(tmp_path / "model"))
    assert best["train_acc"] > 0.9
    root = tmp_path / "model"
    raw = (root / "statistics" / "Demo_genes.txt").read_bytes()
    assert raw.count(b"\r\n") == 160                                            # reference line endings
    state = torch.load(root / "models" / "mouse-Demo.pt", map_location="cpu")
    assert set(state) == {"model", "optimizer"}
    assert set(state["model"]) == {"alpha", "layers.0.fc_neigh.weight", "layers.0.fc_neigh.bias", "linear.weight", "linear.bias"}
    assert state["model"]["alpha"].shape == (162, 1)
    df = clf.predict(str(tmp_path / "mouse_Demo40_data.csv"), model_path=str(root), save_path=str(tmp_path / "out"))
    assert list(df.columns) == ["index", "cell_type", "cell_subtype"] and len(df) == 40
    acc = np.mean([p == t for p, t in zip(df["cell_type"], truth)])
    assert acc > 0.8
    assert (tmp_path / "out" / "mouse_Demo_mouse_Demo40_data.csv").exists()
    # stand-alone predictor on the saved artefacts, unsure disabled
    df2 = DeepSortPredictor("mouse", "Demo", unsure_rate=0.0, model_path=str(root), dense_dim=48, hidden_dim=32,
                            gpu_id=0).predict(str(tmp_path / "mouse_Demo40_data.csv"))
    assert "unsure" not in set(df2["cell_type"])
    with pytest.raises(FileNotFoundError):
        DeepSortPredictor("mouse", "Demo")


def test_sampled_training_two_layers(tmp_path):
    """num_neighbors > 0 (train.py:37-40): sampled NodeFlows through the generic gather kernels, 2 layers, dropout."""
    _write_dataset(tmp_path, seed=3)
    clf = DeepSortClassifier(species="mouse", tissue="Demo", dense_dim=32, hidden_dim=32, batch_size=50, dropout=0.1,
                             gpu_id=0, n_epochs=80, n_layers=2, num_neighbors=12, random_seed=2, learning_rate=5e-3)
    best = clf.fit([(str(tmp_path / "mouse_Demo140_data.csv"), str(tmp_path / "mouse_Demo140_celltype.csv"))])
    assert best["train_acc"] > 0.85


def test_runner_full_graph_fast_path_matches_nodeflow_loop(golden_test):
    """predict.py:61-88 on the reference-built inference graph: the batched NodeFlow loop and the one-pass full-graph
    form (Runner(bipartite=...), what DeepSortPredictor uses) give the same logits and the same labels."""
    import scdeepsort_b200 as sd
    from scdeepsort_b200.trainer import Runner
    from scds_helpers import golden_csr, golden_graph, golden_state, rel_err
    z = golden_test
    gg = golden_graph(z, num_genes=int(z["num_genes"]))
    g = sd.DeepSortGraph.from_edges(gg.src, gg.dst, gg.weight, gg.node_id, gg.features, gg.num_genes)
    nid = torch.from_numpy(z["test_nid"])
    kw = dict(dense_dim=int(z["dense_dim"]), hidden_dim=int(z["hidden"]), n_layers=2, batch_size=16, unsure_rate=2.0,
              device="cuda:0", state_dict=golden_state(z, "L2"))
    pred_a, logits_a = Runner(g, nid, int(z["num_labels"]), **kw).inference()
    bg = sd.BipartiteGraph.from_expression(golden_csr(z), golden_csr(z, "xt"), device="cuda:0").densify(0.1)
    pred_b, logits_b = Runner(g, nid, int(z["num_labels"]), bipartite=bg, **kw).inference()
    assert rel_err(logits_a.cpu(), z["L2/logits"]) < 1e-5 and rel_err(logits_b.cpu(), z["L2/logits"]) < 2e-5
    assert torch.equal(pred_a, pred_b)
