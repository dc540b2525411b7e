"""``DeepSortClassifier`` / ``DeepSortPredictor``: the pip-package façade documented in
/root/reference/docs/api.rst:6-130 (its implementation ships in a release tarball that is not in the
reference tree; semantics follow the in-tree CLI classes ``Trainer`` train.py:16-123 and ``Runner``
predict.py:17-152).  Same constructor keywords and method signatures; the hot path underneath is
``scdeepsort_b200`` (CUDA only — ``gpu_id=-1`` selects ``cuda:0`` because there is no CPU path).

File handling restates the reference's builders in vectorised form:
gene × cell csv/gz tables (docs/input_requirement.rst), gene set = sorted union
(utils/preprocess_internal.py:26-41), labels rarer than ``exclude_rate`` dropped with their cells
(:94-97,126-129), PCA gene features fit on the training matrix (:186-187; utils/preprocess.py:196-197),
cell features = row-normalised expression · gene features (:194-196, through ``wsage_spmm``),
``statistics`` text files with ``\\r\\n`` endings (:59-67) and the ``{'model', 'optimizer'}`` checkpoint
(train.py:117-123).  The celltype→subtype xlsx map (predict.py:125-133) needs xlrd/openpyxl, which are
absent here: ``cell_subtype`` is filled with ``'N/A'`` unless a mapping dict is supplied.
"""
from pathlib import Path
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import pandas as pd
import scipy.sparse as sp
import torch

from .features import cell_features, pca_gene_features
from .graph import BipartiteGraph, DeepSortGraph
from .trainer import Runner, Trainer


def _device(gpu_id):
    if not torch.cuda.is_available():
        raise RuntimeError("scdeepsort_b200 needs a CUDA device (no CPU fallback)")
    return torch.device("cuda", 0 if gpu_id is None or gpu_id < 0 else gpu_id)


def _read_table(path, file_type):
    """gene × cell table with header and index (docs/input_requirement.rst)."""
    kw = dict(index_col=0)
    if file_type == "gz":
        kw["compression"] = "gzip"
    return pd.read_csv(path, **kw)


def _features(x_support: sp.csr_matrix, x_test, gene_feat, dense_dim, seed, device, return_graph=False):
    """cat[gene_feat; (X / (rowsum+1e-6)) · gene_feat]  (preprocess_internal.py:186-202; preprocess.py:196-208).  Neither
    the PCA nor the product ever sees a dense [C, G] array: both run as sparse x thin-dense passes of the aggregation
    kernels (features.py).  ``gene_feat`` None: fit the PCA on the support cells."""
    bg = BipartiteGraph.from_expression(x_support, x_test, device=device)
    if gene_feat is None:
        gf = pca_gene_features(bg, dense_dim, seed=10086 if seed is None else seed)
    else:
        gf = torch.as_tensor(np.ascontiguousarray(gene_feat), dtype=torch.float32, device=device)
    feats = torch.cat([gf, cell_features(bg, gf)], dim=0)
    return (feats, gf, bg) if return_graph else (feats, gf)


class DeepSortClassifier:
    def __init__(self, species, tissue, dense_dim=400, hidden_dim=200, batch_size=256, dropout=0.1, gpu_id=-1,
                 file_type='csv', learning_rate=1e-3, weight_decay=5e-4, n_epochs=300, n_layers=1, threshold=0,
                 num_neighbors=None, exclude_rate=0.005, random_seed=None, validation_fraction=0.1):
        self.species, self.tissue = species, tissue
        self.dense_dim, self.hidden_dim, self.batch_size, self.dropout = dense_dim, hidden_dim, batch_size, dropout
        self.device = _device(gpu_id)
        self.file_type, self.lr, self.weight_decay = file_type, learning_rate, weight_decay
        self.n_epochs, self.n_layers, self.threshold = n_epochs, n_layers, threshold
        self.num_neighbors = 0 if num_neighbors is None else num_neighbors
        self.exclude_rate, self.random_seed, self.validation_fraction = exclude_rate, random_seed, validation_fraction
        self.trainer: Optional[Trainer] = None

    # -- training --------------------------------------------------------------------------------------
    def _load_training(self, files: Sequence[Tuple[str, str]]):
        tables, types = [], []
        for data_file, type_file in files:
            tables.append(_read_table(data_file, self.file_type))
            ct = pd.read_csv(type_file, index_col=0)
            ct.columns = ['cell', 'type']
            ct['type'] = ct['type'].map(str.strip)
            types.append(ct)
        id2gene = sorted(set().union(*[set(map(str, t.index)) for t in tables]))
        gene2id = {g: i for i, g in enumerate(id2gene)}
        counts = pd.concat([t['type'] for t in types]).value_counts()
        total = int(counts.sum())
        id2label = sorted(lbl for lbl, n in counts.items() if n / total > self.exclude_rate)
        label2id = {lbl: i for i, lbl in enumerate(id2label)}
        mats, labels = [], []
        for tab, ct in zip(tables, types):
            keep = ct['type'].isin(label2id).to_numpy()
            if len(ct) != tab.shape[1]:
                raise ValueError("cell type file does not match the data file columns")
            arr = np.nan_to_num(tab.to_numpy(dtype=np.float64).T[keep])           # cells × file genes
            cols = np.array([gene2id[str(g)] for g in tab.index])
            m = sp.lil_matrix((arr.shape[0], len(id2gene)))
            m[:, cols] = np.where(arr > self.threshold, arr, 0.0)
            mats.append(m.tocsr())
            labels += [label2id[t] for t in ct['type'][keep]]
        return sp.vstack(mats).tocsr(), np.asarray(labels, dtype=np.int64), id2gene, id2label

    def fit(self, files, save_path=None):
        if self.random_seed is not None:
            np.random.seed(self.random_seed)
            torch.manual_seed(self.random_seed)
        x, labels, id2gene, id2label = self._load_training(files)
        num_cells, num_genes = x.shape
        feats, gene_feat = _features(x, None, None, self.dense_dim, self.random_seed, self.device)   # tiny inputs: zero-padded
        gene_feat = gene_feat.cpu().numpy()
        graph = DeepSortGraph.from_expression(x, threshold=self.threshold, features=feats.cpu())
        perm = np.random.permutation(np.arange(num_genes, num_genes + num_cells))
        n_val = int(num_cells * self.validation_fraction)
        all_labels = torch.cat([torch.full((num_genes,), -1, dtype=torch.int64), torch.from_numpy(labels)])
        model_file = None
        if save_path is not None:
            model_file = Path(save_path) / 'models' / f'{self.species}-{self.tissue}.pt'
        self.trainer = Trainer(graph, all_labels, perm[n_val:], perm[:n_val], len(id2label), dense_dim=self.dense_dim,
                               hidden_dim=self.hidden_dim, n_layers=self.n_layers, dropout=self.dropout, lr=self.lr,
                               weight_decay=self.weight_decay, batch_size=self.batch_size,
                               num_neighbors=self.num_neighbors, device=self.device, save_path=model_file)
        best = self.trainer.fit(self.n_epochs, verbose=False)      # saves the best-on-validation checkpoint (train.py:53-58)
        if model_file is not None and not model_file.exists():      # n_epochs == 0: nothing was selected, keep the initial weights
            self.trainer.save_model()
        if save_path is not None:
            _save_artifacts(Path(save_path), self.species, self.tissue, x, gene_feat, id2gene, id2label)
        self._support = (x, gene_feat, id2gene, id2label)
        return best

    # -- inference -------------------------------------------------------------------------------------
    def predict(self, input_file, model_path, save_path=None, unsure_rate=2., file_type='csv'):
        pred = DeepSortPredictor(self.species, self.tissue, file_type=file_type, unsure_rate=unsure_rate,
                                 model_path=model_path, dense_dim=self.dense_dim, hidden_dim=self.hidden_dim,
                                 n_layers=self.n_layers, batch_size=self.batch_size, gpu_id=self.device.index)
        return pred.predict(input_file, save_path=save_path)


def _save_artifacts(root: Path, species, tissue, x, gene_feat, id2gene, id2label):
    """pretrained/{species}/graphs|statistics layout of the reference (preprocess_internal.py:59-67,180)."""
    (root / 'graphs').mkdir(parents=True, exist_ok=True)
    (root / 'statistics').mkdir(parents=True, exist_ok=True)
    sp.save_npz(root / 'graphs' / f'{species}_{tissue}_data', x)
    np.save(root / 'graphs' / f'{species}_{tissue}_gene_feat.npy', gene_feat)
    with open(root / 'statistics' / f'{tissue}_genes.txt', 'w', encoding='utf-8', newline='') as f:
        f.writelines(g + '\r\n' for g in id2gene)
    with open(root / 'statistics' / f'{tissue}_cell_type.txt', 'w', encoding='utf-8', newline='') as f:
        f.writelines(lbl + '\r\n' for lbl in id2label)


class DeepSortPredictor:
    """``DeepSortPredictor(species, tissue, file_type='csv', unsure_rate=2.).predict(input_file, save_path=None)``.
    The published package bundles pretrained artefacts; they are not in the reference tree, so the directory
    written by ``DeepSortClassifier.fit(save_path=...)`` (or a reference ``pretrained/{species}`` tree) is passed
    as ``model_path``."""

    def __init__(self, species, tissue, file_type='csv', unsure_rate=2., model_path=None, dense_dim=400,
                 hidden_dim=200, n_layers=1, batch_size=500, gpu_id=-1, subtype_map: Optional[Dict[str, Tuple[str, str]]] = None):
        if model_path is None:
            raise FileNotFoundError("no bundled pretrained models: pass model_path=<dir written by DeepSortClassifier.fit>")
        self.root = Path(model_path)
        self.species, self.tissue, self.file_type, self.unsure_rate = species, tissue, file_type, unsure_rate
        self.dense_dim, self.hidden_dim, self.n_layers, self.batch_size = dense_dim, hidden_dim, n_layers, batch_size
        self.device = _device(gpu_id)
        self.subtype_map = subtype_map or {}
        read = lambda p: [ln.strip() for ln in open(p, encoding='utf-8') if ln.strip()]      # noqa: E731
        self.id2gene = read(self.root / 'statistics' / f'{tissue}_genes.txt')                # preprocess.py:43-56
        self.id2label = read(self.root / 'statistics' / f'{tissue}_cell_type.txt')
        self.support = sp.load_npz(self.root / 'graphs' / f'{species}_{tissue}_data.npz').tocsr()
        gf = self.root / 'graphs' / f'{species}_{tissue}_gene_feat.npy'
        self.gene_feat = np.load(gf) if gf.exists() else None
        self.state = torch.load(self.root / 'models' / f'{species}-{tissue}.pt', map_location='cpu')['model']

    def predict(self, input_file, save_path=None) -> pd.DataFrame:
        tab = _read_table(input_file, self.file_type)
        gene2id = {g: i for i, g in enumerate(self.id2gene)}
        known = [i for i, g in enumerate(map(str, tab.index)) if g in gene2id]          # gene intersection
        arr = np.nan_to_num(tab.to_numpy(dtype=np.float64).T[:, known])
        xt = sp.lil_matrix((arr.shape[0], len(self.id2gene)))
        xt[:, [gene2id[str(tab.index[i])] for i in known]] = np.where(arr > 0, arr, 0.0)
        xt = xt.tocsr()
        # reference artefacts carry no gene features: PCA on the support cells only (preprocess.py:196)
        feats, _, bg = _features(self.support, xt, self.gene_feat, self.dense_dim, 10086, self.device, return_graph=True)
        graph = DeepSortGraph.from_expression(self.support, xt, features=feats.cpu())
        g, ns = graph.num_genes, self.support.shape[0]
        nid = torch.arange(g + ns, g + ns + xt.shape[0])
        runner = Runner(graph, nid, len(self.id2label), dense_dim=self.dense_dim, hidden_dim=self.hidden_dim,
                        n_layers=self.n_layers, batch_size=self.batch_size, unsure_rate=self.unsure_rate,
                        device=self.device, state_dict=self.state, bipartite=bg)
        pred, _ = runner.inference()
        names = ['unsure' if p < 0 else self.id2label[p] for p in pred.cpu().tolist()]
        df = pd.DataFrame({'index': list(tab.columns),
                           'cell_type': [self.subtype_map.get(n, (n, 'N/A'))[0] for n in names],
                           'cell_subtype': [self.subtype_map.get(n, (n, 'N/A'))[1] for n in names]})
        if save_path is not None:
            Path(save_path).mkdir(parents=True, exist_ok=True)
            df.to_csv(Path(save_path) / f'{self.species}_{self.tissue}_{Path(input_file).stem}.csv', index=False)
        return df
