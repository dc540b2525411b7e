"""``DGLGraph`` restatement (mutable multigraph with node/edge frames), DGL 0.4 semantics."""
import numpy as np
import torch


def _as_index(x):
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().to(torch.int64).reshape(-1)
    return torch.as_tensor(np.asarray(x), dtype=torch.int64).reshape(-1)


class _Frame(dict):
    """Column store.  ``frame[k]`` returns the stored tensor itself (no copy), which is what
    lets the reference's ``graph.edata['weight'][eids] = ...`` mutate the graph in place
    (utils/preprocess_internal.py:23)."""

    def __init__(self, owner_len):
        super().__init__()
        self._len = owner_len

    def __setitem__(self, key, value):
        assert value.shape[0] == self._len(), \
            f"frame column {key!r} has {value.shape[0]} rows, expected {self._len()}"
        super().__setitem__(key, value)

    def append_rows(self, n, data):
        data = data or {}
        for key in set(self.keys()) | set(data.keys()):
            new = data.get(key)
            old = self.get(key)
            if old is None:
                # column first appears now: earlier rows are zero-filled (DGL behaviour)
                assert new is not None
                pad = torch.zeros((self._len() - n,) + tuple(new.shape[1:]), dtype=new.dtype, device=new.device)
                super().__setitem__(key, torch.cat([pad, new], dim=0))
            else:
                if new is None:
                    new = torch.zeros((n,) + tuple(old.shape[1:]), dtype=old.dtype, device=old.device)
                super().__setitem__(key, torch.cat([old, new.to(old.dtype)], dim=0))


class DGLGraph:
    def __init__(self):
        self._n = 0
        self._src = torch.zeros(0, dtype=torch.int64)
        self._dst = torch.zeros(0, dtype=torch.int64)
        self.ndata = _Frame(lambda: self._n)
        self.edata = _Frame(lambda: int(self._src.shape[0]))
        self._readonly = False
        self._in_csr = None

    # -- mutation ---------------------------------------------------------------------
    def add_nodes(self, num, data=None):
        assert not self._readonly
        self._n += int(num)
        self.ndata.append_rows(int(num), data)

    def add_edges(self, u, v, data=None):
        assert not self._readonly
        u, v = _as_index(u), _as_index(v)
        assert u.shape == v.shape
        self._src = torch.cat([self._src, u])
        self._dst = torch.cat([self._dst, v])
        self.edata.append_rows(int(u.shape[0]), data)
        self._in_csr = None

    def readonly(self, readonly_state=True):
        self._readonly = readonly_state
        return self

    # -- queries ----------------------------------------------------------------------
    def number_of_nodes(self):
        return self._n

    def number_of_edges(self):
        return int(self._src.shape[0])

    def nodes(self):
        return torch.arange(self._n, dtype=torch.int64)

    def in_degrees(self, v=None):
        deg = torch.bincount(self._dst, minlength=self._n)
        return deg if v is None else deg[_as_index(v)]

    def _build_in_csr(self):
        if self._in_csr is None:
            # stable sort keeps edges of one destination in insertion (edge id) order
            order = torch.sort(self._dst, stable=True).indices
            ptr = torch.zeros(self._n + 1, dtype=torch.int64)
            ptr[1:] = torch.cumsum(torch.bincount(self._dst, minlength=self._n), 0)
            self._in_csr = (ptr, order)
        return self._in_csr

    def in_edges(self, v, form='uv'):
        ptr, order = self._build_in_csr()
        if isinstance(v, (int, np.integer)):
            eid = order[ptr[v]:ptr[v + 1]]
        else:
            eid = torch.cat([order[ptr[i]:ptr[i + 1]] for i in _as_index(v).tolist()])
        if form == 'all':
            return self._src[eid], self._dst[eid], eid
        if form == 'eid':
            return eid
        return self._src[eid], self._dst[eid]


class EdgeBatch:
    """What a message UDF receives: ``src``/``dst`` node data gathered per edge, ``data`` edge data."""

    def __init__(self, src_data, dst_data, edge_data):
        self.src = src_data
        self.dst = dst_data
        self.data = edge_data

    def batch_size(self):
        return next(iter(self.data.values())).shape[0]


class NodeBatch:
    def __init__(self, data):
        self.data = data
