"""TEST INFRASTRUCTURE: a checker-backed stand-in for vtkcloudpoint_b200.distributed.GpuBackend so that the
multi-rank HOST logic (partition, halo exchange, merge, numbering) can run under gloo without a GPU.
It is built on the CPU oracle and NumPy/SciPy and is never imported by the product package."""
import numpy as np
import torch

import oracle_py


class CpuCheckerBackend:
    def slab_local(self, x, y, gidx, eps, min_pts):
        self.x, self.y, self.g = x.numpy().copy(), y.numpy().copy(), gidx.numpy().copy()
        self.eps = eps
        cid, key, cls, amount = oracle_py.dbscan(self.x, self.y, eps, min_pts, 0, variant="grid")
        self.core = key.astype(bool)
        mins = np.full(amount + 1, np.iinfo(np.int32).max, np.int64)
        np.minimum.at(mins, cid[self.core], self.g[self.core])
        self.key = np.where(self.core, mins[cid], -1).astype(np.int32)
        return torch.from_numpy(key.copy()), torch.from_numpy(self.key.copy())

    def slab_finish(self, map_from, map_to):
        mf, mt = map_from.numpy(), map_to.numpy()
        key = self.key.copy()
        if mf.size:
            pos = np.searchsorted(mf, key)
            pos = np.clip(pos, 0, mf.size - 1)
            hit = (mf[pos] == key) & self.core
            key[hit] = mt[pos[hit]]
        out = key.copy()
        cx, cy, ck = self.x[self.core], self.y[self.core], key[self.core]
        for i in np.flatnonzero(~self.core):
            d = np.abs(self.x[i] - cx) + np.abs(self.y[i] - cy)      # reference predicate, same operand order
            near = d <= self.eps
            out[i] = ck[near].max() if near.any() else -1
        return torch.from_numpy(out.astype(np.int32))

    def uf_edges(self, a, b, n_nodes):
        from scipy.sparse import coo_matrix
        from scipy.sparse.csgraph import connected_components
        a, b = a.numpy().astype(np.int64), b.numpy().astype(np.int64)
        g = coo_matrix((np.ones(a.size, np.int8), (a, b)), shape=(n_nodes, n_nodes))
        _, lab = connected_components(g, directed=False)
        mins = np.full(lab.max() + 1, n_nodes, np.int64)
        np.minimum.at(mins, lab, np.arange(n_nodes))
        return torch.from_numpy(mins[lab].astype(np.int32))
