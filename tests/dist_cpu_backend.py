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


class CpuCheckerIcpBackend:
    """Checker-backed stand-in for distributed.GpuIcpBackend (oracle NN search + NumPy solve)."""

    def set_model(self, model_planar):
        self.model = model_planar.numpy().copy()

    def begin(self, data_planar):
        self.n = data_planar.shape[1]
        self.R, self.T, self.have_rt = np.zeros((3, 3)), np.zeros(3), False
        self.round, self.done, self.converged, self.d = 0, False, False, 0.0
        self.d2 = torch.zeros(self.n, dtype=torch.float64)
        self.d2g = torch.zeros(self.n, dtype=torch.float64)
        self.idx = torch.zeros(self.n, dtype=torch.int32)
        self.sums = torch.zeros(16, dtype=torch.float64)

    def _P(self, data_planar):
        d = data_planar.numpy()
        return oracle_py.trans_points(d, self.R, self.T) if self.have_rt else d.copy()

    def nn(self, data_planar, idx_offset):
        if not self.done:
            order, sq = oracle_py.closest_point_set(self.model, self._P(data_planar), "grid")
            self.d2.copy_(torch.from_numpy(sq))
            self.idx.copy_(torch.from_numpy(order + idx_offset))
        return self.d2, self.idx

    def select(self, d2_local, d2_global, idx):
        if not self.done:
            idx[d2_local != d2_global] = np.iinfo(np.int32).max

    def accumulate(self, data_planar, idx_global, idx_offset):
        if not self.done:
            j = idx_global.numpy().astype(np.int64) - idx_offset
            mine = (j >= 0) & (j < self.model.shape[1])
            P, Y = self._P(data_planar)[:, mine], self.model[:, j[mine]]
            s = np.concatenate([P.sum(1), Y.sum(1), (P[:, None, :] * Y[None, :, :]).sum(2).ravel(), [((P - Y) ** 2).sum()]])
            self.sums.copy_(torch.from_numpy(s))
        return self.sums

    def solve(self, sums, e, max_iters):
        if not self.done:
            S, N = sums.numpy(), float(self.n)
            mp, my = S[0:3] / N, S[3:6] / N
            m = S[6:15].reshape(3, 3) / N - np.outer(mp, my)
            A = m - m.T
            tr = np.trace(m)
            Q = np.empty((4, 4))
            Q[0, 0] = tr
            Q[0, 1:] = Q[1:, 0] = [A[1, 2], A[2, 0], A[0, 1]]
            Q[1:, 1:] = m + m.T - tr * np.eye(3)
            w, v = np.linalg.eigh(Q)
            q = v[:, -1] * (1 if v[0, -1] >= 0 else -1)
            R1 = np.array([[q[0]**2 + q[1]**2 - q[2]**2 - q[3]**2, 2 * (q[1]*q[2] - q[0]*q[3]), 2 * (q[1]*q[3] + q[0]*q[2])],
                           [2 * (q[1]*q[2] + q[0]*q[3]), q[0]**2 - q[1]**2 + q[2]**2 - q[3]**2, 2 * (q[2]*q[3] - q[0]*q[1])],
                           [2 * (q[1]*q[3] - q[0]*q[2]), 2 * (q[2]*q[3] + q[0]*q[1]), q[0]**2 - q[1]**2 - q[2]**2 + q[3]**2]])
            T1 = my - R1 @ mp
            pre_d, self.d = self.d, float(S[15])
            self.round += 1
            if abs(self.d - pre_d) >= e:
                if self.round == 1:
                    self.R, self.T = R1, T1
                else:
                    self.R, self.T = R1 @ self.R, R1 @ self.T + T1
                self.have_rt = True
            else:
                self.converged = True
                self.done = True
            if max_iters > 0 and self.round >= max_iters:
                self.done = True
        st = np.concatenate([self.R.ravel(), self.T, [self.d, self.round, float(self.converged), 0.0]])
        return torch.from_numpy(st)


class CpuPipelineBackend:
    """Checker-backed stand-in for vtkcloudpoint_b200.pipeline.GpuPipelineBackend (CPU oracle underneath)."""

    def __init__(self):
        self.db = CpuCheckerBackend()
        self.icp = CpuCheckerIcpBackend()

    def dbscan_single(self, mx, my, eps, min_pts):
        cid, key, cls, amount = oracle_py.dbscan(mx.numpy(), my.numpy(), eps, min_pts, 0, variant="grid")
        return torch.from_numpy(cid), torch.from_numpy(key), amount

    def cluster_stats(self, cid_local, k_local, vals5):
        v = vals5.numpy()
        r = oracle_py.cluster_stats(cid_local.numpy(), k_local, v[0:3], v[3], v[4], circles3d=True, circles2d=False)
        return torch.from_numpy(r["means"]), torch.from_numpy(r["counts"]), torch.from_numpy(r["circle3d"]), torch.from_numpy(r["status3d"])

    def radius_flag(self, circ, status, k, thr):
        f = (status.numpy() == 1) & (circ.numpy()[2] > thr)
        f[0] = False
        return torch.from_numpy(f.astype(np.uint8))

    def match_within(self, truth_planar, centres_planar, match_distance):
        mid, d = oracle_py.match_within(truth_planar.numpy(), centres_planar.numpy(), match_distance)
        return torch.from_numpy(mid), torch.from_numpy(d)
