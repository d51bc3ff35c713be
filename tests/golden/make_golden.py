"""Generates tests/golden/*.npz.  The reference has no golden vectors and cannot run here, so these are produced by
the literal oracle and CROSS-CHECKED at generation time against independent implementations (scikit-learn
DBSCAN(metric='manhattan'), NumPy brute-force argmin, Kabsch SVD); the script refuses to write a fixture that fails
its cross-check.  Run: python tests/golden/make_golden.py"""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle"))
import oracle_py as O  # noqa: E402
from vtkcloudpoint_b200 import synth  # noqa: E402


def sk_check(mx, my, eps, min_pts, cid, key):
    from sklearn.cluster import DBSCAN
    fin = np.isfinite(mx) & np.isfinite(my)
    db = DBSCAN(eps=eps, min_samples=min_pts, metric="manhattan", algorithm="brute").fit(np.c_[mx[fin], my[fin]])
    core = np.zeros(fin.sum(), bool); core[db.core_sample_indices_] = True
    assert np.array_equal(core, key[fin].astype(bool)) and np.array_equal(db.labels_ == -1, cid[fin] == 0)
    assert len(set(zip(db.labels_[core].tolist(), cid[fin][core].tolist()))) == len(set(cid[fin][core].tolist()))


def db_case(name, mx, my, eps, min_pts, cf0):
    cid, key, cls, amount = O.dbscan(mx, my, eps, min_pts, cf0, variant="literal")
    sk_check(mx, my, eps, min_pts, cid, key)
    np.savez_compressed(HERE / f"dbscan_{name}.npz", mx=mx, my=my, eps=eps, min_pts=min_pts, cf0=cf0, cluster_id=cid, is_key=key,
                        is_classed=cls, cluster_amount=amount)
    print(name, len(mx), "pts ->", amount - cf0, "clusters")


def icp_case(name, model, data, e, max_iters):
    order, sq = O.closest_point_set(model, data, "literal")
    d = data.T[:, None, :] - model.T[None, :, :]
    d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
    assert np.array_equal(order, np.argmin(d2, axis=1)) and np.array_equal(sq, d2.min(axis=1))
    R, T, it, sse, olast = O.icp_rigid(model, data, e, max_iters, use_grid=False)
    P = O.trans_points(data, R, T) if it > 1 else data
    np.savez_compressed(HERE / f"icp_{name}.npz", model=model, data=data, e=e, max_iters=max_iters, order0=order, sqdist0=sq, R=R, T=T,
                        iters=it, sse=sse, order_last=olast)
    print(name, model.shape[1], "x", data.shape[1], "->", it, "rounds, sse", sse)


def stats_case(name, mx, my, dist, eps, min_pts, radius):
    """C1-style statistics block: polar -> XYZ, DBSCAN labels, centroids, both circle sets, nearest-truth ids.  Cross-checks: the
    centroids equal NumPy's sequential sums, every circle encloses its members and no smaller circle through two members does."""
    xyz, keep = O.polar_to_xyz(mx, my, dist, 149.0, 307.0)
    assert keep.all()
    cid, key, cls, k = O.dbscan(mx, my, eps, min_pts, 0, variant="literal")
    st = O.cluster_stats(cid, k, xyz, mx, my)
    for c in range(1, k + 1):
        m = np.flatnonzero(cid == c)
        assert st["counts"][c] == len(m) and st["means"][0, c] == np.cumsum(xyz[0, m])[-1] / len(m)
        if st["status3d"][c] == 1:
            cx, cy, r = st["circle3d"][:, c]
            d = np.hypot(xyz[0, m] - cx, xyz[1, m] - cy)
            assert d.max() <= r * (1 + 1e-12)
            pd = np.hypot(xyz[0, m][:, None] - xyz[0, m][None], xyz[1, m][:, None] - xyz[1, m][None])
            assert r >= pd.max() / 2 * (1 - 1e-12)            # at least half the diameter
    tid = np.arange(1, k + 1, dtype=np.int32)
    near = O.nearest_truth_2d(st["means"][3, 1:], st["means"][4, 1:], tid, mx, my, radius)
    d = np.sqrt((st["means"][3, 1:][None] - mx[:, None]) ** 2 + (st["means"][4, 1:][None] - my[:, None]) ** 2)
    want = np.where(d.min(1) < radius, tid[d.shape[1] - 1 - np.argmin(d[:, ::-1], axis=1)], 0)   # ties -> highest index
    assert np.array_equal(near, want)
    np.savez_compressed(HERE / f"stats_{name}.npz", mx=mx, my=my, dist=dist, xyz=xyz, cluster_id=cid, n_clusters=k, radius=radius,
                        means=st["means"], counts=st["counts"], circle3d=st["circle3d"], status3d=st["status3d"],
                        circle2d=st["circle2d"], status2d=st["status2d"], nearest=near)
    print(name, len(mx), "pts ->", k, "clusters,", int((st["status3d"] == 1).sum()), "circles")


if __name__ == "__main__":
    rng = np.random.default_rng(20261018)
    mx, my = synth.dbscan_cloud(0xC1, 14, n_total=10_000, decimals=3)            # config C1
    db_case("c1", mx, my, 0.07, 7, 0)
    lx, ly = rng.integers(0, 50, 3000) * 0.25, rng.integers(0, 50, 3000) * 0.25    # exact ties on the eps boundary
    db_case("lattice_ties", lx.astype(np.float64), ly.astype(np.float64), 0.25, 8, 9)
    bx, by = rng.uniform(0, 1, 800), rng.uniform(0, 1, 800)
    bx[[5, 99]] = np.nan; by[300] = np.inf
    db_case("nonfinite", bx, by, 0.06, 3, 0)
    stats_case("c1", mx, my, np.round(41.7 + 0.42 * synth.uniform(0xC1, 40, np.arange(len(mx), dtype=np.uint64)), 3), 0.07, 7, 0.088)
    g = np.arange(14, dtype=np.float64) * 0.5                                     # C1 ICP: centroids vs checkerboard truth
    truth = np.stack([np.repeat(g, 14), np.tile(g, 14), np.zeros(196)])
    Rz = synth.rotation_about_axis((0, 0, 1), np.radians(3.0))
    data = Rz.T @ (truth + rng.normal(0, 0.003, truth.shape) - np.array([[0.025], [-0.015], [0.01]]))
    icp_case("c1_checkerboard", truth, data, 1e-4, 0)
    model, data, _, _ = synth.icp_clouds(0xC3, 5000, 800)
    icp_case("c3_small", model, data, -1.0, 6)
    lat = np.stack(np.meshgrid(*[np.arange(5.0)] * 3, indexing="ij")).reshape(3, -1)
    icp_case("lattice_ties", np.concatenate([lat, lat[:, ::-1]], axis=1), rng.integers(0, 9, (3, 500)) * 0.5, -1.0, 1)
