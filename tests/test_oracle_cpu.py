"""CPU suite: the oracle against independent implementations (scikit-learn, NumPy) and against itself
(literal vs grid variants).  The reference ships no golden vectors, so these
cross-checks plus tests/golden are what pins the oracle."""
import numpy as np
import pytest

from vtkcloudpoint_b200 import synth


def _sk_check(mx, my, eps, min_pts, cid, key):
    from sklearn.cluster import DBSCAN
    db = DBSCAN(eps=eps, min_samples=min_pts, metric="manhattan", algorithm="brute").fit(np.c_[mx, my])
    core = np.zeros(len(mx), bool)
    core[db.core_sample_indices_] = True
    np.testing.assert_array_equal(core, key.astype(bool))
    np.testing.assert_array_equal(db.labels_ == -1, cid == 0)
    # core points: same partition
    a, b = db.labels_[core], cid[core]
    pairs = set(zip(a.tolist(), b.tolist()))
    assert len(pairs) == len(set(a.tolist())) == len(set(b.tolist()))
    # cluster numbering: ascending minimum core index (DBImproved.cs:93-110)
    firsts = [np.flatnonzero(core & (cid == c))[0] for c in sorted(set(b.tolist()))]
    assert firsts == sorted(firsts)
    # border rule: max adjacent cluster id (DBImproved.cs:87)
    for i in np.flatnonzero(~core & (cid != 0)):
        d = np.abs(mx[i] - mx) + np.abs(my[i] - my)
        assert cid[i] == cid[core & (d <= eps)].max()


@pytest.mark.parametrize("seed", range(8))
def test_dbscan_literal_vs_sklearn_and_grid(oracle, seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(50, 1500))
    k = int(rng.integers(1, 6))
    centres = rng.uniform(0, 1, (k, 2))
    pts = np.vstack([centres[rng.integers(0, k, n // 2)] + rng.normal(0, 0.03, (n // 2, 2)), rng.uniform(0, 1, (n - n // 2, 2))])
    mx, my = pts[:, 0].copy(), pts[:, 1].copy()
    eps, min_pts = float(rng.uniform(0.02, 0.1)), int(rng.integers(2, 8))
    cid, key, cls, amount = oracle.dbscan(mx, my, eps, min_pts, variant="literal")
    _sk_check(mx, my, eps, min_pts, cid, key)
    np.testing.assert_array_equal(cls, (cid != 0).astype(np.uint8))
    assert amount == len(set(cid[cid != 0].tolist()))
    for kw in (dict(variant="literal", dedup_loop=True), dict(variant="grid"), dict(variant="grid", n_threads=4)):
        c2, k2, s2, a2 = oracle.dbscan(mx, my, eps, min_pts, **kw)
        np.testing.assert_array_equal(c2, cid)
        np.testing.assert_array_equal(k2, key)
        np.testing.assert_array_equal(s2, cls)
        assert a2 == amount


def test_dbscan_c1_config(oracle):
    mx, my = synth.dbscan_cloud(0xC1, 14, n_total=10_000, decimals=3)
    cid, key, cls, amount = oracle.dbscan(mx, my, 0.07, 7, variant="literal")
    assert amount == 196
    _sk_check(mx, my, 0.07, 7, cid, key)
    c2, k2, s2, a2 = oracle.dbscan(mx, my, 0.07, 7, variant="grid")
    np.testing.assert_array_equal(c2, cid)


def test_dbscan_seeded_cf_and_edge_cases(oracle):
    rng = np.random.default_rng(3)
    mx, my = rng.uniform(0, 1, 300), rng.uniform(0, 1, 300)
    base = oracle.dbscan(mx, my, 0.08, 3, 0, variant="literal")
    seeded = oracle.dbscan(mx, my, 0.08, 3, 41, variant="literal")
    np.testing.assert_array_equal(np.where(base[0] > 0, base[0] + 41, 0), seeded[0])
    assert seeded[3] == base[3] + 41
    mx[7] = np.nan
    for eps, min_pts in ((0.08, 3), (0.08, 1), (0.08, 0), (-1.0, 2), (-1.0, 0), (float("nan"), 1), (0.0, 1)):
        lit = oracle.dbscan(mx, my, eps, min_pts, 5, variant="literal")
        grd = oracle.dbscan(mx, my, eps, min_pts, 5, variant="grid")
        for a, b in zip(lit, grd):
            np.testing.assert_array_equal(a, b)
    lit = oracle.dbscan(mx, my, 0.08, 3, variant="literal")
    assert lit[0][7] == 0 and lit[1][7] == 0          # NaN point: no neighbours, not even itself
    lit0 = oracle.dbscan(mx, my, 0.08, 0, variant="literal")
    assert lit0[1][7] == 1 and lit0[0][7] != 0 and lit0[2][7] == 0   # minPts <= 0: own cluster, isClassed stays false
    e = np.empty(0)
    assert oracle.dbscan(e, e, 0.1, 3, 9, variant="literal")[3] == 9


@pytest.mark.parametrize("seed", range(4))
def test_closest_point_set(oracle, seed):
    rng = np.random.default_rng(seed)
    m, n = int(rng.integers(1, 3000)), int(rng.integers(1, 500))
    model = rng.uniform(-3, 3, (3, m))
    data = rng.uniform(-4, 4, (3, n))
    if seed == 3:
        model = np.round(model)            # heavy ties
        data = np.round(data * 2) / 2
    o1, s1 = oracle.closest_point_set(model, data, "literal")
    d = data.T[:, None, :] - model.T[None, :, :]
    d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
    np.testing.assert_array_equal(o1, np.argmin(d2, axis=1))     # argmin returns the first minimum = lowest index
    np.testing.assert_array_equal(s1, d2.min(axis=1))
    for kw in (dict(variant="grid"), dict(variant="grid", n_threads=3), dict(variant="literal", n_threads=3)):
        o2, s2 = oracle.closest_point_set(model, data, **kw)
        np.testing.assert_array_equal(o2, o1)
        np.testing.assert_array_equal(s2, s1)


def test_jacobi_reference_matrix(oracle):
    # the only fixed input in the reference: FrmMain.cs:2637-2640 (no expected output recorded there)
    a = np.array([[1.0, 0, 2], [0, 2, 0], [2, 0, 3]])
    ok, w, v, _ = oracle.jacobi_eig(a, 100, 1e-4)
    assert ok
    np.testing.assert_allclose(np.sort(w), [2 - 5 ** 0.5, 2.0, 2 + 5 ** 0.5], atol=1e-7)
    ok, w, v, _ = oracle.jacobi_eig(a, 100, 1e-14)
    np.testing.assert_allclose(a @ v, v * w, atol=1e-12)
    rng = np.random.default_rng(0)
    for _ in range(20):
        q = rng.normal(size=(4, 4)); q = q + q.T
        ok, w, v, _ = oracle.jacobi_eig(q, 100, 1e-14)
        assert ok
        np.testing.assert_allclose(np.sort(w), np.linalg.eigvalsh(q), atol=1e-12)
        np.testing.assert_allclose(q @ v, v * w, atol=1e-11)


def test_rigid_step_vs_kabsch(oracle):
    rng = np.random.default_rng(1)
    for _ in range(10):
        n = int(rng.integers(3, 400))
        P = rng.normal(size=(3, n)) * rng.uniform(0.1, 50)
        R = synth.rotation_about_axis(rng.normal(size=3), rng.uniform(-3, 3))
        t = rng.normal(size=3) * 10
        Y = R @ P + t[:, None] + rng.normal(size=(3, n)) * 1e-3
        R1, T1, sse = oracle.rigid_step(P, Y)
        Pc, Yc = P - P.mean(1, keepdims=True), Y - Y.mean(1, keepdims=True)
        U, S, Vt = np.linalg.svd(Yc @ Pc.T)
        D = np.diag([1, 1, np.sign(np.linalg.det(U @ Vt))])
        Rk = U @ D @ Vt
        np.testing.assert_allclose(R1, Rk, atol=1e-9)
        np.testing.assert_allclose(T1, Y.mean(1) - Rk @ P.mean(1), atol=1e-8)
        np.testing.assert_allclose(sse, ((P - Y) ** 2).sum(), rtol=1e-12)
        np.testing.assert_allclose(R1 @ R1.T, np.eye(3), atol=1e-12)


def test_icp_recovers_transform_and_control_flow(oracle):
    model, data, R, T = synth.icp_clouds(0xC3, 20_000, 2_000)
    Rg, Tg, it_g, sse_g, ord_g = oracle.icp_rigid(model, data, 1e-4, 0, use_grid=True)
    Rl, Tl, it_l, sse_l, ord_l = oracle.icp_rigid(model, data, 1e-4, 0, use_grid=False)
    assert it_g == it_l and sse_g == sse_l
    np.testing.assert_array_equal(ord_g, ord_l)
    np.testing.assert_array_equal(Rg, Rl)
    assert np.abs(Rg - R).max() < 1e-4 and np.abs(Tg - T).max() < 5e-3
    assert (ord_g == np.arange(2000)).all()
    # capped run: exactly max_iters rounds; e < 0 disables the convergence test
    _, _, it, _, _ = oracle.icp_rigid(model, data, -1.0, 3)
    assert it == 3
    # converged on round 1: R, T untouched (ICP.cs:149-180)
    R0, T0 = np.arange(9.0), np.array([1.0, 2, 3])
    Rr, Tr, it, sse, _ = oracle.icp_rigid(model[:, :10], model[:, :10], 1e-4, 0, R0, T0)
    assert it == 1 and sse == 0.0
    np.testing.assert_array_equal(Rr.ravel(), R0)
    np.testing.assert_array_equal(Tr, T0)


def test_trans_points_order_of_operations(oracle):
    rng = np.random.default_rng(2)
    src = rng.normal(size=(3, 100))
    R = rng.normal(size=(3, 3)); T = rng.normal(size=3)
    dst = oracle.trans_points(src, R, T)
    exp = np.stack([((0.0 + R[i, 0] * src[0]) + R[i, 1] * src[1]) + R[i, 2] * src[2] + T[i] for i in range(3)])
    np.testing.assert_array_equal(dst, exp)


# ---- the C# exactly as written: its defects as executable facts (oracle/vpc_oracle_aswritten.cpp) -------------------------
def test_jacobi_as_written_is_wrong_on_the_references_own_test_matrix(oracle):
    # the one fixed input in the reference (FrmMain.cs:2637-2640): true eigenvalues 2 and 2 +- sqrt(5)
    a = np.array([[1, 0, 2], [0, 2, 0], [2, 0, 3.0]])
    ok, w, v, _ = oracle.jacobi_eig_as_written(a, 100, 1e-4)
    true = np.array([2 - np.sqrt(5), 2, 2 + np.sqrt(5)])
    assert ok                                              # it even reports success ...
    assert np.abs(np.sort(w) - true).max() > 1.0           # ... with eigenvalues that are off by more than 1
    assert abs(np.linalg.det(v)) < 0.5                     # and "eigenvectors" that are not even orthonormal
    ok2, w2, v2, _ = oracle.jacobi_eig(a, 100, 1e-12)      # the restored routine (indices put back) is right
    assert ok2 and np.abs(np.sort(w2) - true).max() < 1e-9


def test_icp_as_written_throws_on_round_two_and_is_not_rigid(oracle):
    rng = np.random.default_rng(0)
    model = rng.uniform(0, 10, (3, 200))
    th = 0.05
    Rz = np.array([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1]])
    data = Rz.T @ (model - np.array([[1.0], [2.0], [0.5]]))
    r = oracle.icp_as_written(model, data, 1e-4)
    assert r["rc"] == -7 and r["rounds"] == 2              # IndexOutOfRangeException in the composition loop (ICP.cs:170-174)
    assert np.abs(r["R"] @ r["R"].T - np.eye(3)).max() > 0.1   # what it had written to R is not a rotation
    Ro, To, it, sse, _ = oracle.icp_rigid(model, data, 1e-4, 50, use_grid=False)
    assert np.abs(Ro - Rz).max() < 1e-9 and np.abs(To - [1.0, 2.0, 0.5]).max() < 1e-9 and sse < 1e-12   # the intended algorithm


def test_icp_as_written_integer_division_zeroes_the_cross_covariance(oracle):
    # one point: 1 / P.Count == 1, more points: == 0 -- with n = 1 and model = data the loop converges in round 2 (d == pre_d == 0)
    p = np.array([[1.0], [2.0], [3.0]])
    r = oracle.icp_as_written(p, p, 1e-4)
    assert r["rc"] == 0 and r["rounds"] == 1 and r["sse"][0] == 0.0
