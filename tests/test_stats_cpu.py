"""CPU suite: the oracle's restatements of the statistics / matching / ingest steps against independent
NumPy formulations.  The reference ships no fixtures for these either ("parity unpinned"); these checks pin the oracle."""
import itertools

import numpy as np


def _brute_mcc(px, py):
    """Smallest enclosing circle by exhaustive search over all pairs and triples of the points (reference-free)."""
    pts = np.stack([px, py], 1)
    best = np.inf
    for i, j in itertools.combinations(range(len(pts)), 2):
        c = (pts[i] + pts[j]) / 2
        r2 = ((pts[i] - c) ** 2).sum()
        if r2 < best and (((pts - c) ** 2).sum(1) <= r2 * (1 + 1e-12)).all():
            best = r2
    for i, j, k in itertools.combinations(range(len(pts)), 3):
        a, b, c = pts[i], pts[j], pts[k]
        d = 2 * (a[0] * (b[1] - c[1]) + b[0] * (c[1] - a[1]) + c[0] * (a[1] - b[1]))
        if abs(d) < 1e-14:
            continue
        ux = ((a @ a) * (b[1] - c[1]) + (b @ b) * (c[1] - a[1]) + (c @ c) * (a[1] - b[1])) / d
        uy = ((a @ a) * (c[0] - b[0]) + (b @ b) * (a[0] - c[0]) + (c @ c) * (b[0] - a[0])) / d
        ctr = np.array([ux, uy])
        r2 = ((a - ctr) ** 2).sum()
        if r2 < best and (((pts - ctr) ** 2).sum(1) <= r2 * (1 + 1e-9)).all():
            best = r2
    return np.sqrt(best)


def test_cluster_stats_vs_numpy(oracle):
    rng = np.random.default_rng(5)
    n, k = 260, 12
    cid = rng.integers(0, k + 1, n).astype(np.int32)
    cid[cid == 7] = 0                      # an empty cluster
    cid[np.flatnonzero(cid == 9)[3:]] = 0  # a cluster of exactly 3 points: skipped by getCircles (Tools.cs:400)
    xyz = rng.normal(size=(3, n)) * [[3.0], [1.0], [10.0]]
    mx, my = rng.uniform(149, 156, n), rng.uniform(307, 314, n)
    r = oracle.cluster_stats(cid, k, xyz, mx, my)
    for c in range(1, k + 1):
        m = np.flatnonzero(cid == c)
        assert r["counts"][c] == len(m)
        if len(m) == 0:
            assert np.isnan(r["means"][:, c]).all()
            continue
        for f, v in enumerate((xyz[0], xyz[1], xyz[2], mx, my)):
            assert r["means"][f, c] == np.cumsum(v[m])[-1] / len(m)      # sequential sum, one division
        if len(m) <= 3:
            assert r["status3d"][c] == 0 and r["status2d"][c] == 0
            continue
        assert r["status3d"][c] == 1 and r["status2d"][c] == 1
        for circ, hx, hy in ((r["circle3d"], xyz[0], xyz[1]), (r["circle2d"], mx, my)):
            cx, cy, rad = circ[:, c]
            d = np.hypot(hx[m] - cx, hy[m] - cy)
            assert d.max() <= rad * (1 + 1e-12)
            np.testing.assert_allclose(rad, _brute_mcc(hx[m], hy[m]), rtol=1e-9)


def test_circles_degenerate_clusters(oracle):
    # identical points -> hull of one point -> radius 0, centre = points[0]; collinear points -> a two-point circle
    cid = np.array([1] * 5 + [2] * 6, np.int32)
    x = np.array([2.0] * 5 + [0, 1, 2, 3, 4, 5.0])
    y = np.array([3.0] * 5 + [0, 1, 2, 3, 4, 5.0])
    xyz = np.stack([x, y, np.zeros_like(x)])
    r = oracle.cluster_stats(cid, 2, xyz, x, y)
    assert r["status3d"].tolist() == [0, 1, 1]
    assert r["circle3d"][:, 1].tolist() == [2.0, 3.0, 0.0]
    np.testing.assert_allclose(r["circle3d"][:, 2], [2.5, 2.5, np.hypot(2.5, 2.5)], rtol=1e-15)


def test_nearest_truth_2d_vs_numpy(oracle):
    rng = np.random.default_rng(6)
    g = np.arange(8, dtype=np.float64)
    tx, ty = [a.ravel() for a in np.meshgrid(g, g, indexing="ij")]
    tx, ty = np.concatenate([tx, tx]), np.concatenate([ty, ty])       # every truth twice: exact ties
    tid = rng.permutation(len(tx)).astype(np.int32) + 1
    px, py = rng.integers(0, 15, 500) * 0.5, rng.integers(0, 15, 500) * 0.5
    for radius in (0.4, 0.75, 5.0):
        got = oracle.nearest_truth_2d(tx, ty, tid, px, py, radius)
        d = np.sqrt((tx[None] - px[:, None]) * (tx[None] - px[:, None]) + (ty[None] - py[:, None]) * (ty[None] - py[:, None]))
        for i in range(len(px)):
            dm = d[i].min()
            want = tid[np.flatnonzero(d[i] == dm).max()] if dm < radius else 0   # ties -> highest index
            assert got[i] == want


def test_polar_dedupe_parse_vs_numpy(oracle):
    rng = np.random.default_rng(7)
    n = 300
    mx, my = rng.uniform(140, 160, n).round(3), rng.uniform(300, 320, n).round(3)
    ds = rng.uniform(41, 43, n).round(3)
    ds[3], ds[4] = 0.0, 1000.5
    mx[10:20], my[10:20], ds[10:20] = mx[30:40], my[30:40], ds[30:40]   # duplicates of LATER rows: the earlier copy stays
    xyz, keep = oracle.polar_to_xyz(mx, my, ds, 149.0, 307.0)
    ya = -2 * (mx - 149.0) / 180 * np.pi
    fa = 2 * (my - 307.0) / 180 * np.pi
    np.testing.assert_allclose(xyz[0], ds * np.cos(ya) * np.sin(fa), rtol=1e-14, atol=1e-13)
    np.testing.assert_allclose(xyz[1], ds * np.sin(ya) * np.cos(fa), rtol=1e-14, atol=1e-13)
    np.testing.assert_allclose(xyz[2], ds * np.cos(ya), rtol=1e-14)
    assert keep[3] == 0 and keep[4] == 0 and keep.sum() == n - 2
    k2, first, ndup = oracle.dedupe_xyz(xyz, keep)
    seen = {}
    for i in range(n):
        if not keep[i]:
            assert k2[i] == 0 and first[i] == -1
            continue
        key = tuple(xyz[:, i])
        assert first[i] == seen.setdefault(key, i) and k2[i] == (first[i] == i)
    assert ndup == 10 and k2[30:40].sum() == 0 and k2[10:20].sum() == 10
    text = "motor_x\tmotor_y\tDistance\n" + "".join(f"{a:.3f}\t{b:.3f}\t{c:.3f}\r\n" for a, b, c in zip(mx, my, ds))
    pmx, pmy, pds, st = oracle.parse_rows(text.encode())
    assert (st == 0).all()
    np.testing.assert_array_equal(pmx, mx)
    np.testing.assert_array_equal(pmy, my)
    np.testing.assert_array_equal(pds, ds)
