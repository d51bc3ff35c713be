"""GPU parity: sort / cluster statistics / nearest truth / ingest through the C ABI vs the CPU oracle.
Integer and index results bit-exact; centroids, circle centres and radii bit-exact too (same operation order, no FMA);
only the sin/cos stage of the polar conversion is compared to a tolerance."""
import numpy as np
import pytest
import torch

from vtkcloudpoint_b200 import synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _eq(a, b):
    np.testing.assert_array_equal(np.asarray(a), np.asarray(b))


def test_sort_pairs_is_stable(ctx):
    rng = np.random.default_rng(11)
    for n, hi in ((1, 4), (257, 3), (5000, 70000), (300_000, 1 << 40), (1_000_003, 1000)):
        keys = rng.integers(0, hi, n).astype(np.int64)
        dk = torch.from_numpy(keys.copy()).to(DEV)
        bits = max(8, int(np.ceil(np.log2(hi) / 8)) * 8)
        sk, perm = ctx.sort_pairs_dev(dk, None, 0, bits)
        want = np.argsort(keys, kind="stable")
        _eq(perm.cpu().numpy(), want)
        _eq(sk.cpu().numpy(), keys[want])


def test_argsort_f64_compare_to_order(ctx):
    rng = np.random.default_rng(12)
    v = rng.normal(size=100_000)
    v[::7] = np.round(v[::7], 1)                      # many ties
    v[5], v[6], v[100], v[101], v[102] = np.nan, -np.inf, 0.0, -0.0, np.inf
    order = ctx.argsort_f64_dev(torch.from_numpy(v).to(DEV)).cpu().numpy()
    key = np.where(np.isnan(v), -np.inf, v)           # Double.CompareTo: NaN first; then stable ascending, -0.0 == 0.0
    nan_first = np.lexsort((np.arange(len(v)), key, ~np.isnan(v)))
    _eq(order, nan_first)


def _random_clusters(rng, n, k):
    cid = rng.integers(0, k + 1, n).astype(np.int32)
    xyz = rng.normal(size=(3, n))
    centres = rng.uniform(-50, 50, (2, k + 1))
    xyz[0] = xyz[0] * 0.4 + centres[0, cid]
    xyz[1] = xyz[1] * 0.2 + centres[1, cid]
    mx = 149 + (xyz[0] + 50) * 0.07 + rng.normal(size=n) * 1e-3
    my = 307 + (xyz[1] + 50) * 0.07 + rng.normal(size=n) * 1e-3
    return cid, xyz, mx, my


def test_cluster_stats_bit_exact(ctx, oracle):
    rng = np.random.default_rng(13)
    for n, k in ((0, 3), (50, 4), (4000, 60), (60_000, 1500)):
        cid, xyz, mx, my = _random_clusters(rng, n, k)
        if n:
            cid[cid == 2] = 0                          # empty cluster
            cid[np.flatnonzero(cid == 3)[3:]] = 0      # exactly three members: skipped
        got = ctx.cluster_stats(cid, k, xyz, mx, my)
        want = oracle.cluster_stats(cid, k, xyz, mx, my)
        for key in ("counts", "status3d", "status2d"):
            _eq(got[key], want[key])
        for key in ("means", "circle3d", "circle2d"):
            _eq(got[key][:, 1:].view(np.int64), want[key][:, 1:].view(np.int64))     # bit patterns (NaN-safe)


def test_circles_ties_duplicates_collinear(ctx, oracle):
    # lattice points (equal angles, equal radii, collinear triples, duplicates): every tie rule of the gift wrap and of
    # the circle search is exercised; results must still match bit for bit
    rng = np.random.default_rng(14)
    k = 300
    n = 9000
    cid = rng.integers(1, k + 1, n).astype(np.int32)
    x = rng.integers(0, 6, n).astype(np.float64) + 10 * (cid % 17)
    y = rng.integers(0, 6, n).astype(np.float64) + 10 * (cid // 17)
    cid[:40] = 1; x[:40] = 3.0; y[:40] = 4.0                        # a cluster of identical points (plus lattice points)
    line = np.flatnonzero(cid == 5)
    y[line] = x[line]                                               # a collinear cluster
    xyz = np.stack([x, y, rng.normal(size=n)])
    got = ctx.cluster_stats(cid, k, xyz, y, x)
    want = oracle.cluster_stats(cid, k, xyz, y, x)
    _eq(got["status3d"], want["status3d"]); _eq(got["status2d"], want["status2d"])
    _eq(got["circle3d"][:, 1:].view(np.int64), want["circle3d"][:, 1:].view(np.int64))
    _eq(got["circle2d"][:, 1:].view(np.int64), want["circle2d"][:, 1:].view(np.int64))


def test_circles_nonfinite_status(ctx, oracle):
    cid = np.array([1] * 5 + [2] * 5 + [3] * 4, np.int32)
    x = np.arange(14, dtype=np.float64); y = x * x
    x[6] = np.nan                                   # cluster 2: one NaN coordinate -> -2
    x[10:] = np.nan; y[10:] = np.nan                # cluster 3: everything culled -> the C# throws -> -1
    xyz = np.stack([x, y, x * 0])
    got = ctx.cluster_stats(cid, 3, xyz, x, y)
    want = oracle.cluster_stats(cid, 3, xyz, x, y)
    _eq(got["status3d"], want["status3d"])
    assert got["status3d"].tolist() == [0, 1, -2, -1]


def test_groups_means_circles_device_forms(ctx, oracle):
    rng = np.random.default_rng(15)
    n, k = 200_000, 5000
    cid, xyz, mx, my = _random_clusters(rng, n, k)
    d_cid = torch.from_numpy(cid).to(DEV)
    members, offsets = ctx.cluster_groups_dev(d_cid, k)
    order = np.argsort(cid, kind="stable")
    _eq(members.cpu().numpy(), order)
    _eq(offsets.cpu().numpy(), np.searchsorted(cid[order], np.arange(k + 2)))
    vals = torch.from_numpy(np.stack([xyz[0], xyz[1], xyz[2], mx, my])).to(DEV)
    means, counts = ctx.cluster_means_ordered_dev(members, offsets, k, vals)
    circ, status = ctx.cluster_circles_dev(members, offsets, k, vals[0].contiguous(), vals[1].contiguous())
    flag = ctx.radius_filter_dev(circ[2].contiguous(), status, k, 0.9)
    want = oracle.cluster_stats(cid, k, xyz, mx, my, circles2d=False)
    _eq(means.cpu().numpy()[:, 1:].view(np.int64), want["means"][:, 1:].view(np.int64))
    _eq(counts.cpu().numpy(), want["counts"])
    _eq(circ.cpu().numpy()[:, 1:].view(np.int64), want["circle3d"][:, 1:].view(np.int64))
    # MainForm.FilterClustersByRadius: strict '>' on the 3-D circle's radius (FrmMain.cs:1910)
    _eq(flag.cpu().numpy()[1:], ((want["status3d"][1:] == 1) & (want["circle3d"][2, 1:] > 0.9)).astype(np.uint8))
    # the atomics-based means agree with the ordered ones to rounding
    m2, c2 = ctx.cluster_means_dev(d_cid, k, vals)
    np.testing.assert_allclose(m2.cpu().numpy()[:, 1:], want["means"][:, 1:], rtol=1e-12)


def test_nearest_truth_2d_ties_highest_index(ctx, oracle):
    rng = np.random.default_rng(16)
    g = np.arange(12, dtype=np.float64) * 0.5
    tx, ty = [a.ravel() for a in np.meshgrid(g, g, indexing="ij")]
    tx, ty = np.concatenate([tx, tx[::-1]]), np.concatenate([ty, ty[::-1]])      # duplicates in reversed order
    tid = (rng.permutation(len(tx)) + 1).astype(np.int32)
    tid[5] = 0                                                                   # a truth whose clusterId is 0
    px, py = rng.integers(-4, 28, 20000) * 0.25, rng.integers(-4, 28, 20000) * 0.25
    px[7], py[8] = np.nan, np.inf
    for radius in (0.2, 0.25, 0.36, 3.0, np.inf, np.nan, -1.0):
        _eq(ctx.nearest_truth_2d(tx, ty, tid, px, py, radius), oracle.nearest_truth_2d(tx, ty, tid, px, py, radius))
    _eq(ctx.nearest_truth_2d(tx, ty, None, px, py, 0.3), oracle.nearest_truth_2d(tx, ty, None, px, py, 0.3))


def test_nearest_truth_2d_c1_like(ctx, oracle):
    # the checkerboard of config C1: 196 truths at pitch 0.5, scan points around them
    mx, my = synth.dbscan_cloud(0xC1, 14, n_total=10_000)
    cx, cy = [a.ravel() for a in np.meshgrid(149.0 + 0.5 * np.arange(14), 307.0 + 0.5 * np.arange(14), indexing="ij")]
    tid = np.arange(1, 197, dtype=np.int32)
    for radius in (0.05, 0.088, 0.3):
        _eq(ctx.nearest_truth_2d(cx, cy, tid, mx, my, radius), oracle.nearest_truth_2d(cx, cy, tid, mx, my, radius))


def test_polar_and_dedupe(ctx, oracle):
    rng = np.random.default_rng(17)
    n = 50_000
    mx, my = rng.uniform(140, 160, n).round(3), rng.uniform(300, 320, n).round(3)
    ds = rng.uniform(41, 43, n).round(3)
    ds[::97] = 0.0; ds[5::211] = 1000.001; ds[11] = np.nan; ds[12] = 1000.0
    src = rng.integers(0, n, 4000); dst = rng.integers(0, n, 4000)
    mx[dst], my[dst], ds[dst] = mx[src], my[src], ds[src]            # exact duplicates scattered through the file
    d = [torch.from_numpy(a).to(DEV) for a in (mx, my, ds)]
    for xdir, ydir in ((2, 1), (1, 2), (3, 4)):
        xyz, keep = ctx.polar_to_xyz_dev(*d, 149.0, 307.0, xdir, ydir)
        oxyz, okeep = oracle.polar_to_xyz(mx, my, ds, 149.0, 307.0, xdir, ydir)
        _eq(keep.cpu().numpy(), okeep)
        np.testing.assert_allclose(xyz.cpu().numpy(), oxyz, rtol=1e-12, atol=1e-12)     # sin/cos differ by ulps
    # duplicate removal is exact on whatever coordinates it is given: feed both sides the oracle's XYZ
    oxyz, okeep = oracle.polar_to_xyz(mx, my, ds, 149.0, 307.0)
    oxyz[0, 20], oxyz[0, 21] = 0.0, -0.0                             # -0.0 == 0.0
    oxyz[1, 20] = oxyz[1, 21]; oxyz[2, 20] = oxyz[2, 21]
    oxyz[:, 30] = np.nan; oxyz[:, 31] = np.nan                       # NaN equals nothing: both rows stay
    k, f, nd = ctx.dedupe_xyz_dev(torch.from_numpy(oxyz).to(DEV), torch.from_numpy(okeep).to(DEV))
    wk, wf, wnd = oracle.dedupe_xyz(oxyz[:, :6000].copy(), okeep[:6000])
    k6, f6, nd6 = ctx.dedupe_xyz_dev(torch.from_numpy(oxyz[:, :6000].copy()).to(DEV), torch.from_numpy(okeep[:6000].copy()).to(DEV))
    _eq(k6.cpu().numpy(), wk); _eq(f6.cpu().numpy(), wf); assert int(nd6.item()) == wnd
    # full size: first-occurrence property checked with a dictionary
    k, f = k.cpu().numpy(), f.cpu().numpy()
    seen = {}
    for i in range(n):
        if not okeep[i]:
            assert k[i] == 0 and f[i] == -1
            continue
        key = tuple(oxyz[:, i] + 0.0)
        if any(np.isnan(key)):
            assert k[i] == 1 and f[i] == i
            continue
        first = seen.setdefault(key, i)
        assert f[i] == first and k[i] == (first == i)
    assert int(nd.item()) == int(okeep.sum() - k.sum())


def test_ingest_text(ctx, oracle):
    rng = np.random.default_rng(18)
    n = 30_000
    mx, my = rng.uniform(140, 160, n).round(3), rng.uniform(300, 320, n).round(3)
    ds = rng.uniform(41, 43, n).round(3)
    ds[::101] = 0.0
    mx[100:200], my[100:200], ds[100:200] = mx[300:400], my[300:400], ds[300:400]
    rows = [f"{a:.3f}\t{b:.3f}\t{c:.3f}" for a, b, c in zip(mx, my, ds)]
    rows[7] = "1.5e1\t-2.50\t+42"               # exponent / sign forms Convert.ToDouble accepts
    rows[8] = "abc\t1\t2"                       # FormatException in the C#
    rows[9] = "1.0\t2.0"                        # too few fields
    rows[10] = " 150.25 \t 310.5\t41.125\t99"   # padding and an extra field
    rows[11] = "0.1234567890123456789012\t1\t42"  # more digits than the exact path handles -> status 2
    for eol, tail in (("\n", "\n"), ("\r\n", ""), ("\n", "")):
        text = ("motor_x\tmotor_y\tDistance" + eol + eol.join(rows) + tail).encode()
        got = ctx.ingest_text(text, 149.0, 307.0, remove_duplicates=True)
        pmx, pmy, pds, st = oracle.parse_rows(text)
        assert len(got["mx"]) == n
        want_st = st.copy(); want_st[11] = 2
        _eq(got["row_status"], want_st)
        ok = want_st == 0
        _eq(got["mx"][ok], pmx[ok]); _eq(got["my"][ok], pmy[ok]); _eq(got["dist"][ok], pds[ok])
        oxyz, okeep = oracle.polar_to_xyz(got["mx"], got["my"], got["dist"], 149.0, 307.0)
        np.testing.assert_allclose(got["xyz"], oxyz, rtol=1e-12, atol=1e-12)
        wk, _, wnd = oracle.dedupe_xyz(got["xyz"], okeep)      # de-dup is exact given the XYZ the library produced
        _eq(got["keep"], wk)
        assert got["n_duplicates"] == wnd and got["n_kept"] == int(wk.sum())
    empty = ctx.ingest_text(b"", 149.0, 307.0)
    assert len(empty["mx"]) == 0
    hdr = ctx.ingest_text(b"motor_x\tmotor_y\tDistance\n", 149.0, 307.0)
    assert len(hdr["mx"]) == 0


def test_circles_large_clusters(ctx, oracle):
    # clusters of thousands of points with hulls of 100+ vertices (points on and inside circles / ellipses): the pair and
    # triple loops stride past one warp, the gift wrap runs for hundreds of rounds
    rng = np.random.default_rng(19)
    xs, ys, cs = [], [], []
    for c, (m, ring) in enumerate(((3000, 150), (800, 40), (5000, 0), (64, 64), (33, 33)), start=1):
        ang = np.sort(rng.uniform(0, 2 * np.pi, ring))
        rx, ry = 3.0 * np.cos(ang) + 20 * c, 2.0 * np.sin(ang) - 7 * c
        r = np.sqrt(rng.uniform(0, 0.97, m - ring))
        th = rng.uniform(0, 2 * np.pi, m - ring)
        xs.append(np.concatenate([rx, 3.0 * r * np.cos(th) + 20 * c])); ys.append(np.concatenate([ry, 2.0 * r * np.sin(th) - 7 * c]))
        cs.append(np.full(m, c, np.int32))
    x, y, cid = np.concatenate(xs), np.concatenate(ys), np.concatenate(cs)
    perm = rng.permutation(len(x))
    x, y, cid = x[perm], y[perm], cid[perm]
    xyz = np.stack([x, y, np.zeros_like(x)])
    got = ctx.cluster_stats(cid, 5, xyz, y, x)
    want = oracle.cluster_stats(cid, 5, xyz, y, x)
    _eq(got["status3d"], want["status3d"]); _eq(got["status2d"], want["status2d"])
    _eq(got["circle3d"][:, 1:].view(np.int64), want["circle3d"][:, 1:].view(np.int64))
    _eq(got["circle2d"][:, 1:].view(np.int64), want["circle2d"][:, 1:].view(np.int64))
    _eq(got["means"][:, 1:].view(np.int64), want["means"][:, 1:].view(np.int64))


def test_nearest_truth_more_truths_than_points(ctx, oracle):
    rng = np.random.default_rng(20)
    tx, ty = rng.uniform(0, 50, 40_000), rng.uniform(0, 50, 40_000)
    tid = rng.integers(0, 5000, 40_000).astype(np.int32)
    px, py = rng.uniform(-5, 55, 3000), rng.uniform(-5, 55, 3000)
    for radius in (0.05, 0.3, 100.0):
        _eq(ctx.nearest_truth_2d(tx, ty, tid, px, py, radius), oracle.nearest_truth_2d(tx, ty, tid, px, py, radius))
    # a single truth, and truths that are all non-finite
    _eq(ctx.nearest_truth_2d([1.0], [2.0], [9], px, py, 60.0), oracle.nearest_truth_2d([1.0], [2.0], [9], px, py, 60.0))
    nan = np.full(5, np.nan)
    _eq(ctx.nearest_truth_2d(nan, nan, None, px, py, 60.0), np.zeros(len(px), np.int32))
