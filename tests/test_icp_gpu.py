"""GPU parity: vpc_closest_point_set / vpc_icp_rigid (through the C ABI) vs the CPU oracle.
Correspondences index-exact; R, T, SSE within 1e-6 relative (north_star tolerance)."""
import numpy as np
import pytest

from vtkcloudpoint_b200 import synth

pytestmark = pytest.mark.gpu
RTOL = 1e-6


def test_closest_random(ctx, oracle):
    rng = np.random.default_rng(1)
    for m, n in ((1, 5), (2, 7), (300, 200), (5000, 3000)):
        model = rng.uniform(-5, 5, (3, m))
        data = rng.uniform(-6, 6, (3, n))
        order, sq = ctx.closest_point_set(model, data)
        o, s = oracle.closest_point_set(model, data, "literal")
        np.testing.assert_array_equal(order, o)
        np.testing.assert_array_equal(sq, s)


def test_closest_ties_lowest_index(ctx, oracle):
    # integer lattice model with duplicates, queries at lattice midpoints: massive exact ties
    g = np.arange(6, dtype=np.float64)
    model = np.stack(np.meshgrid(g, g, g, indexing="ij")).reshape(3, -1)
    model = np.concatenate([model, model[:, ::-1]], axis=1)          # every point twice
    rng = np.random.default_rng(2)
    data = rng.integers(0, 11, (3, 4000)) * 0.5
    order, sq = ctx.closest_point_set(model, data)
    o, s = oracle.closest_point_set(model, data, "literal")
    np.testing.assert_array_equal(order, o)
    np.testing.assert_array_equal(sq, s)


def test_closest_planar_and_far_queries(ctx, oracle):
    rng = np.random.default_rng(3)
    model = rng.uniform(0, 100, (3, 2000))
    model[2] = 0.0                                                   # production ICP feeds z = 0 (Tools.cs:701)
    data = rng.uniform(-300, 400, (3, 1500))
    order, sq = ctx.closest_point_set(model, data)
    o, s = oracle.closest_point_set(model, data, "literal")
    np.testing.assert_array_equal(order, o)
    np.testing.assert_array_equal(sq, s)


def test_closest_nonfinite(ctx, oracle):
    rng = np.random.default_rng(4)
    model = rng.uniform(0, 10, (3, 400))
    data = rng.uniform(0, 10, (3, 300))
    data[0, 5] = np.nan
    data[1, 9] = np.inf
    model[2, 17] = np.nan
    model[0, 33] = -np.inf
    for variant_model in (model, np.concatenate([np.full((3, 1), np.nan), model], axis=1),
                          np.concatenate([np.array([[np.inf], [0.0], [0.0]]), model], axis=1)):
        order, sq = ctx.closest_point_set(variant_model, data)
        o, s = oracle.closest_point_set(variant_model, data, "literal")
        np.testing.assert_array_equal(order, o)
        np.testing.assert_array_equal(sq, s)


def test_closest_bad_args(ctx):
    from vtkcloudpoint_b200 import VpcError
    with pytest.raises(VpcError):
        ctx.closest_point_set(np.empty((3, 0)), np.zeros((3, 4)))


def _rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300)


def test_icp_c1_like(ctx, oracle):
    # ~196 centroids against the checkerboard truth (config C1), e = 1e-4 (FrmMain.cs:2689)
    g = np.arange(14, dtype=np.float64) * 0.5
    truth = np.stack([np.repeat(g, 14), np.tile(g, 14), np.zeros(196)])
    rng = np.random.default_rng(5)
    R = synth.rotation_about_axis((0, 0, 1), np.radians(3.0))
    data = R.T @ (truth + rng.normal(0, 0.003, truth.shape) - np.array([[0.025], [-0.015], [0.01]]))
    res = ctx.icp_rigid(truth, data, 1e-4)
    Ro, To, it, sse, order = oracle.icp_rigid(truth, data, 1e-4, use_grid=False)
    assert res.iters_done == it
    np.testing.assert_array_equal(res.order_last, order)
    assert _rel(res.R, Ro) < RTOL and _rel(res.T, To) < RTOL and abs(res.sse_last - sse) <= RTOL * sse


def test_icp_fixed_iterations(ctx, oracle):
    model, data, R, T = synth.icp_clouds(0xC3, 50_000, 5_000)
    for iters in (1, 2, 7):
        res = ctx.icp_rigid(model, data, -1.0, iters)
        Ro, To, it, sse, order = oracle.icp_rigid(model, data, -1.0, iters)
        assert res.iters_done == it == iters
        np.testing.assert_array_equal(res.order_last, order)
        assert _rel(res.R, Ro) < RTOL and _rel(res.T, To) < RTOL and abs(res.sse_last - sse) <= RTOL * sse
    assert _rel(res.R, R) < 1e-3


def test_icp_round1_convergence_keeps_rt(ctx, oracle):
    # |d - 0| < e on round 1: R and T are left exactly as the caller passed them (ICP.cs:149-180)
    model = np.array([[0.0, 1, 0], [0, 0, 1], [0, 0, 0]])
    data = model.copy()
    R0 = np.arange(9.0)
    T0 = np.array([7.0, 8, 9])
    res = ctx.icp_rigid(model, data, 1e-4, 0, R0, T0)
    Ro, To, it, sse, order = oracle.icp_rigid(model, data, 1e-4, 0, R0, T0)
    assert res.iters_done == it == 1
    np.testing.assert_array_equal(res.R, Ro)
    np.testing.assert_array_equal(res.T, To)


def test_icp_device_entry(ctx, oracle):
    import torch
    model, data, R, T = synth.icp_clouds(0xC3, 100_000, 10_000)
    tm = torch.from_numpy(model).cuda()
    td = torch.from_numpy(data).cuda()
    ctx.icp_set_model_dev(tm)
    order, sq = ctx.closest_point_set_dev(td)
    o, s = oracle.closest_point_set(model, data, "grid")
    np.testing.assert_array_equal(order.cpu().numpy(), o)
    np.testing.assert_array_equal(sq.cpu().numpy(), s)
    state, order_last = ctx.icp_rigid_dev(td, -1.0, 10)
    torch.cuda.synchronize()
    st = state.cpu().numpy()
    Ro, To, it, sse, oo = oracle.icp_rigid(model, data, -1.0, 10)
    assert int(st[13]) == 10
    np.testing.assert_array_equal(order_last.cpu().numpy(), oo)
    assert _rel(st[:9].reshape(3, 3), Ro) < RTOL and _rel(st[9:12], To) < RTOL and abs(st[12] - sse) <= RTOL * sse


def test_icp_shard_steps_single_rank(ctx, oracle):
    # the vpc_icp_shard_* exports through the distributed driver at world size 1
    import torch
    from vtkcloudpoint_b200.distributed import GpuIcpBackend, icp_rigid_sharded
    model, data, R, T = synth.icp_clouds(0xC3, 60_000, 6_000)
    state, order = icp_rigid_sharded(GpuIcpBackend(ctx), torch.from_numpy(model).cuda(), 0, torch.from_numpy(data).cuda(), -1.0, 6)
    torch.cuda.synchronize()
    st = state.cpu().numpy()
    Ro, To, it, sse, oo = oracle.icp_rigid(model, data, -1.0, 6)
    assert int(st[13]) == it == 6
    np.testing.assert_array_equal(order.cpu().numpy(), oo)
    assert _rel(st[:9].reshape(3, 3), Ro) < RTOL and _rel(st[9:12], To) < RTOL and abs(st[12] - sse) <= RTOL * sse


def test_match_within(ctx, oracle):
    # MainForm.RecorrectMatchingPtsByDistance (FrmMain.cs:3588-3618): sqrt-ranked nearest truth point + threshold
    rng = np.random.default_rng(9)
    for m, n, thr in ((196, 196, 0.05), (1, 10, 1.0), (3000, 2000, 0.3)):
        truth = rng.uniform(0, 10, (3, m))
        cen = truth[:, rng.integers(0, m, n)] + rng.normal(0, 0.04, (3, n))
        mid, dist = ctx.match_within(truth, cen, thr)
        omid, odist = oracle.match_within(truth, cen, thr)
        np.testing.assert_array_equal(mid, omid)
        np.testing.assert_array_equal(dist, odist)
    # lattice with duplicated truth points: exact ties, also ties created by the rounding of sqrt
    g = np.arange(5.0)
    lat = np.stack(np.meshgrid(g, g, g, indexing="ij")).reshape(3, -1)
    truth = np.concatenate([lat, lat[:, ::-1]], axis=1)
    cen = rng.integers(0, 9, (3, 800)) * 0.5
    mid, dist = ctx.match_within(truth, cen, 0.8)
    omid, odist = oracle.match_within(truth, cen, 0.8)
    np.testing.assert_array_equal(mid, omid)
    np.testing.assert_array_equal(dist, odist)


def test_cluster_means_dev(ctx):
    # Tools.GetClusList's per-cluster averages (Tools.cs:187-194) as a segmented reduction
    import torch
    rng = np.random.default_rng(10)
    n, k = 20_000, 300
    cid = rng.integers(0, k + 1, n).astype(np.int32)
    cid[cid == 7] = 0                                     # an empty cluster
    vals = rng.normal(size=(5, n)) * 100 + 1000
    means, counts = ctx.cluster_means_dev(torch.from_numpy(cid).cuda(), k, torch.from_numpy(vals).cuda())
    means, counts = means.cpu().numpy(), counts.cpu().numpy()
    for c in range(1, k + 1):
        sel = cid == c
        assert counts[c] == sel.sum()
        if sel.any():
            np.testing.assert_allclose(means[:, c], [np.cumsum(v[sel])[-1] / sel.sum() for v in vals], rtol=1e-12)
        else:
            assert np.isnan(means[:, c]).all()
