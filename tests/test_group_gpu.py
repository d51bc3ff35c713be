"""ONE context over several GPUs from one process (vpc_create with n_devices > 1, csrc/group_api.cuh): the host-array calls a
P/Invoke shim makes -- DBImprovedGpu.dbscan / ICPGpu.go_hell_ICP -- spread over the devices inside libvpc.  On a one-GPU box the
devices are [0, 0, ...] (emulation: ranks share the GPU, phases enqueued in order); with >= 2 GPUs the real devices are used too."""
import os

import numpy as np
import pytest

from vtkcloudpoint_b200 import Context, synth

pytestmark = pytest.mark.gpu


def _device_sets():
    import torch
    n = torch.cuda.device_count()
    sets = [[0, 0], [0, 0, 0, 0]]
    if n >= 2:
        sets.append([0, 1])
    if n >= 4:
        sets.append([0, 1, 2, 3])
    if n >= 8:
        sets.append(list(range(8)))
    return sets


@pytest.fixture(scope="module", params=_device_sets(), ids=lambda d: "dev" + "".join(map(str, d)))
def gctx(request):
    os.environ["VPC_GROUP_MIN_POINTS"] = "2000"          # small clouds must take the multi-device path in these tests
    c = Context(request.param)
    yield c
    c.close()
    os.environ.pop("VPC_GROUP_MIN_POINTS", None)


def _check(gctx, oracle, mx, my, eps, min_pts, cf0=0, variant="grid"):
    before = gctx.launch_count
    got = gctx.dbscan(mx, my, eps, min_pts, cf0)
    cid, key, cls, amount = oracle.dbscan(mx, my, eps, min_pts, cf0, variant=variant)
    assert got.cluster_amount == amount
    np.testing.assert_array_equal(got.is_key, key)
    np.testing.assert_array_equal(got.is_classed, cls)
    np.testing.assert_array_equal(got.cluster_id, cid)
    return gctx.launch_count - before


def test_group_dbscan_c2_recipe_arbitrary_order(gctx, oracle):
    mx, my = synth.dbscan_cloud(0xC2, 44, n_total=100_000)          # shuffled: every chunk holds points of every slab
    launches = _check(gctx, oracle, mx, my, 0.07, 7, cf0=3)
    assert launches >= 10 * len(gctx.devices)                        # every rank ran the slab step
    _check(gctx, oracle, mx, my, 0.07, 7, cf0=0)                     # second call: heaps, epochs and workspaces are reused


def test_group_dbscan_sorted_input_and_chain(gctx, oracle):
    # spatially sorted input (a chunk holds a few slabs only) and one dense diagonal band through every slab
    rng = np.random.default_rng(8)
    t = np.sort(rng.uniform(0, 10, 50_000))
    mx, my = t + rng.normal(0, 0.01, t.size), t + rng.normal(0, 0.01, t.size)
    _check(gctx, oracle, mx, my, 0.05, 5)


def test_group_dbscan_nonfinite_and_ties(gctx, oracle):
    rng = np.random.default_rng(12)
    n = 30_000
    mx, my = rng.integers(0, 150, n) * 0.25, rng.integers(0, 150, n) * 0.25
    mx[::997] = np.nan; my[5::1013] = np.inf
    _check(gctx, oracle, mx, my, 0.25, 3)
    _check(gctx, oracle, mx, my, 0.5, 6, cf0=11)


def test_group_small_cloud_falls_back_to_one_device(gctx, oracle):
    rng = np.random.default_rng(1)
    mx, my = rng.uniform(0, 1, 500), rng.uniform(0, 1, 500)
    _check(gctx, oracle, mx, my, 0.05, 3, variant="literal")


def test_group_icp(gctx, oracle):
    model, data, _, _ = synth.icp_clouds(0xC3, 30_000, 4_000)
    res = gctx.icp_rigid(model, data, -1.0, 6)
    Ro, To, itd, sse, oo = oracle.icp_rigid(model, data, -1.0, 6)
    assert res.iters_done == itd == 6
    np.testing.assert_array_equal(res.order_last, oo)
    assert np.abs(res.R - Ro).max() < 1e-6 and np.abs(res.T - To).max() < 1e-6 * max(1.0, np.abs(To).max())
    assert abs(res.sse_last - sse) <= 1e-6 * sse
    # unbounded like the reference (ICP.cs:180): runs to |d - pre_d| < e
    res = gctx.icp_rigid(model, data, 1e-7, 0)
    Ro, To, itd, sse, oo = oracle.icp_rigid(model, data, 1e-7, 0)
    assert res.iters_done == itd
    np.testing.assert_array_equal(res.order_last, oo)
    assert abs(res.sse_last - sse) <= 1e-6 * sse


def test_pageable_and_page_locked_host_arrays_agree(ctx, oracle):
    """The host-pointer call with pageable NumPy arrays (worker threads + page-locked ring, csrc/host/staging.hpp) and with arrays
    registered as page-locked (vpc_host_register) gives the same result; sizes straddle the 1 MiB staging chunks."""
    import ctypes as C
    for n in (1, 131_071, 131_072, 131_073, 700_001):
        mx, my = synth.dbscan_cloud(0xC2, int((n * 0.7 / 40) ** 0.5) or 1, n_total=max(n, 40 * (int((n * 0.7 / 40) ** 0.5) or 1) ** 2))
        mx, my = np.ascontiguousarray(mx[:n]), np.ascontiguousarray(my[:n])
        a = ctx.dbscan(mx, my, 0.07, 7)
        for arr in (mx, my):
            assert ctx._lib.vpc_host_register(C.c_void_p(arr.ctypes.data), arr.nbytes) == 0
        try:
            b = ctx.dbscan(mx, my, 0.07, 7)
        finally:
            for arr in (mx, my):
                ctx._lib.vpc_host_unregister(C.c_void_p(arr.ctypes.data))
        assert a.cluster_amount == b.cluster_amount
        np.testing.assert_array_equal(a.cluster_id, b.cluster_id)
        np.testing.assert_array_equal(a.is_key, b.is_key)
        if n <= 131_073:
            cid, key, cls, amount = oracle.dbscan(mx, my, 0.07, 7, 0)
            np.testing.assert_array_equal(a.cluster_id, cid)


def test_group_blocked_cells_over_devices(gctx, oracle):
    """The StartCode work items (one per cell, FrmMain.cs:1356-1359) spread over the group's devices: same result as the oracle."""
    mx, my = synth.dbscan_cloud(0xC1, 14, n_total=10_000, decimals=3)
    for ppc in (200, 650):
        exp = oracle.blocked(mx, my, 0.07, 7, ppc)
        before = gctx.launch_count
        got = gctx.dbscan_blocked_ref(mx, my, 0.07, 7, ppc)
        assert gctx.launch_count - before > 20
        for k in ("rows", "cols", "n_unassigned", "cluster_sum", "del_sum"):
            assert got[k] == exp[k], k
        np.testing.assert_array_equal(got["cluster_id"], exp["cluster_id"])
        np.testing.assert_array_equal(got["merge_cid"], exp["merge_cid"])


def test_device_generator_matches_numpy_recipe(ctx):
    """vpc_synth_dbscan_cloud_dev (config C4 is generated on the device) reproduces synth.dbscan_cloud bit for bit."""
    for seed, grid, n, start, count in ((0xC2, 44, 100_000, 0, 100_000), (0xC4, 1400, 100_000_000, 73_000_123, 50_000), (0xC1, 14, 10_000, 17, 1000)):
        mx, my = synth.dbscan_cloud(seed, grid, n_total=n, start=start, count=count)
        dx, dy = ctx.synth_dbscan_cloud_dev(seed, grid, n, start, count)
        np.testing.assert_array_equal(dx.cpu().numpy().view(np.int64), mx.view(np.int64))
        np.testing.assert_array_equal(dy.cpu().numpy().view(np.int64), my.view(np.int64))
