"""The CUDA path (through the C ABI) against the committed fixtures: bit-exact labels / flags / correspondences,
R, T, SSE within 1e-6 relative."""
from pathlib import Path

import numpy as np
import pytest

GOLD = Path(__file__).resolve().parent / "golden"
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["c1", "lattice_ties", "nonfinite"])
def test_dbscan_golden(ctx, name):
    g = np.load(GOLD / f"dbscan_{name}.npz")
    got = ctx.dbscan(g["mx"], g["my"], float(g["eps"]), int(g["min_pts"]), int(g["cf0"]))
    np.testing.assert_array_equal(got.cluster_id, g["cluster_id"])
    np.testing.assert_array_equal(got.is_key, g["is_key"])
    np.testing.assert_array_equal(got.is_classed, g["is_classed"])
    assert got.cluster_amount == int(g["cluster_amount"])


@pytest.mark.parametrize("name", ["c1_checkerboard", "c3_small", "lattice_ties"])
def test_icp_golden(ctx, name):
    g = np.load(GOLD / f"icp_{name}.npz")
    order, sq = ctx.closest_point_set(g["model"], g["data"])
    np.testing.assert_array_equal(order, g["order0"])
    np.testing.assert_array_equal(sq, g["sqdist0"])
    res = ctx.icp_rigid(g["model"], g["data"], float(g["e"]), int(g["max_iters"]))
    assert res.iters_done == int(g["iters"])
    np.testing.assert_array_equal(res.order_last, g["order_last"])
    scale = max(np.abs(g["T"]).max(), 1.0)
    assert np.abs(res.R - g["R"]).max() < 1e-6 and np.abs(res.T - g["T"]).max() < 1e-6 * scale
    assert abs(res.sse_last - float(g["sse"])) <= 1e-6 * max(float(g["sse"]), 1e-30)


def test_stats_golden(ctx):
    g = np.load(GOLD / "stats_c1.npz")
    st = ctx.cluster_stats(g["cluster_id"], int(g["n_clusters"]), g["xyz"], g["mx"], g["my"])
    for k in ("means", "circle3d", "circle2d"):
        np.testing.assert_array_equal(st[k][:, 1:].view(np.int64), g[k][:, 1:].view(np.int64))      # bit patterns
    for k in ("counts", "status3d", "status2d"):
        np.testing.assert_array_equal(st[k], g[k])
    tid = np.arange(1, int(g["n_clusters"]) + 1, dtype=np.int32)
    np.testing.assert_array_equal(ctx.nearest_truth_2d(g["means"][3, 1:], g["means"][4, 1:], tid, g["mx"], g["my"], float(g["radius"])), g["nearest"])
    got = ctx.dbscan(g["mx"], g["my"], 0.07, 7, 0)
    np.testing.assert_array_equal(got.cluster_id, g["cluster_id"])
