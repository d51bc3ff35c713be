"""The oracle against the committed fixtures (tests/golden, made by tests/golden/make_golden.py)."""
from pathlib import Path

import numpy as np
import pytest

GOLD = Path(__file__).resolve().parent / "golden"


@pytest.mark.parametrize("name", ["c1", "lattice_ties", "nonfinite"])
def test_oracle_dbscan_golden(oracle, name):
    g = np.load(GOLD / f"dbscan_{name}.npz")
    for variant in ("literal", "grid"):
        cid, key, cls, amount = oracle.dbscan(g["mx"], g["my"], float(g["eps"]), int(g["min_pts"]), int(g["cf0"]), variant=variant)
        np.testing.assert_array_equal(cid, g["cluster_id"])
        np.testing.assert_array_equal(key, g["is_key"])
        np.testing.assert_array_equal(cls, g["is_classed"])
        assert amount == int(g["cluster_amount"])


@pytest.mark.parametrize("name", ["c1_checkerboard", "c3_small", "lattice_ties"])
def test_oracle_icp_golden(oracle, name):
    g = np.load(GOLD / f"icp_{name}.npz")
    for variant in ("literal", "grid"):
        order, sq = oracle.closest_point_set(g["model"], g["data"], variant)
        np.testing.assert_array_equal(order, g["order0"])
        np.testing.assert_array_equal(sq, g["sqdist0"])
    R, T, it, sse, olast = oracle.icp_rigid(g["model"], g["data"], float(g["e"]), int(g["max_iters"]), use_grid=True)
    assert it == int(g["iters"])
    np.testing.assert_array_equal(olast, g["order_last"])
    np.testing.assert_allclose(R, g["R"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(T, g["T"], rtol=0, atol=1e-12)


def test_oracle_stats_golden(oracle):
    g = np.load(GOLD / "stats_c1.npz")
    xyz, keep = oracle.polar_to_xyz(g["mx"], g["my"], g["dist"], 149.0, 307.0)
    np.testing.assert_allclose(xyz, g["xyz"], rtol=1e-13, atol=1e-13)      # libm's sin/cos may differ in the last bit across hosts
    st = oracle.cluster_stats(g["cluster_id"], int(g["n_clusters"]), g["xyz"], g["mx"], g["my"])
    for k in ("means", "circle3d", "circle2d"):
        np.testing.assert_array_equal(st[k][:, 1:].view(np.int64), g[k][:, 1:].view(np.int64))
    for k in ("counts", "status3d", "status2d"):
        np.testing.assert_array_equal(st[k], g[k])
    tid = np.arange(1, int(g["n_clusters"]) + 1, dtype=np.int32)
    np.testing.assert_array_equal(oracle.nearest_truth_2d(g["means"][3, 1:], g["means"][4, 1:], tid, g["mx"], g["my"], float(g["radius"])), g["nearest"])
