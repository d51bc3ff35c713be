"""Blocked clustering flow with libvpc as the engine vs the same flow with the oracle as the engine: identical."""
import numpy as np
import pytest

from vtkcloudpoint_b200 import blocked, synth

from blocked_helpers import oracle_dbscan, oracle_dbscan_cells

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("pts_in_cell", [200, 650])
def test_blocked_flow_gpu_equals_oracle(ctx, pts_in_cell):
    mx, my = synth.dbscan_cloud(0xC1, 14, n_total=10_000, decimals=3)
    xyz = np.stack([mx * 2.0, my * 3.0, mx - my], axis=1)
    got = blocked.cluster_blocked(mx, my, 0.07, 7, pts_in_cell, ctx.dbscan, ctx.dbscan_cells, points_xyz=xyz, argsort=ctx.argsort_f64)
    exp = blocked.cluster_blocked(mx, my, 0.07, 7, pts_in_cell, oracle_dbscan, oracle_dbscan_cells, points_xyz=xyz)
    assert got.cluster_amount == exp.cluster_amount and got.del_sum == exp.del_sum and got.cluster_sum_cells == exp.cluster_sum_cells
    np.testing.assert_array_equal(got.cluster_id, exp.cluster_id)
    np.testing.assert_array_equal(got.merge_order, exp.merge_order)
    np.testing.assert_array_equal(got.centers, exp.centers)
    np.testing.assert_array_equal(got.centers2d, exp.centers2d)
    d_got = blocked.merge_ids_by_distance(got.centers2d, got.center_ids, 0.1, ctx.dbscan)
    d_exp = blocked.merge_ids_by_distance(exp.centers2d, exp.center_ids, 0.1, oracle_dbscan)
    assert d_got == d_exp


@pytest.mark.parametrize("case", ["c1_200", "c1_650", "lattice", "uniform"])
def test_blocked_ref_single_call(ctx, case):
    """vpc_dbscan_blocked_ref (one C-ABI call) against the same flow driven by the oracle through the Python host mirror."""
    rng = np.random.default_rng(41)
    if case.startswith("c1"):
        mx, my = synth.dbscan_cloud(0xC1, 14, n_total=10_000, decimals=3)
        eps, min_pts, ppc = 0.07, 7, int(case.split("_")[1])
    elif case == "lattice":      # many points ON box edges and on xmin / ymin: the strict / inclusive bounds and the cut ties matter
        mx, my = 149.0 + rng.integers(0, 40, 6000) * 0.05, 307.0 + rng.integers(0, 40, 6000) * 0.05
        eps, min_pts, ppc = 0.05, 4, 300
    else:
        mx, my = rng.uniform(0, 3, 8000), rng.uniform(0, 2, 8000)
        eps, min_pts, ppc = 0.04, 5, 500
    got = ctx.dbscan_blocked_ref(mx, my, eps, min_pts, ppc)
    exp = blocked.cluster_blocked(mx, my, eps, min_pts, ppc, oracle_dbscan, oracle_dbscan_cells)
    assert got["rows"] == exp.partition.rows and got["cols"] == exp.partition.cols
    assert got["n_unassigned"] == len(exp.partition.dropped)
    assert got["cluster_sum"] == exp.cluster_amount and got["del_sum"] == exp.del_sum
    np.testing.assert_array_equal(got["cluster_id"], exp.cluster_id)
