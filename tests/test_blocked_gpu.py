"""The device-resident blocked clustering (vpc_dbscan_blocked_ref_ex, vpc_merge_ids_by_distance; csrc/blocked.cuh) against the
oracle's literal List-based restatement (oracle/vpc_oracle_blocked.cpp): identical, clusForMerge order and centroids included."""
import numpy as np
import pytest

from vtkcloudpoint_b200 import VpcError, blocked, synth

pytestmark = pytest.mark.gpu


def _cases():
    rng = np.random.default_rng(41)
    c1 = synth.dbscan_cloud(0xC1, 14, n_total=10_000, decimals=3)
    yield "c1_200", c1[0], c1[1], 0.07, 7, 200
    yield "c1_650", c1[0], c1[1], 0.07, 7, 650
    # many points ON box edges and on xmin / ymin: the strict / inclusive bounds and the cut ties matter
    yield "lattice", 149.0 + rng.integers(0, 40, 6000) * 0.05, 307.0 + rng.integers(0, 40, 6000) * 0.05, 0.05, 4, 300
    yield "uniform", rng.uniform(0, 3, 8000), rng.uniform(0, 2, 8000), 0.04, 5, 500
    yield "small_minpts", rng.uniform(0, 3, 4000), rng.uniform(0, 2, 4000), 0.03, 2, 150          # the <= 3 drop and its off-by-one
    yield "tiny_cells", rng.uniform(0, 1, 3000), rng.uniform(0, 1, 3000), 0.02, 2, 7              # hundreds of cells, many empty
    for seed in (1, 2, 5, 9, 10):                                                                  # clouds with SHARED points (see the oracle)
        r = np.random.default_rng(seed)
        yield f"shared_{seed}", np.round(r.uniform(-1, 2, 1500), 3), np.round(r.uniform(-0.5, 1.5, 1500), 3), 0.06, 3, 100


@pytest.mark.parametrize("case", list(_cases()), ids=lambda c: c[0])
def test_blocked_ref_vs_literal_oracle(ctx, oracle, case):
    name, mx, my, eps, min_pts, ppc = case
    exp = oracle.blocked(mx, my, eps, min_pts, ppc)                       # copy semantics = the product's defined behaviour
    got = blocked.cluster_blocked(ctx, mx, my, eps, min_pts, ppc)
    for k in ("rows", "cols", "n_unassigned", "n_shared", "cluster_sum", "del_sum", "cluster_sum_cells"):
        assert got[k] == exp[k], k
    np.testing.assert_array_equal(got["cluster_id"], exp["cluster_id"])
    np.testing.assert_array_equal(got["merge_order"], exp["merge_order"])
    np.testing.assert_array_equal(got["merge_cid"], exp["merge_cid"])
    if name.startswith("shared"):
        assert exp["n_shared"] > 0


@pytest.mark.parametrize("pts_in_cell", [200, 650])
def test_merge_ids_by_distance_vs_literal_oracle(ctx, oracle, pts_in_cell):
    mx, my = synth.dbscan_cloud(0xC1, 14, n_total=10_000, decimals=3)
    xyz = np.stack([mx, my, mx * 0.5 - my])
    res = blocked.cluster_blocked(ctx, mx, my, 0.07, 7, pts_in_cell)
    got = blocked.merge_clusters(ctx, res, xyz, mx, my, 0.1)
    order = res["merge_order"]
    exp = oracle.merge_ids(res["merge_cid"], xyz[:, order], mx[order], my[order], res["cluster_sum"], 0.1)
    assert got["cluster_amount"] == exp["cluster_amount"] and got["dict"] == exp["dict"] and len(exp["dict"]) > 0
    np.testing.assert_array_equal(got["cluster_id"], exp["cluster_id"])
    np.testing.assert_array_equal(got["center_ids"], exp["center_ids"])
    np.testing.assert_array_equal(got["centers5"].view(np.int64), exp["centers5"].view(np.int64))            # bit-identical centroids
    np.testing.assert_array_equal(got["new_centers5"].view(np.int64), exp["new_centers5"].view(np.int64))


def test_merge_ids_reports_where_the_reference_throws(ctx, oracle):
    mx, my = synth.dbscan_cloud(0xC1, 14, n_total=10_000, decimals=3)
    res = blocked.cluster_blocked(ctx, mx, my, 0.07, 7, 200)
    bad = dict(res); bad["merge_cid"] = res["merge_cid"].copy(); bad["merge_cid"][bad["merge_cid"] == 3] = 0
    with pytest.raises(VpcError) as e:
        blocked.merge_clusters(ctx, bad, np.stack([mx, my, mx]), mx, my, 0.1)
    assert e.value.code == -6


def test_blocked_ref_rejects_what_the_reference_cannot_do(ctx):
    with pytest.raises(VpcError):
        ctx.dbscan_blocked_ref(np.array([1.0, np.nan, 2.0]), np.array([0.0, 1.0, 2.0]), 0.1, 2, 2)        # NaN keys: order-dependent Min / Sort
    with pytest.raises(VpcError):
        ctx.dbscan_blocked_ref(np.ones(50), np.arange(50.0), 0.1, 2, 10)                                   # cell_x == 0: the C# divides by zero


def test_blocked_ref_1m_points_vs_fast_oracle(ctx, oracle):
    """Config-C2-sized cloud (5000 cells): the oracle's fast variant (same lists, separable box scan, grid re-cluster; equal to the
    literal one by tests/test_blocked_cpu.py) on all host cores."""
    import os
    mx, my = synth.dbscan_cloud(0xC2, 140, n_total=1_000_000)
    exp = oracle.blocked(mx, my, 0.07, 7, 200, fast=True, n_threads=os.cpu_count() or 8)
    got = blocked.cluster_blocked(ctx, mx, my, 0.07, 7, 200)
    for k in ("rows", "cols", "n_unassigned", "n_shared", "cluster_sum", "del_sum"):
        assert got[k] == exp[k], k
    np.testing.assert_array_equal(got["cluster_id"], exp["cluster_id"])
    np.testing.assert_array_equal(got["merge_cid"], exp["merge_cid"])
