"""Host mirror of the blocked clustering (FrmMain.cs:1214-1291, 1432-1544; Tools.cs:162-195, 521-621) with the oracle as engine."""
import numpy as np
import pytest

from vtkcloudpoint_b200 import blocked, synth

from blocked_helpers import oracle_dbscan, oracle_dbscan_cells


def test_partition_matches_reference_predicates():
    rng = np.random.default_rng(0)
    mx, my = rng.uniform(10, 12, 3000), rng.uniform(-3, -1, 3000)
    part = blocked.partition_cells(mx, my, 200)
    xmin, ymin, xmax, ymax = mx.min(), my.min(), mx.max(), my.max()
    key = np.maximum(mx - xmin, my - ymin)
    cell0 = np.argsort(key, kind="stable")[:200]
    assert set(part.order[part.offsets[0]:part.offsets[1]].tolist()) == set(cell0.tolist())
    cx, cy = mx[cell0].max() - xmin, my[cell0].max() - ymin
    assert part.rows == int((ymax - ymin) / cy) + 1 and part.cols == int((xmax - xmin) / cx) + 1
    seen = np.zeros(len(mx), int)
    for c in range(part.rows * part.cols):
        idx = part.order[part.offsets[c]:part.offsets[c + 1]]
        seen[idx] += 1
        if c == 0:
            continue
        p, q = divmod(c, part.cols)
        hx = xmax if q == part.cols - 1 else xmin + (q + 1) * cx
        hy = ymax if p == part.rows - 1 else ymin + (p + 1) * cy
        assert ((mx[idx] > xmin + q * cx) & (my[idx] > ymin + p * cy) & (mx[idx] <= hx) & (my[idx] <= hy)).all()
    assert seen.max() == 1                                      # cells are disjoint
    assert set(np.flatnonzero(seen == 0).tolist()) == set(part.dropped.tolist())
    # the strict lower bounds drop the points sitting on the x_min / y_min edges outside cell 0
    edge = ((mx == xmin) | (my == ymin)) & ~np.isin(np.arange(len(mx)), cell0)
    assert set(np.flatnonzero(edge).tolist()) <= set(part.dropped.tolist())


def test_complete_work3_small_cluster_rule_and_off_by_one():
    # two cells given directly: local ids chosen to hit (a) a <=3-point cluster that is dropped, (b) the C#'s off-by-one
    # when a cell starts with a non-noise point, (c) the never-checked last cluster of a cell
    order = np.arange(20)
    offsets = np.array([0, 12, 20])
    local = np.array([0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 3, 3,            # cell 0: noise x2, A(4), B(2: too small), C(4)
                      1, 1, 2, 2, 2, 2, 3, 3])                       # cell 1: starts non-noise: A'(2) counted as 3 -> dropped with one extra
    part = blocked.CellPartition(order, offsets, 1, 2, np.empty(0, np.int64))
    mx = np.arange(20) * 10.0                                        # far apart: the noise re-cluster finds nothing
    my = np.zeros(20)
    cid, amount, merge, del_sum, cluster_sum = blocked.complete_work3(part, local, np.array([3, 3]), mx, my, 0.5, 2, oracle_dbscan)
    assert cluster_sum == 7 and del_sum == 2
    assert cid[2:6].tolist() == [1] * 4                              # A keeps id 1
    assert cid[6:8].tolist() == [0, 0]                               # B (2 points) zeroed, its id is reused
    assert cid[8:11].tolist() == [2] * 3                             # C gets id 2 (its last point is hit by cell 1's walk-back, below)
    # cell 1: A' has 2 points but clusLen = 3 (off-by-one) -> dropped, and the walk-back zeroes 3 entries: A' and the
    # LAST point of cell 0's cluster C that happens to precede them in clusForMerge
    assert cid[12:14].tolist() == [0, 0] and cid[11] == 0
    assert cid[14:18].tolist() == [3] * 4
    assert cid[18:20].tolist() == [4, 4]                             # last cluster of a cell is never size-checked
    assert amount == cluster_sum - del_sum - 1                       # nothing re-clustered: clusterAmount = cf


def test_blocked_flow_c1_with_oracle_engine():
    mx, my = synth.dbscan_cloud(0xC1, 14, n_total=10_000, decimals=3)
    res = blocked.cluster_blocked(mx, my, 0.07, 7, 200, oracle_dbscan, oracle_dbscan_cells)
    assert res.partition.rows * res.partition.cols == len(res.partition.offsets) - 1
    assert res.cluster_id.min() == 0 and res.cluster_id.max() <= res.cluster_amount
    assert (res.cluster_id[res.partition.dropped] == 0).all()
    # blocking splits some clusters at cell borders (no halo): at least the 196 true clusters, not wildly more
    n_found = len(np.unique(res.cluster_id[res.cluster_id > 0]))
    assert 196 <= n_found <= 2 * 196
    assert len(res.center_ids) == n_found and res.centers2d.shape == (n_found, 2)
    # centroid merge at the reference's default threshold 0.1 (Clustering.Designer.cs:228) re-joins split clusters
    dick = blocked.merge_ids_by_distance(res.centers2d, res.center_ids, 0.1, oracle_dbscan)
    new_id, new_amount = blocked.refresh_by_dictionary(res.cluster_id, res.cluster_amount, dick)
    assert new_amount == res.cluster_amount - len(dick)
    assert len(np.unique(new_id[new_id > 0])) <= n_found
