"""The oracle's literal, List-based restatement of the blocked clustering (oracle/vpc_oracle_blocked.cpp; FrmMain.cs:1214-1291,
1340-1361, 2782-2794, 1432-1544; Tools.cs:162-195, 510-513, 521-621) checked against what the C# text implies."""
import numpy as np
import pytest

from vtkcloudpoint_b200 import synth


def _box_predicates_hold(mx, my, res, ppc):
    """Every assigned point satisfies the literal box predicate of its cell (Tools.cs:510-513) or is one of the first ppc sorted points."""
    xmin, ymin, xmax, ymax = mx.min(), my.min(), mx.max(), my.max()
    key = np.maximum(mx - xmin, my - ymin)
    cell0 = np.argsort(key, kind="stable")[:ppc]
    cx, cy = mx[cell0].max() - xmin, my[cell0].max() - ymin
    assert res["rows"] == int((ymax - ymin) / cy) + 1 and res["cols"] == int((xmax - xmin) / cx) + 1
    return cell0


def test_blocked_c1(oracle):
    mx, my = synth.dbscan_cloud(0xC1, 14, n_total=10_000, decimals=3)
    res = oracle.blocked(mx, my, 0.07, 7, 200)
    _box_predicates_hold(mx, my, res, 200)
    assert res["cluster_sum_cells"] - res["del_sum"] - 1 <= res["cluster_sum"]
    assert res["cluster_id"].min() == 0 and res["cluster_id"].max() <= res["cluster_sum"]
    # blocking splits some clusters at cell borders (no halo): at least the 196 true clusters, not wildly more
    n_found = len(np.unique(res["cluster_id"][res["cluster_id"] > 0]))
    assert 196 <= n_found <= 2 * 196
    # points that fall into no cell keep 0 and are not part of clusForMerge
    taken = np.zeros(len(mx), bool); taken[res["merge_order"]] = True
    assert (~taken).sum() == res["n_unassigned"] and (res["cluster_id"][~taken] == 0).all()
    # the strict lower bounds drop points sitting on the x_min / y_min edges outside the first cell
    cell0 = _box_predicates_hold(mx, my, res, 200)
    edge = ((mx == mx.min()) | (my == my.min())) & ~np.isin(np.arange(len(mx)), cell0)
    assert not taken[edge].any()
    # clusForMerge: kept entries first, then the re-clustered noise (FrmMain.cs:1511, 1517-1520)
    assert np.array_equal(res["cluster_id"][res["merge_order"]], res["merge_cid"]) or res["n_shared"] > 0


@pytest.mark.parametrize("seed", range(6))
def test_fast_variant_equals_literal(oracle, seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(500, 4000))
    mx, my = np.round(rng.uniform(-1, 2, n), 3), np.round(rng.uniform(-0.5, 1.5, n), 3)
    args = (mx, my, float(rng.uniform(0.03, 0.08)), int(rng.integers(2, 6)), int(rng.integers(50, 400)))
    try:
        a = oracle.blocked(*args)
    except oracle.ReferenceThrows:
        with pytest.raises(oracle.ReferenceThrows):
            oracle.blocked(*args, fast=True, n_threads=3)
        return
    b = oracle.blocked(*args, fast=True, n_threads=3)
    for k in a:
        assert np.array_equal(a[k], b[k]), k


def test_small_cluster_drop_and_off_by_one(oracle):
    """minPts = 2: clusters of 2-3 points exist, so the <= 3 drop (FrmMain.cs:1481) and its off-by-one (a cell whose first sorted entry
    is not noise counts its first cluster one too long, :1461-1465 / :1498) fire.  Consequences checked: every dropped cluster lowers the
    id budget by one (cf = clusterSum - delSum - 1), and the re-cluster continues the numbering at cf + 1."""
    rng = np.random.default_rng(4)
    mx, my = rng.uniform(0, 3, 4000), rng.uniform(0, 2, 4000)
    res = oracle.blocked(mx, my, 0.03, 2, 150)
    assert res["del_sum"] > 50
    cf = res["cluster_sum_cells"] - res["del_sum"] - 1
    assert res["cluster_sum"] >= cf
    ids = np.unique(res["merge_cid"][res["merge_cid"] > 0])
    assert ids.max() == res["cluster_sum"]
    # ids up to cf come from the per-cell pass, ids above from the global noise re-cluster, which is dense 1.. by construction
    assert np.array_equal(ids[ids > cf], np.arange(cf + 1, res["cluster_sum"] + 1))


def test_shared_object_race_is_visible_and_counted(oracle):
    """x_Min + 1 * cell_x can round one ulp below the first cell's own maximum: that point then also passes the strict lower bound of box
    (0,1) and sits in two cells -- one Point3D written by two pool threads in the C#.  n_shared counts such points; with none the
    copy semantics (product) and the sequential shared-object schedule agree exactly, with some they may differ."""
    agree_when_unshared, seen_shared, seen_diff = 0, 0, 0
    for seed in range(60):
        rng = np.random.default_rng(seed)
        mx, my = np.round(rng.uniform(-1, 2, 1500), 3), np.round(rng.uniform(-0.5, 1.5, 1500), 3)
        try:
            a = oracle.blocked(mx, my, 0.06, 3, 100)
            s = oracle.blocked(mx, my, 0.06, 3, 100, shared_objects=True)
        except oracle.ReferenceThrows:
            continue
        assert a["n_shared"] == s["n_shared"]
        if a["n_shared"] == 0:
            assert np.array_equal(a["cluster_id"], s["cluster_id"]) and a["cluster_sum"] == s["cluster_sum"]
            agree_when_unshared += 1
        else:
            seen_shared += 1
            seen_diff += int(not np.array_equal(a["cluster_id"], s["cluster_id"]))
    assert agree_when_unshared > 10 and seen_shared > 5 and seen_diff > 0
    # motor-angle data (149.., 307..: the subtraction mx - x_Min is exact by Sterbenz) never shares
    mx, my = synth.dbscan_cloud(0xC1, 14, n_total=10_000, decimals=3)
    assert oracle.blocked(mx, my, 0.07, 7, 200)["n_shared"] == 0


def test_merge_ids_by_distance(oracle):
    mx, my = synth.dbscan_cloud(0xC1, 14, n_total=10_000, decimals=3)
    res = oracle.blocked(mx, my, 0.07, 7, 200)
    order = res["merge_order"]
    xyz = np.stack([mx, my, mx - my])          # X = motor_x, Y = motor_y: the merge threshold 0.1 is in those units
    m = oracle.merge_ids(res["merge_cid"], xyz[:, order], mx[order], my[order], res["cluster_sum"], 0.1)
    # every dictionary entry removes one id; the survivors are renumbered 1.. without gaps
    assert m["cluster_amount"] == res["cluster_sum"] - len(m["dict"])
    ids = np.unique(m["cluster_id"][m["cluster_id"] > 0])
    assert np.array_equal(ids, np.arange(1, m["cluster_amount"] + 1))
    assert len(m["dict"]) > 0                                   # cell borders split clusters; the centroid merge re-joins them
    # merged pairs were within the threshold's reach of each other as centre clusters (chains allowed): same new id
    old = res["merge_cid"]
    for src, dst in m["dict"][:20]:
        assert len(set(m["cluster_id"][old == src].tolist()) | set(m["cluster_id"][old == dst].tolist())) == 1
    # centroids are LINQ Averages in list order
    c = int(m["center_ids"][0])
    sel = np.flatnonzero(old == c)
    assert m["centers5"][0, 0] == np.cumsum(xyz[0, order][sel])[-1] / len(sel)
    # a cluster id without points makes the C# throw (Average over an empty list, Tools.cs:565)
    bad = res["merge_cid"].copy(); bad[bad == 3] = 0
    with pytest.raises(oracle.ReferenceThrows):
        oracle.merge_ids(bad, xyz[:, order], mx[order], my[order], res["cluster_sum"], 0.1)
