"""GPU parity: vpc_dbscan_l1_2d (through the C ABI) vs the CPU oracle, bit-exact."""
import numpy as np
import pytest

from vtkcloudpoint_b200 import synth

pytestmark = pytest.mark.gpu


def _check(ctx, oracle, mx, my, eps, min_pts, cf0=0, variant="grid"):
    got = ctx.dbscan(mx, my, eps, min_pts, cf0)
    cid, key, cls, amount = oracle.dbscan(mx, my, eps, min_pts, cf0, variant=variant)
    assert got.cluster_amount == amount
    np.testing.assert_array_equal(got.is_key, key)
    np.testing.assert_array_equal(got.is_classed, cls)
    np.testing.assert_array_equal(got.cluster_id, cid)
    return got


def test_c1_literal(ctx, oracle):
    # config C1: 10k-point cloud written with 3 decimals, eps 0.07, minPts 7 -- against the LITERAL oracle
    mx, my = synth.dbscan_cloud(0xC1, 14, n_total=10_000, decimals=3)
    got = _check(ctx, oracle, mx, my, 0.07, 7, variant="literal")
    assert got.cluster_amount == 196


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5])
def test_random_uniform_small(ctx, oracle, seed):
    rng = np.random.default_rng(seed)
    n = int(rng.integers(1, 3000))
    mx = rng.uniform(0, 2.0, n)
    my = rng.uniform(0, 2.0, n)
    _check(ctx, oracle, mx, my, float(rng.uniform(0.02, 0.2)), int(rng.integers(1, 9)), int(rng.integers(0, 50)), variant="literal")


def test_quantised_ties(ctx, oracle):
    # lattice coordinates: many pairs sit exactly on |dx|+|dy| == eps (inclusive '<=')
    rng = np.random.default_rng(7)
    n = 4000
    mx = rng.integers(0, 60, n) * 0.25
    my = rng.integers(0, 60, n) * 0.25
    for eps in (0.25, 0.5, 0.75):
        for min_pts in (2, 4, 6):
            _check(ctx, oracle, mx, my, eps, min_pts, variant="literal")


def test_border_last_writer_wins(ctx, oracle):
    # two dense groups and a NON-core point within eps of one member of each: it must take the LARGER id (:87)
    t = np.arange(6) * 0.001
    a = np.stack([-t, np.zeros(6)], axis=1)            # a[0] = (0, 0)
    b = np.stack([0.2 + t, np.zeros(6)], axis=1)       # b[0] = (0.2, 0)
    bridge = np.array([[0.1, 0.0]])
    pts = np.vstack([a, bridge, b])
    got = _check(ctx, oracle, pts[:, 0].copy(), pts[:, 1].copy(), 0.1, 6, variant="literal")
    assert got.cluster_amount == 2
    assert got.cluster_id[6] == 2 and got.is_key[6] == 0 and got.is_classed[6] == 1
    # same cloud with the groups swapped in index order: the bridge still takes the larger id
    pts = np.vstack([b, bridge, a])
    got = _check(ctx, oracle, pts[:, 0].copy(), pts[:, 1].copy(), 0.1, 6, variant="literal")
    assert got.cluster_id[6] == 2


def test_edge_cases(ctx, oracle):
    rng = np.random.default_rng(11)
    mx = rng.uniform(0, 1, 500)
    my = rng.uniform(0, 1, 500)
    mx2, my2 = mx.copy(), my.copy()
    mx2[[3, 77]] = np.nan
    my2[[5]] = np.inf
    mx2[[9]] = -np.inf
    for (x, y) in ((mx, my), (mx2, my2)):
        for eps, min_pts in ((0.05, 4), (0.05, 1), (0.05, 0), (0.05, -3), (-1.0, 3), (-1.0, 0), (float("nan"), 2), (0.0, 1), (0.0, 2),
                             (5.0, 3), (1e-300, 1)):
            _check(ctx, oracle, x, y, eps, min_pts, 7, variant="literal")
    # duplicates count toward each other's density
    d = np.repeat(rng.uniform(0, 1, 40), 5)
    _check(ctx, oracle, d, d[::-1].copy(), 0.0, 5, variant="literal")
    # n = 0 and n = 1
    e = np.empty(0)
    got = ctx.dbscan(e, e, 0.07, 7, 5)
    assert got.cluster_amount == 5 and got.cluster_id.size == 0
    _check(ctx, oracle, np.array([1.0]), np.array([2.0]), 0.07, 1, variant="literal")
    _check(ctx, oracle, np.array([1.0]), np.array([2.0]), 0.07, 2, variant="literal")


def test_huge_extent_and_offsets(ctx, oracle):
    rng = np.random.default_rng(13)
    n = 2000
    mx = np.concatenate([rng.normal(0, 0.02, n // 2), rng.normal(1e9, 0.02, n // 2)])
    my = np.concatenate([rng.normal(-1e12, 0.02, n // 2), rng.normal(5e11, 0.02, n // 2)])
    _check(ctx, oracle, mx, my, 0.03, 5, variant="literal")
    mx = rng.normal(0, 1e-3, n) + 1e6
    my = rng.normal(0, 1e-3, n) - 1e6
    _check(ctx, oracle, mx, my, 2e-4, 4, variant="literal")


def test_eps_inf_rejected(ctx):
    from vtkcloudpoint_b200 import VpcError
    with pytest.raises(VpcError):
        ctx.dbscan(np.zeros(3), np.zeros(3), float("inf"), 2)


def test_c2_full_size_vs_grid_oracle(ctx, oracle):
    # config C2: 1M points; oracle = grid variant (validated against the literal one in the CPU suite)
    mx, my = synth.dbscan_cloud(0xC2, 140, n_total=1_000_000)
    got = _check(ctx, oracle, mx, my, 0.07, 7)
    assert got.cluster_amount >= 19_600
    # idempotence / permutation property: relabelling is stable under a second run
    again = ctx.dbscan(mx, my, 0.07, 7)
    np.testing.assert_array_equal(again.cluster_id, got.cluster_id)


def test_device_pointer_entry(ctx, oracle):
    import torch
    mx, my = synth.dbscan_cloud(0xC1, 14, n_total=10_000)
    tx = torch.from_numpy(mx).cuda()
    ty = torch.from_numpy(my).cuda()
    cid, key, cls, amount = ctx.dbscan_dev(tx, ty, 0.07, 7, 3)
    torch.cuda.synchronize()
    ocid, okey, ocls, oamount = oracle.dbscan(mx, my, 0.07, 7, 3)
    assert int(amount.item()) == oamount
    np.testing.assert_array_equal(cid.cpu().numpy(), ocid)
    np.testing.assert_array_equal(key.cpu().numpy(), okey)
    np.testing.assert_array_equal(cls.cpu().numpy(), ocls)


def test_cells_batched_matches_per_cell_oracle(ctx, oracle):
    # all StartCode work items (FrmMain.cs:2782-2794) in one launch: independent clouds, cell-local ids
    rng = np.random.default_rng(21)
    sizes = [0, 300, 1, 0, 777, 64, 1500, 0]
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    n = int(offs[-1])
    # overlapping extents on purpose: points of different cells must never interact
    mx = rng.uniform(0, 1.0, n)
    my = rng.uniform(0, 1.0, n)
    for eps, min_pts in ((0.05, 4), (0.03, 2), (0.08, 7)):
        got, per_cell = ctx.dbscan_cells(mx, my, offs, eps, min_pts)
        for k in range(len(sizes)):
            a, b = int(offs[k]), int(offs[k + 1])
            cid, key, cls, amount = oracle.dbscan(mx[a:b], my[a:b], eps, min_pts, 0, variant="literal")
            np.testing.assert_array_equal(got.cluster_id[a:b], cid)
            np.testing.assert_array_equal(got.is_key[a:b], key)
            np.testing.assert_array_equal(got.is_classed[a:b], cls)
            assert per_cell[k] == amount
    with pytest.raises(Exception):
        ctx.dbscan_cells(mx, my, np.array([0, 5, 3, n]), 0.05, 4)


def test_slab_kernels_single_rank(ctx, oracle):
    # the slab-mode exports (local keys -> merge table -> finish) through the distributed driver at world size 1
    import torch
    from vtkcloudpoint_b200.distributed import GpuBackend, dbscan_slabs
    mx, my = synth.dbscan_cloud(0xC1, 14, n_total=10_000)
    cid, key, cls, amount = dbscan_slabs(GpuBackend(ctx), torch.from_numpy(mx).cuda(), torch.from_numpy(my).cuda(), 0, 0.07, 7, 11)
    ocid, okey, ocls, oamount = oracle.dbscan(mx, my, 0.07, 7, 11)
    assert amount == oamount
    np.testing.assert_array_equal(cid.cpu().numpy(), ocid)
    np.testing.assert_array_equal(key.cpu().numpy(), okey)
    np.testing.assert_array_equal(cls.cpu().numpy(), ocls)


def test_uf_edges(ctx):
    import torch
    from vtkcloudpoint_b200.distributed import GpuBackend
    rng = np.random.default_rng(5)
    n_nodes, n_edges = 5000, 4000
    a = rng.integers(0, n_nodes, n_edges).astype(np.int32)
    b = rng.integers(0, n_nodes, n_edges).astype(np.int32)
    root = GpuBackend(ctx).uf_edges(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), n_nodes).cpu().numpy()
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    _, lab = connected_components(coo_matrix((np.ones(n_edges), (a, b)), shape=(n_nodes, n_nodes)), directed=False)
    mins = np.full(lab.max() + 1, n_nodes)
    np.minimum.at(mins, lab, np.arange(n_nodes))
    np.testing.assert_array_equal(root, mins[lab])


def test_slab_lean_single_rank(ctx, oracle):
    # the sync-free pre-cut path (count-prefixed buffers, NaN padding) at world size 1
    import torch
    from vtkcloudpoint_b200.distributed import LeanSlabPlan, dbscan_slabs_lean
    mx, my = synth.dbscan_cloud(0xC1, 14, n_total=10_000)
    dev = torch.device("cuda", 0)
    plan = LeanSlabPlan(ctx, len(mx), [], 0.07, float(np.abs(mx).max() + np.abs(my).max()), dev)
    for first in (0, 5):
        cid, key, cls, amount, overflow = dbscan_slabs_lean(plan, torch.from_numpy(mx).to(dev), torch.from_numpy(my).to(dev), 0, 7, first)
        torch.cuda.synchronize()
        ocid, okey, ocls, oamount = oracle.dbscan(mx, my, 0.07, 7, first)
        assert int(overflow.item()) == 0 and int(amount.item()) == oamount
        np.testing.assert_array_equal(cid.cpu().numpy(), ocid)
        np.testing.assert_array_equal(key.cpu().numpy(), okey)
        np.testing.assert_array_equal(cls.cpu().numpy(), ocls)


@pytest.fixture
def banded_mode():
    """Force the band-partitioned counting sort (normally taken by clouds of >= 3M points) at any size."""
    import os
    os.environ["VPC_DB_BAND_MIN"] = "1"
    yield
    os.environ.pop("VPC_DB_BAND_MIN", None)


def test_banded_path_small_cases(ctx, oracle, banded_mode):
    # the same semantics through the band partition: C1, lattice ties, non-finite points, min_pts <= 0, eps < 0, huge extents
    mx, my = synth.dbscan_cloud(0xC1, 14, n_total=10_000, decimals=3)
    assert _check(ctx, oracle, mx, my, 0.07, 7, variant="literal").cluster_amount == 196
    rng = np.random.default_rng(21)
    lx, ly = rng.integers(0, 60, 4000) * 0.25, rng.integers(0, 60, 4000) * 0.25
    _check(ctx, oracle, lx, ly, 0.5, 4, variant="literal")
    x, y = rng.uniform(0, 1, 700), rng.uniform(0, 1, 700)
    x[[3, 77]] = np.nan; y[5] = np.inf; x[9] = -np.inf
    for eps, min_pts in ((0.05, 4), (0.05, 1), (0.05, 0), (0.05, -3), (-1.0, 3), (-1.0, 0), (float("nan"), 2), (0.0, 1), (5.0, 3)):
        _check(ctx, oracle, x, y, eps, min_pts, 7, variant="literal")
    x = np.concatenate([rng.uniform(0, 1, 300), 1e9 + rng.uniform(0, 1, 300)])
    _check(ctx, oracle, x, rng.uniform(0, 1, 600), 0.05, 3, variant="literal")
    for n in (1, 2, 255, 257, 4097):
        _check(ctx, oracle, rng.uniform(0, 1, n), rng.uniform(0, 1, n), 0.05, 2, variant="literal")


def test_banded_equals_direct_at_c2(ctx, oracle, banded_mode):
    mx, my = synth.dbscan_cloud(0xC2, 140, n_total=1_000_000)
    banded = ctx.dbscan(mx, my, 0.07, 7, 0)
    import os
    os.environ.pop("VPC_DB_BAND_MIN", None)
    direct = ctx.dbscan(mx, my, 0.07, 7, 0)
    assert banded.cluster_amount == direct.cluster_amount >= 19600
    np.testing.assert_array_equal(banded.cluster_id, direct.cluster_id)
    np.testing.assert_array_equal(banded.is_key, direct.is_key)
    np.testing.assert_array_equal(banded.is_classed, direct.is_classed)


def test_big_adjacent_cells(ctx, oracle):
    # cells holding hundreds of core points next to each other (many candidate pairs per cell pair): connected and
    # unconnected variants, plus a one-pair bridge
    rng = np.random.default_rng(31)
    eps = 1.0
    for gap, bridge in ((1.85, False), (1.85, True), (0.9, False), (1.02, False)):
        a = rng.uniform(0, 0.08, (400, 2))
        b = rng.uniform(0, 0.08, (400, 2)) + [gap, 0.0]
        pts = [a, b, rng.uniform(-3, 5, (300, 2))]
        if bridge:
            pts.append(np.array([[0.08 + 0.9, 0.04]]))         # within eps of both blobs' near edges
        pts = np.vstack(pts)
        perm = rng.permutation(len(pts))
        _check(ctx, oracle, pts[perm, 0].copy(), pts[perm, 1].copy(), eps, 5)
    # lattice data with eps on the lattice: pairs at exactly eps across cell boundaries two cells apart
    g = rng.integers(0, 40, (6000, 2)) * 0.5
    _check(ctx, oracle, g[:, 0].copy(), g[:, 1].copy(), 1.0, 3)
    _check(ctx, oracle, g[:, 0].copy(), g[:, 1].copy(), 0.5, 2)


def test_launch_and_staging_switches_do_not_change_results(oracle, monkeypatch):
    # VPC_PDL=0 (plain launches), VPC_COPY_THREADS=0 (plain cudaMemcpy) and a forced worker count: contexts read the switches when
    # they are created; every variant must give the oracle's result on a cloud large enough for several staging chunks
    from vtkcloudpoint_b200 import Context
    mx, my = synth.dbscan_cloud(0x5D, 60, n_total=200_003)
    want = oracle.dbscan(mx, my, 0.07, 7, 3, variant="grid")
    for env in ({"VPC_PDL": "0"}, {"VPC_COPY_THREADS": "0"}, {"VPC_COPY_THREADS": "5"}, {}):
        for k in ("VPC_PDL", "VPC_COPY_THREADS"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        with Context(0) as c:
            got = c.dbscan(mx, my, 0.07, 7, 3)
        assert got.cluster_amount == want[3], env
        np.testing.assert_array_equal(got.cluster_id, want[0], err_msg=str(env))
        np.testing.assert_array_equal(got.is_key, want[1], err_msg=str(env))
        np.testing.assert_array_equal(got.is_classed, want[2], err_msg=str(env))
