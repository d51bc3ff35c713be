"""BASELINE.json's full sizes on one GPU, checked through size-independent properties (the oracle cannot finish these in
seconds): C4's 100M-point cloud and a 10M-point C5 scene.
  * idempotence / order independence: clustering a cloud and clustering a PERMUTATION of it give the same partition,
    the same core flags, and the same numbering rule (ids ascend with the clusters' minimum core index);
  * the planted structure: every one of the grid x grid clusters of the recipe is found, noise stays unclustered in bulk;
  * consistency of the outputs: is_classed == (cluster_id != 0), ids fill 1..amount, every cluster has a core point."""
import numpy as np
import pytest
import torch

from vtkcloudpoint_b200 import synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _cloud(n, seed, chunk=5_000_000):
    grid = int(round((n * 0.784 / 40) ** 0.5))
    xs, ys = [], []
    for s in range(0, n, chunk):
        mx, my = synth.dbscan_cloud(seed, grid, n_total=n, start=s, count=min(chunk, n - s))
        xs.append(torch.from_numpy(mx).to(DEV)); ys.append(torch.from_numpy(my).to(DEV))
    return torch.cat(xs), torch.cat(ys), grid


def _properties(ctx, dx, dy, grid):
    n = dx.numel()
    cid, key, cls, amount = [t.clone() for t in ctx.dbscan_dev(dx, dy, 0.07, 7, 0)]
    k = int(amount.item())
    assert k >= grid * grid                                    # every planted cluster (a few noise clumps come on top)
    assert bool(((cid != 0) == (cls != 0)).all()) and int(cid.min()) == 0 and int(cid.max()) == k
    assert bool((cid[key != 0] > 0).all())
    has_core = torch.zeros(k + 1, dtype=torch.bool, device=DEV)
    has_core[cid[key != 0].long()] = True
    assert bool(has_core[1:].all())                            # ids fill 1..k and every cluster has a core point
    # numbering rule: the minimum core index per cluster ascends with the id (DBImproved.cs:93-110)
    first_core = torch.full((k + 1,), n, dtype=torch.int64, device=DEV)
    core_idx = torch.nonzero(key != 0).squeeze(1)
    first_core.scatter_reduce_(0, cid[core_idx].long(), core_idx, reduce="amin")
    assert bool((first_core[2:] > first_core[1:-1]).all())
    return cid, key, k


def _slab_vs_oracle(dx, dy, cid, key, eps=0.07, min_pts=7, frac=0.10):
    """a ~10M-point u-slab of the C4 cloud against the grid oracle on all host cores.  DBSCAN of a window equals
    DBSCAN of the whole cloud away from the window's edges: core flags of points at least eps inside, and the partition (canonicalised
    by minimum point index) of every cluster that keeps 2 eps clear of the edges, must agree bit for bit."""
    import os
    import oracle_py
    u = dx + dy
    sample = u[:: max(1, u.numel() // 1_000_000)]
    lo, hi = float(torch.quantile(sample, 0.5 - frac / 2)), float(torch.quantile(sample, 0.5 + frac / 2))
    idx = torch.nonzero((u >= lo) & (u < hi)).squeeze(1)
    wx, wy = dx[idx].cpu().numpy(), dy[idx].cpu().numpy()
    ocid, okey, _, _ = oracle_py.dbscan(wx, wy, eps, min_pts, 0, variant="grid", n_threads=os.cpu_count() or 8)
    gcid, gkey = cid[idx].cpu().numpy(), key[idx].cpu().numpy()
    wu = wx + wy
    inner = (wu >= lo + 1.01 * eps) & (wu < hi - 1.01 * eps)
    assert inner.sum() > 0.9 * len(wu) * (1 - 4 * eps / (hi - lo))
    np.testing.assert_array_equal(gkey[inner], okey[inner])                    # core flags
    # clusters (of either labelling) that come within 2 eps of the window's edges are excluded, the others must be the same sets
    near = ~((wu >= lo + 2.02 * eps) & (wu < hi - 2.02 * eps))
    bad_o = np.unique(ocid[near & (ocid > 0)]); bad_g = np.unique(gcid[near & (gcid > 0)])
    keep = ~np.isin(ocid, bad_o) & ~np.isin(gcid, bad_g)
    assert keep.sum() > 0.5 * len(wu)
    a, b = ocid[keep], gcid[keep]
    np.testing.assert_array_equal(a > 0, b > 0)                                # noise set
    nz = a > 0
    first_a = np.full(int(a.max()) + 1, -1, np.int64); first_b = np.full(int(b.max()) + 1, -1, np.int64)
    pos = np.flatnonzero(nz)
    first_a[a[pos][::-1]] = pos[::-1]; first_b[b[pos][::-1]] = pos[::-1]       # minimum member position per cluster
    np.testing.assert_array_equal(first_a[a[pos]], first_b[b[pos]])           # same partition, canonicalised by the minimum point index
    # numbering: ids ascend with the minimum core index in both (DBImproved.cs:93-110), so the rank order of the kept clusters agrees
    ua, ub = np.unique(a[nz]), np.unique(b[nz])
    assert len(ua) == len(ub)
    np.testing.assert_array_equal(first_a[ua], first_b[ub])


@pytest.mark.parametrize("n", [100_000_000])
def test_c4_100m_order_independence(ctx, n):
    dx, dy, grid = _cloud(n, 0xC4)
    cid, key, k = _properties(ctx, dx, dy, grid)
    _slab_vs_oracle(dx, dy, cid, key)
    # a permuted copy of the cloud: same partition, same core flags
    perm = torch.randperm(n, device=DEV, generator=torch.Generator(device=DEV).manual_seed(1))
    px, py = dx[perm].contiguous(), dy[perm].contiguous()
    cid2, key2, cls2, amount2 = ctx.dbscan_dev(px, py, 0.07, 7, 0)
    assert int(amount2.item()) == k
    assert bool((key2 == key[perm]).all())
    a, b = cid[perm].long(), cid2.long()                       # a bijection between the two numberings must exist
    fwd = torch.full((k + 1,), -1, dtype=torch.int64, device=DEV)
    fwd[a] = b
    assert bool((fwd[a] == b).all()) and int(fwd[0]) == 0
    back = torch.full((k + 1,), -1, dtype=torch.int64, device=DEV)
    back[b] = a
    assert bool((back[b] == a).all())
    del perm, px, py
    torch.cuda.empty_cache()


def test_c5_10m_pipeline_properties(ctx):
    import pipeline_ref
    import oracle_py
    from vtkcloudpoint_b200.pipeline import GpuPipelineBackend, run_pipeline
    n = 10_000_000
    grid = int(round((n * 0.784 / 40) ** 0.5))
    mx, my, xyz = pipeline_ref.scene(0xC5, grid, n)
    gx, gy = np.meshgrid(np.arange(grid), np.arange(grid), indexing="ij")
    cxyz, _ = oracle_py.polar_to_xyz(149.0 + 0.5 * gx.ravel(), 307.0 + 0.5 * gy.ravel(), np.full(grid * grid, 41.91), 149.0, 307.0)
    truth = pipeline_ref.truth_for(cxyz[:2])
    t = lambda v: torch.from_numpy(np.ascontiguousarray(v)).to(DEV)   # noqa: E731
    res = run_pipeline(GpuPipelineBackend(ctx), t(mx), t(my), t(xyz), 0, t(truth), eps=0.07, min_pts=7, radius_threshold=0.088, icp_e=1e-9,
                       icp_max_iters=10, match_distance=0.05)
    assert res.cluster_amount >= grid * grid
    # the stretched clusters (every 17th) are the ones the radius filter removes
    frac = float(res.filtered.sum()) / (grid * grid)
    assert 0.02 < frac < 0.09, frac
    st = res.icp_state.cpu().numpy()
    th = np.deg2rad(0.4)
    assert abs(st[0] - np.cos(th)) < 1e-4 and abs(st[3] - np.sin(th)) < 1e-4 and abs(st[9] - 0.011) < 5e-3 and abs(st[10] + 0.007) < 5e-3
    assert (res.matched >= 0).float().mean().item() > 0.98
    # centroids are means of members: recompute a sample of them on the host in list order, bit for bit
    cid = res.cluster_id.cpu().numpy()
    for c in (1, 2, res.cluster_amount // 2, res.cluster_amount):
        m = np.flatnonzero(cid == c)
        for f, v in enumerate((xyz[0], xyz[1], xyz[2], mx, my)):
            assert res.centres[f, c].item() == np.cumsum(v[m])[-1] / len(m)


def test_c3_full_size_vs_oracle(ctx):
    """Config C3 at full size: 100k source vs 1M target, 50 rounds with the convergence test off, against the
    grid oracle on all host cores: correspondences of the last round index-exact, R / T / SSE (hence RMSE) within 1e-6 relative
    (ICP.cs:224-250, :126-180)."""
    import os
    import oracle_py
    model, data, _, _ = synth.icp_clouds(0xC3, 1_000_000, 100_000)
    res = ctx.icp_rigid(model, data, -1.0, 50)
    Ro, To, itd, sse, oo = oracle_py.icp_rigid(model, data, -1.0, 50, use_grid=True, n_threads=os.cpu_count() or 8)
    assert res.iters_done == itd == 50
    np.testing.assert_array_equal(res.order_last, oo)
    assert np.abs(res.R - Ro).max() <= 1e-6 * np.abs(Ro).max()
    assert np.abs(res.T - To).max() <= 1e-6 * np.abs(To).max()
    assert abs(res.sse_last - sse) <= 1e-6 * sse
    # and the device-resident entry point the bench times
    dm, dd = torch.from_numpy(model).to(DEV), torch.from_numpy(data).to(DEV)
    ctx.icp_set_model_dev(dm)
    state, order = ctx.icp_rigid_dev(dd, -1.0, 50)
    st = state.cpu().numpy()
    np.testing.assert_array_equal(order.cpu().numpy(), oo)
    assert abs(st[12] - sse) <= 1e-6 * sse and int(st[13]) == 50
