"""World-size 2 and 3 gloo tests of the multi-GPU HOST logic (slab partition, halo exchange, cross-slab
union-find merge, global numbering, result return) with the checker-backed CPU backend.  The result must
equal DBImproved.dbscan on the whole cloud (oracle), bit for bit."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _cloud(case):
    rng = np.random.default_rng(100 + case)
    if case == 0:       # long diagonal chains crossing every slab boundary + noise
        t = np.linspace(0, 4, 1500)
        pts = np.vstack([np.c_[t, 0.3 * np.sin(3 * t)] + rng.normal(0, 0.01, (1500, 2)),
                         np.c_[t, 2 - t] + rng.normal(0, 0.01, (1500, 2)), rng.uniform(-0.5, 4.5, (1200, 2))])
    elif case == 1:     # blobs, many sitting on the u-quantiles
        c = rng.uniform(0, 3, (40, 2))
        pts = np.vstack([c[rng.integers(0, 40, 2500)] + rng.normal(0, 0.03, (2500, 2)), rng.uniform(0, 3, (800, 2))])
    else:               # lattice: exact ties on |dx|+|dy| == eps, duplicates, a NaN and an inf
        pts = np.c_[rng.integers(0, 40, 3000) * 0.05, rng.integers(0, 40, 3000) * 0.05].astype(np.float64)
        pts[17, 0] = np.nan
        pts[333, 1] = np.inf
    perm = rng.permutation(len(pts))
    return pts[perm, 0].copy(), pts[perm, 1].copy()


def _worker(rank, world, port, case, eps, min_pts, cf0, presplit=False):
    for p in (str(ROOT), str(ROOT / "oracle"), str(ROOT / "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle_py
        from dist_cpu_backend import CpuCheckerBackend
        from vtkcloudpoint_b200.distributed import dbscan_slabs
        mx, my = _cloud(case)
        n = len(mx)
        splitters = None
        if presplit:
            # pre-cut cloud: order the points by slab (rank r = the r-th u-quantile band), non-finite ones anywhere
            u = mx + my
            fin = np.isfinite(u)
            qs = np.quantile(u[fin], [j / world for j in range(1, world)]) if world > 1 else np.empty(0)
            band = np.where(fin, np.searchsorted(qs, u, side="right"), 0)
            order = np.argsort(band, kind="stable")
            mx, my, band = mx[order], my[order], band[order]
            bounds = [int(np.searchsorted(band, r, side="left")) for r in range(world)] + [n]
            splitters = torch.from_numpy(np.asarray(qs, dtype=np.float64))
        else:
            bounds = [n * r // world for r in range(world + 1)]
        a, b = bounds[rank], bounds[rank + 1]
        stats = {}
        cid, key, cls, amount = dbscan_slabs(CpuCheckerBackend(), torch.from_numpy(mx[a:b].copy()), torch.from_numpy(my[a:b].copy()),
                                             a, eps, min_pts, cf0, stats=stats, splitters=splitters)
        ocid, okey, ocls, oamount = oracle_py.dbscan(mx, my, eps, min_pts, cf0, variant="grid")
        assert amount == oamount, (amount, oamount)
        np.testing.assert_array_equal(key.numpy(), okey[a:b])
        np.testing.assert_array_equal(cid.numpy(), ocid[a:b])
        np.testing.assert_array_equal(cls.numpy(), ocls[a:b])
        if world > 1:
            assert stats["n_local"] >= stats["n_owned"] > 0
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,case,eps,min_pts", [(2, 0, 0.06, 4), (2, 1, 0.05, 5), (2, 2, 0.05, 3), (3, 0, 0.06, 4), (3, 1, 0.08, 6),
                                                     (3, 2, 0.1, 4), (1, 1, 0.05, 5)])
def test_dbscan_slabs_matches_whole_cloud(world, case, eps, min_pts):
    mp.spawn(_worker, args=(world, _free_port(), case, eps, min_pts, 7), nprocs=world, join=True)


@pytest.mark.parametrize("world,case,eps,min_pts", [(2, 0, 0.06, 4), (3, 1, 0.05, 5), (3, 2, 0.1, 4)])
def test_dbscan_slabs_presplit(world, case, eps, min_pts):
    mp.spawn(_worker, args=(world, _free_port(), case, eps, min_pts, 0, True), nprocs=world, join=True)


def _icp_worker(rank, world, port, e, max_iters):
    for p in (str(ROOT), str(ROOT / "oracle"), str(ROOT / "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle_py
        from dist_cpu_backend import CpuCheckerIcpBackend
        from vtkcloudpoint_b200 import synth
        from vtkcloudpoint_b200.distributed import icp_rigid_sharded
        model, data, R, T = synth.icp_clouds(0xC3, 30_000, 3_000)
        model[:, 7] = model[:, 20_007]                      # a duplicate across shards: the tie must go to the lower global index
        m = model.shape[1]
        a, b = m * rank // world, m * (rank + 1) // world
        state, order = icp_rigid_sharded(CpuCheckerIcpBackend(), torch.from_numpy(model[:, a:b].copy()), a, torch.from_numpy(data), e, max_iters)
        st = state.numpy()
        Ro, To, it, sse, oo = oracle_py.icp_rigid(model, data, e, max_iters)
        assert int(st[13]) == it, (st[13], it)
        np.testing.assert_array_equal(order.numpy(), oo)
        assert np.abs(st[:9].reshape(3, 3) - Ro).max() < 1e-6 and np.abs(st[9:12] - To).max() < 1e-6 and abs(st[12] - sse) <= 1e-6 * max(sse, 1e-30)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,e,max_iters", [(2, -1.0, 4), (3, 1e-4, 12), (1, -1.0, 3)])
def test_icp_sharded_matches_whole_model(world, e, max_iters):
    mp.spawn(_icp_worker, args=(world, _free_port(), e, max_iters), nprocs=world, join=True)
