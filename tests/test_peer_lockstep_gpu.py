"""The peer-memory multi-GPU paths (csrc/slab.cuh, icp_dist.cuh) on ONE GPU: `world` ranks emulated phase by phase, compared
with the CPU oracle on the whole cloud.  Covers everything but the physical NVLink hop (tests/test_multi_gpu.py does that)."""
import numpy as np
import pytest

from vtkcloudpoint_b200 import synth

from peer_helpers import canon, run_icp_lockstep, run_slabs_lockstep

pytestmark = pytest.mark.gpu


def _check_slabs(oracle, fx, fy, world, eps, min_pts, cf0=0, variant="grid", **kw):
    cid, key, cls, amount, status, errs, order = run_slabs_lockstep(fx, fy, world, eps, min_pts, cf0, **kw)
    assert errs == [0] * world and all(int(s[1]) == 0 for s in status), (errs, [s[:6] for s in status])
    # the oracle clusters the cloud in the slab order the ranks hold it in (global index = position in that order)
    ocid, okey, ocls, oamount = oracle.dbscan(fx[order], fy[order], eps, min_pts, cf0, variant=variant)
    assert amount == oamount
    np.testing.assert_array_equal(key[order], okey)
    np.testing.assert_array_equal(cls[order], ocls)
    np.testing.assert_array_equal(cid[order], ocid)
    return status


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_slabs_c2_recipe(oracle, world):
    fx, fy = synth.dbscan_cloud(0xC2, 44, n_total=100_000)
    status = _check_slabs(oracle, fx, fy, world, 0.07, 7, cf0=5)
    assert all(int(s[5]) == 2 for s in status)          # two steps ran: epochs advance in lockstep


@pytest.mark.parametrize("world", [2, 4])
def test_slabs_chain_through_all_slabs(oracle, world):
    # one dense diagonal band crossing every slab boundary: the cluster key must travel through all ranks (merge chains)
    rng = np.random.default_rng(3)
    t = rng.uniform(0, 10, 60_000)
    fx, fy = t + rng.normal(0, 0.01, t.size), t + rng.normal(0, 0.01, t.size)
    noise = rng.uniform(0, 10, (2, 5000))
    fx, fy = np.concatenate([fx, noise[0]]), np.concatenate([fy, noise[1]])
    _check_slabs(oracle, fx, fy, world, 0.05, 5)


@pytest.mark.parametrize("world", [2, 3])
def test_slabs_lattice_ties_and_nonfinite(oracle, world):
    # lattice: pairs exactly at distance eps, points exactly on slab boundaries; NaN / inf owned points are noise (DBImproved.cs:41)
    rng = np.random.default_rng(11)
    n = 20_000
    fx, fy = rng.integers(0, 120, n) * 0.25, rng.integers(0, 120, n) * 0.25
    fx[::997] = np.nan; fy[5::1013] = np.inf; fx[7::1999] = -np.inf
    for eps, mp in ((0.25, 3), (0.5, 6)):
        _check_slabs(oracle, fx, fy, world, eps, mp, variant="grid")


def test_slabs_small_literal(oracle):
    rng = np.random.default_rng(5)
    fx, fy = rng.uniform(0, 4, 3000), rng.uniform(0, 1, 3000)
    _check_slabs(oracle, fx, fy, 3, 0.06, 3, cf0=2, variant="literal")


def test_slabs_overflow_is_reported(oracle):
    fx, fy = synth.dbscan_cloud(0xC2, 44, n_total=100_000)
    cid, key, cls, amount, status, errs, order = run_slabs_lockstep(fx, fy, 2, 0.07, 7, cap_frac=0.0)   # capacities 1024: too small
    assert all(e & 2 for e in errs) or any(int(s[1]) & 2 for s in status)


def _check_icp(oracle, model, data, world, mode, iters, e=-1.0):
    outs, errs = run_icp_lockstep(model, data, world, mode, e, iters)
    assert errs == [0] * world
    Ro, To, itd, sse, oo = oracle.icp_rigid(model, data, e, iters, use_grid=model.shape[1] > 3000)
    for st, order in outs:
        assert int(st[13]) == itd
        np.testing.assert_array_equal(order, oo)
        assert np.abs(st[:9].reshape(3, 3) - Ro).max() < 1e-6
        assert np.abs(st[9:12] - To).max() < 1e-6 * max(1.0, np.abs(To).max())
        assert abs(st[12] - sse) <= 1e-6 * sse + 1e-18 * data.shape[1]   # north_star: transform and RMSE within 1e-6 relative (absolute floor: an exact fit leaves rounding noise only)
    for st, order in outs[1:]:                                      # the replicated solve is bit-identical on every rank
        np.testing.assert_array_equal(st, outs[0][0])


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_icp_dist_c3_recipe(oracle, world, mode):
    model, data, _, _ = synth.icp_clouds(0xC3, 30_000, 4_000)
    _check_icp(oracle, model, data, world, mode, 6)


@pytest.mark.parametrize("mode", [0, 1])
def test_icp_dist_convergence_stops_all_ranks(oracle, mode):
    model, data, _, _ = synth.icp_clouds(0xC3, 5_000, 600, jitter=0.0)
    _check_icp(oracle, model, data, 3, mode, 40, e=1e-9)


@pytest.mark.parametrize("mode", [0, 1])
def test_icp_dist_ties_and_nonfinite(oracle, mode):
    # lattice model: equidistant candidates in different shards resolve to the lowest GLOBAL index (ICP.cs:240); a NaN data point pins
    # its match to model[0] (ICP.cs:233-244)
    rng = np.random.default_rng(2)
    g = np.stack(np.meshgrid(np.arange(12.0), np.arange(12.0), np.arange(12.0), indexing="ij")).reshape(3, -1)
    model = np.ascontiguousarray(g[:, rng.permutation(g.shape[1])])
    data = np.ascontiguousarray(rng.integers(0, 23, (3, 700)) * 0.5)       # half-integer points: 2-, 4- and 8-way ties
    data[0, 13] = np.nan
    outs, errs = run_icp_lockstep(model, data, 4, mode, -1.0, 1)
    assert errs == [0] * 4
    order, _ = oracle.closest_point_set(model, data, "literal")
    for st, o in outs:
        np.testing.assert_array_equal(o, order)
