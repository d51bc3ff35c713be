"""TEST INFRASTRUCTURE: the C5 pipeline evaluated by the CPU oracle on the WHOLE cloud in one process -- the expected result
for vtkcloudpoint_b200.pipeline.run_pipeline at any world size -- and the small synthetic scene the tests feed it."""
import numpy as np

import oracle_py
from vtkcloudpoint_b200 import synth


def scene(seed: int, grid: int, n_total: int, fat_every: int = 17):
    """C5 recipe at a chosen size: grid x grid clusters + noise in motor units; XYZ by the import formulas; a few clusters are
    stretched so that their bounding circle exceeds the 0.088 threshold; truth = the cluster centres, rotated and shifted."""
    mx, my = synth.dbscan_cloud(seed, grid, n_total=n_total)
    # stretch every fat_every-th cluster along x (done on the generated points: those nearest to such a centre)
    cx = np.round((mx - 149.0) / 0.5).astype(np.int64)
    cy = np.round((my - 307.0) / 0.5).astype(np.int64)
    near = (np.abs(mx - (149.0 + 0.5 * cx)) < 0.06) & (np.abs(my - (307.0 + 0.5 * cy)) < 0.06)
    fat = near & ((cx + grid * cy) % fat_every == 0)
    mx = np.where(fat, 149.0 + 0.5 * cx + (mx - (149.0 + 0.5 * cx)) * 2.6, mx)
    dist = 41.91 + 0.004 * (synth.uniform(seed, 40, np.arange(len(mx), dtype=np.uint64)) - 0.5)   # a flat target: range spread << pitch
    xyz, keep = oracle_py.polar_to_xyz(mx, my, dist, 149.0, 307.0)
    assert keep.all()
    return mx, my, np.ascontiguousarray(xyz)


def truth_for(centres_xy: np.ndarray, seed: int = 5):
    """Truth points = the (X, Y) centres moved by a small rigid motion plus noise, shuffled."""
    th = np.deg2rad(0.4)
    R = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    t = R @ centres_xy + np.array([[0.011], [-0.007]])
    rng = np.random.default_rng(seed)
    t = t + rng.normal(0, 2e-3, t.shape)          # measurement noise: the SSE at convergence is a real number, not rounding dust
    return np.ascontiguousarray(t[:, rng.permutation(t.shape[1])])


def run(mx, my, xyz, truth_xy, eps, min_pts, radius_threshold, icp_e, icp_max_iters, match_distance=None):
    cid, key, cls, amount = oracle_py.dbscan(mx, my, eps, min_pts, 0, variant="grid")
    st = oracle_py.cluster_stats(cid, amount, xyz, mx, my, circles3d=True, circles2d=False)
    filtered = (st["status3d"] == 1) & (st["circle3d"][2] > radius_threshold)
    filtered[0] = False
    keep = (st["counts"] > 0) & ~filtered
    keep[0] = False
    kept = np.flatnonzero(keep).astype(np.int32)
    cen = st["means"][:, kept]
    centres = np.stack([cen[0], cen[1], np.zeros(len(kept))])
    truth = np.stack([truth_xy[0], truth_xy[1], np.zeros(truth_xy.shape[1])])
    R, T, iters, sse, order = oracle_py.icp_rigid(truth, centres, icp_e, icp_max_iters)
    out = dict(cluster_id=cid, amount=amount, means=st["means"], counts=st["counts"], circle=st["circle3d"], status=st["status3d"],
               filtered=filtered.astype(np.uint8), kept=kept, R=R, T=T, iters=iters, sse=sse, order=order)
    if match_distance is not None:
        moved = oracle_py.trans_points(centres, R, T) if np.any(R != 0) else centres
        out["matched"], _ = oracle_py.match_within(truth, moved, match_distance)
    return out


def check(res, ref, a, b, rtol=1e-6):
    """res: PipelineResult of one rank holding points [a, b); ref: run(...) on the whole cloud."""
    t = lambda x: x.cpu().numpy()   # noqa: E731
    assert res.cluster_amount == ref["amount"]
    np.testing.assert_array_equal(t(res.cluster_id), ref["cluster_id"][a:b])
    np.testing.assert_array_equal(t(res.counts), ref["counts"])
    np.testing.assert_array_equal(t(res.circle_status), ref["status"])
    np.testing.assert_array_equal(t(res.centres)[:, 1:].view(np.int64), ref["means"][:, 1:].view(np.int64))     # bit patterns
    np.testing.assert_array_equal(t(res.circle)[:, 1:].view(np.int64), ref["circle"][:, 1:].view(np.int64))
    np.testing.assert_array_equal(t(res.filtered)[1:], ref["filtered"][1:])
    np.testing.assert_array_equal(t(res.kept_ids), ref["kept"])
    st = t(res.icp_state)
    assert int(st[13]) == ref["iters"], (st[13], ref["iters"])
    np.testing.assert_array_equal(t(res.icp_order), ref["order"])
    assert np.abs(st[:9].reshape(3, 3) - ref["R"]).max() <= rtol and np.abs(st[9:12] - ref["T"]).max() <= rtol * max(1.0, np.abs(ref["T"]).max())
    assert abs(st[12] - ref["sse"]) <= rtol * max(ref["sse"], 1e-30)
    if "matched" in ref and res.matched is not None:
        np.testing.assert_array_equal(t(res.matched), ref["matched"])
