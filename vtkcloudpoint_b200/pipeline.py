"""The reference's whole clustering-to-matching pipeline on N GPUs (BASELINE.json config C5):

  DBSCAN over the scan points          Clustering.DoClusteringBtn_Click -> DBImproved.dbscan   (FrmMain.cs:1214.., DBImproved.cs:91)
  per-cluster centroids + circles      Tools.GetClusList, Tools.getCircles                      (Tools.cs:162-195, 394-409; FrmMain.cs:1521-1540)
  radius filter                        MainForm.FilterClustersByRadius (+ MCC threshold 0.088)  (FrmMain.cs:1905-1920, MCC.Designer.cs:71)
  drop the filtered clusters' centres  Tools.removeFilterPointFromClustering                    (Tools.cs:70-74, FrmMain.cs:3744-3745)
  ICP of the centres (z = 0) to truth  hand-written ICP.go_hell_ICP on (X, Y, 0)                (ICP.cs:18-181; z = 0 as in Tools.cs:701)
  thresholded match                    MainForm.RecorrectMatchingPtsByDistance                  (FrmMain.cs:3588-3618)

One process per GPU.  DBSCAN is exact across the GPUs (distributed.dbscan_slabs).  The statistics shard by CLUSTER: cluster
ids are cut into `world` contiguous ranges, every non-noise point travels once to the rank owning its cluster (all_to_all, in
rawData order, which the centroid sums and the gift-wrapping hull depend on), each rank computes centroids and circles for its
clusters, and the per-cluster results (a few doubles each) are all-gathered.  ICP then runs with the truth points sharded and
the centres replicated (distributed.icp_rigid_sharded).  With world == 1 the same code runs without collectives.

The compute steps are injected (`PipelineBackend`): libvpc.so through the C ABI in the product, a checker built on the CPU
oracle in the gloo tests.  There is no CPU fallback in this module.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.distributed as dist

from .distributed import _all_gather_var, _all_to_all_var, _world, dbscan_slabs, icp_rigid_sharded


class GpuPipelineBackend:
    """The product backend: every step is a libvpc.so call on device tensors."""

    def __init__(self, ctx):
        from .distributed import GpuBackend, GpuIcpBackend
        self.ctx = ctx
        self.db = GpuBackend(ctx)
        self.icp = GpuIcpBackend(ctx)

    def dbscan_single(self, mx, my, eps, min_pts):
        cid, key, cls, amount = self.ctx.dbscan_dev(mx, my, eps, min_pts, 0)
        return cid, key, int(amount.item())

    def cluster_stats(self, cid_local, k_local, vals5):
        """cid_local in 0..k_local, vals5 = [X, Y, Z, motor_x, motor_y] planar.  -> means [5,k+1], counts, circle [3,k+1], status"""
        members, offsets = self.ctx.cluster_groups_dev(cid_local, k_local)
        means, counts = self.ctx.cluster_means_ordered_dev(members, offsets, k_local, vals5)
        circ, status = self.ctx.cluster_circles_dev(members, offsets, k_local, vals5[0], vals5[1])     # is3D = true: the hull is on X, Y
        return means, counts, circ, status

    def radius_flag(self, circ, status, k, thr):
        return self.ctx.radius_filter_dev(circ[2].contiguous(), status, k, thr)

    def match_within(self, truth_planar, centres_planar, match_distance):
        self.ctx.icp_set_model_dev(truth_planar)
        n = centres_planar.shape[1]
        mid = torch.empty(n, dtype=torch.int32, device=centres_planar.device)
        d = torch.empty(n, dtype=torch.float64, device=centres_planar.device)
        s = torch.cuda.current_stream(centres_planar.device).cuda_stream
        self.ctx._check(self.ctx._lib.vpc_match_within_dev(self.ctx._h, centres_planar.data_ptr(), n, float(match_distance), mid.data_ptr(), d.data_ptr(), s))
        return mid, d


@dataclass
class PipelineResult:
    cluster_id: torch.Tensor        # [n_local] int32, this rank's chunk (global ids, DBImproved numbering)
    cluster_amount: int
    centres: torch.Tensor           # [5, k+1] means of X, Y, Z, motor_x, motor_y per cluster id (NaN: no members) -- replicated
    counts: torch.Tensor            # [k+1]
    circle: torch.Tensor            # [3, k+1] cx, cy, radius of the 3-D circle (radius -1: no circle)
    circle_status: torch.Tensor     # [k+1]
    filtered: torch.Tensor          # [k+1] uint8, 1 = radius > threshold (MainForm.filterID)
    kept_ids: torch.Tensor          # cluster ids whose centres go to the matching
    icp_state: torch.Tensor         # f64[16] = R[9] T[3] sse iters converged 0
    icp_order: torch.Tensor         # nearest truth index per kept centre in the last round
    matched: torch.Tensor | None    # truth index per kept centre after the final transform, -1 = farther than match_distance


def trans_points(state, pts):
    """ICP.TransPoint (ICP.cs:195-219) in its operation order, on planar [3, n] tensors."""
    R, T = state[:9], state[9:12]
    x, y, z = pts[0], pts[1], pts[2]
    return torch.stack([(((0.0 + R[0] * x) + R[1] * y) + R[2] * z) + T[0],
                        (((0.0 + R[3] * x) + R[4] * y) + R[5] * z) + T[1],
                        (((0.0 + R[6] * x) + R[7] * y) + R[8] * z) + T[2]])


def run_pipeline(backend, mx, my, xyz, gidx0: int, truth_xy, *, eps: float, min_pts: int, radius_threshold: float, icp_e: float,
                 icp_max_iters: int, match_distance: float | None = None, group=None) -> PipelineResult:
    """mx, my: this rank's chunk of the scan (float64 device tensors; element i has global index gidx0 + i, the chunks of all
    ranks tile the cloud in rank order); xyz: planar [3, n_local] Cartesian coordinates of the same points; truth_xy: planar
    [2, m] truth positions, identical on every rank."""
    rank, world = _world(group)
    dev = mx.device
    n = mx.numel()

    # ---- 1. DBSCAN, exact across the ranks -------------------------------------------------------------------------
    if world == 1:
        cid, _, amount = backend.dbscan_single(mx, my, eps, min_pts)
    else:
        cid, _, _, amount = dbscan_slabs(backend.db, mx, my, gidx0, eps, min_pts, 0, group=group)
    k = int(amount)
    k1 = k + 1

    # ---- 2. statistics, sharded by cluster id ----------------------------------------------------------------------------
    lo = [1 + (k * r) // world for r in range(world + 1)]            # rank r owns ids lo[r] .. lo[r+1]-1
    vals5 = torch.stack([xyz[0], xyz[1], xyz[2], mx, my])
    if world == 1:
        cid_l, vals_l = cid, vals5
    else:
        sel = torch.nonzero(cid > 0).squeeze(1)                        # ascending local order = ascending global order
        c_sel = cid[sel]
        bounds = torch.tensor(lo[1:-1], dtype=torch.int32, device=dev)
        dest = torch.searchsorted(bounds, c_sel, right=True)
        order = torch.sort(dest, stable=True).indices                   # stable: rawData order survives inside a destination
        sel, c_sel, dest = sel[order], c_sel[order], dest[order]
        send_counts = torch.bincount(dest, minlength=world).to(torch.int64)
        cols = [c_sel.contiguous()] + [vals5[f][sel].contiguous() for f in range(5)]
        recv, _ = _all_to_all_var(cols, send_counts, group)             # chunks arrive in rank order = global index order
        cid_l = (recv[0] - (lo[rank] - 1)).to(torch.int32).contiguous()
        vals_l = torch.stack(recv[1:]).contiguous()
    k_loc = lo[rank + 1] - lo[rank]
    means_l, counts_l, circ_l, status_l = backend.cluster_stats(cid_l.contiguous(), k_loc, vals_l.contiguous())
    nan = float("nan")
    if world == 1:
        means, counts, circ, status = means_l, counts_l, circ_l, status_l
    else:
        means = torch.full((5, k1), nan, dtype=torch.float64, device=dev)
        counts = torch.zeros(k1, dtype=torch.int32, device=dev)
        circ = torch.full((3, k1), nan, dtype=torch.float64, device=dev)
        circ[2] = -1.0
        status = torch.zeros(k1, dtype=torch.int32, device=dev)
        packed = torch.cat([means_l[:, 1:], circ_l[:, 1:], counts_l[1:].to(torch.float64)[None], status_l[1:].to(torch.float64)[None]]).T.contiguous()
        allp = _all_gather_var(packed.reshape(-1), group).reshape(-1, 10)   # rank order = cluster id order
        assert allp.shape[0] == k
        means[:, 1:] = allp[:, 0:5].T
        circ[:, 1:] = allp[:, 5:8].T
        counts[1:] = allp[:, 8].to(torch.int32)
        status[1:] = allp[:, 9].to(torch.int32)

    # ---- 3. radius filter, centres that go on --------------------------------------------------------------------------
    filtered = backend.radius_flag(circ, status, k, radius_threshold)
    keep = (counts > 0) & (filtered == 0)
    keep[0] = False
    kept_ids = torch.nonzero(keep).squeeze(1).to(torch.int32)
    cen = means[:, kept_ids.long()]
    centres_planar = torch.stack([cen[0], cen[1], torch.zeros_like(cen[0])]).contiguous()      # z = 0 (Tools.cs:701)

    # ---- 4. ICP: truth sharded, centres replicated -----------------------------------------------------------------------
    m = truth_xy.shape[1]
    a, b = (m * rank) // world, (m * (rank + 1)) // world
    truth_planar = torch.stack([truth_xy[0], truth_xy[1], torch.zeros_like(truth_xy[0])]).contiguous()
    if 0 < m < world:
        # every rank must own at least one truth point (an empty shard would leave its rank out of the collectives); all ranks see
        # the same m, so all raise together
        raise ValueError(f"the truth set ({m} points) is smaller than the number of ranks ({world})")
    if centres_planar.shape[1] == 0 or m == 0:
        state = torch.zeros(16, dtype=torch.float64, device=dev)
        order = torch.empty(0, dtype=torch.int32, device=dev)
        matched = None
    else:
        state, order = icp_rigid_sharded(backend.icp, truth_planar[:, a:b].contiguous(), a, centres_planar, icp_e, icp_max_iters, group=group)
        matched = None
        if match_distance is not None:
            moved = trans_points(state, centres_planar).contiguous() if float(state[13]) > 0 and bool(torch.any(state[:9] != 0)) else centres_planar
            matched, _ = backend.match_within(truth_planar, moved, match_distance)
    return PipelineResult(cid, k, means, counts, circ, status, filtered, kept_ids, state, order, matched)
