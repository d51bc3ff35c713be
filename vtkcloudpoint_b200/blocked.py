"""The reference's blocked ("分块") multithreaded clustering, SURVEY.md 8a rows a5-a8: a thin Python face of the two C-ABI calls
that implement it on the device (csrc/blocked.cuh, csrc/blocked_api.cuh):

  cluster_blocked   MainForm.getClusterFromMotor -> DoWork3 / StartCode -> CompleteWork3   FrmMain.cs:1214-1291, 1340-1361, 2782-2794, 1432-1544
                    = Context.dbscan_blocked_ref  (vpc_dbscan_blocked_ref_ex)
  merge_clusters    Clustering.MergeBtn_Click: Tools.GetClusList -> MergeIDByDistance -> refreshCensAndClusByDictionary
                    Clustering.cs:141-153, Tools.cs:162-195, 580-621, 521-572
                    = Context.merge_ids_by_distance  (vpc_merge_ids_by_distance)

There is ONE implementation of this flow in the product (the library); the literal List-based restatement it is checked against lives
in the literal restatement in oracle/ (part 4, blocked clustering).  Nothing here computes: the functions only arrange arrays for the calls.
"""
from __future__ import annotations

import numpy as np


def cluster_blocked(ctx, mx, my, eps: float, min_pts: int, pts_in_cell: int) -> dict:
    """Returns the dict of Context.dbscan_blocked_ref (cluster_id per input point, cluster_sum = MainForm.clusterSum, clusForMerge
    as merge_order / merge_cid, del_sum, rows, cols, n_unassigned, n_shared)."""
    return ctx.dbscan_blocked_ref(np.ascontiguousarray(mx, np.float64), np.ascontiguousarray(my, np.float64), eps, min_pts, pts_in_cell)


def merge_clusters(ctx, blocked_result: dict, points_xyz, mx, my, thre: float) -> dict:
    """The centroid merge on the result of cluster_blocked.  points_xyz: (n, 3) or planar (3, n) X, Y, Z of the INPUT points; the
    per-entry arrays of the clusForMerge list are gathered here (list order = the order the C# sums centroids in)."""
    xyz = np.asarray(points_xyz, np.float64)
    if xyz.shape[0] != 3 or xyz.ndim != 2 or xyz.shape[1] == 3:
        xyz = xyz.T
    order = blocked_result["merge_order"]
    return ctx.merge_ids_by_distance(blocked_result["merge_cid"], np.ascontiguousarray(xyz[:, order]), np.asarray(mx, np.float64)[order],
                                     np.asarray(my, np.float64)[order], blocked_result["cluster_sum"], thre)
