"""Host-side mirror of the reference's blocked ("分块") multithreaded clustering, SURVEY.md 8a rows a5-a8.

The C# keeps this logic on the host and calls DBImproved.dbscan from it; a drop-in replaces only those calls.
This module restates the host logic in Python over flat arrays so that the whole flow can be driven -- and
tested -- above the C ABI:

  partition_cells   MainForm.getClusterFromMotor      FrmMain.cs:1214-1291 (+ Tools.getListByScale2, Tools.cs:510-513)
  (per-cell DBSCAN) MainForm.DoWork3 / StartCode      FrmMain.cs:1340-1361, 2782-2794  -> ONE vpc_dbscan_l1_2d_cells call
  complete_work3    MainForm.CompleteWork3            FrmMain.cs:1432-1544 (renumber, drop clusters of <= 3 points incl. its
                                                      off-by-one, re-cluster all noise globally with a seeded cf)
  centroids         Tools.GetClusList                 Tools.cs:162-195
  merge_ids_by_distance / refresh_by_dictionary       Tools.cs:580-621, 521-572

The clustering engine is injected (`dbscan`, `dbscan_cells` callables): vtkcloudpoint_b200.Context in the
product, the CPU oracle in the tests.  Two things the C# leaves to chance are pinned here and documented:
List.Sort is unstable (ties are ordered by original position here) and the statics sumPts/threadCount/clusterSum
are updated without synchronisation by the pool threads (FrmMain.cs:2787-2789; summed deterministically here).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np


@dataclass
class CellPartition:
    order: np.ndarray          # point indices grouped by cell (cell-major), int64
    offsets: np.ndarray        # CSR offsets into `order`, int64 [rows*cols + 1]
    rows: int
    cols: int
    dropped: np.ndarray        # indices that fall into no cell (strict lower bounds / ties at the cell-0 cut)


def partition_cells(mx: np.ndarray, my: np.ndarray, pts_in_cell: int, argsort=None) -> CellPartition:
    """MainForm.getClusterFromMotor, FrmMain.cs:1224-1285.  argsort(key) -> stable ascending permutation; the product passes
    Context.argsort_f64 (the device radix sort, vpc_argsort_f64_dev), the default is NumPy's stable sort."""
    n = len(mx)
    if n == 0:
        raise ValueError("empty cloud (the C# returns early, FrmMain.cs:1228)")
    x_min, y_min, x_max, y_max = mx.min(), my.min(), mx.max(), my.max()                 # :1224-1227
    key = np.maximum(mx - x_min, my - y_min)                                            # :1232-1233
    srt = np.asarray(argsort(key), np.int64) if argsort is not None else np.argsort(key, kind="stable")   # :1229 (List.Sort: ties pinned to index order)
    cell0 = srt[:pts_in_cell]                                                           # :1253
    cell_x = mx[cell0].max() - x_min                                                    # :1255
    cell_y = my[cell0].max() - y_min                                                    # :1256
    if not (cell_x > 0 and cell_y > 0):
        raise ValueError("degenerate first cell: the C# divides by zero here (FrmMain.cs:1257-1258)")
    rows = int((y_max - y_min) / cell_y) + 1                                            # :1257
    cols = int((x_max - x_min) / cell_x) + 1                                            # :1258
    sx, sy = mx[srt], my[srt]                                                           # FindAll scans the SORTED rawData
    groups = [cell0]
    for p in range(rows):                                                               # :1262-1285
        for q in range(cols):
            if p == 0 and q == 0:
                continue
            lo_x, lo_y = x_min + q * cell_x, y_min + p * cell_y
            hi_x = x_max if q == cols - 1 else x_min + (q + 1) * cell_x
            hi_y = y_max if p == rows - 1 else y_min + (p + 1) * cell_y
            m = (sx > lo_x) & (sy > lo_y) & (sx <= hi_x) & (sy <= hi_y)                 # Tools.cs:512
            groups.append(srt[m])
    offsets = np.zeros(len(groups) + 1, np.int64)
    np.cumsum([len(g) for g in groups], out=offsets[1:])
    order = np.concatenate(groups).astype(np.int64)
    taken = np.zeros(n, bool)
    taken[order] = True
    return CellPartition(order, offsets, rows, cols, np.flatnonzero(~taken))


@dataclass
class BlockedResult:
    cluster_id: np.ndarray                 # final Point3D.clusterId per ORIGINAL point (dropped points keep 0)
    cluster_amount: int                    # MainForm.clusterSum after CompleteWork3 (:1538)
    merge_order: np.ndarray                # clusForMerge as point indices, in its final order (:1517-1520)
    del_sum: int
    cluster_sum_cells: int                 # clusterSum before the merge = 1 + sum of per-cell amounts (:1346, :2789)
    partition: CellPartition
    centers: np.ndarray = field(default=None)     # [k,3] mean X,Y,Z per non-empty cluster (Tools.cs:192)
    centers2d: np.ndarray = field(default=None)   # [k,2] mean motor_x, motor_y (Tools.cs:193)
    center_ids: np.ndarray = field(default=None)  # clusId of each centre


def complete_work3(part: CellPartition, local_id: np.ndarray, per_cell_amount: np.ndarray, mx, my, eps, min_pts, dbscan):
    """MainForm.CompleteWork3, FrmMain.cs:1443-1520.  local_id: cell-local cluster ids in `part.order` layout."""
    cid = np.asarray(local_id, np.int64).copy()         # clusterId of order[k]
    cluster_sum = 1 + int(np.sum(per_cell_amount))       # :1346 clusterSum = 1; :2789 += clusterAmount
    id_now, del_sum = 0, 0
    merge = []                                           # clusForMerge: positions into part.order
    for c in range(len(part.offsets) - 1):
        a, b = int(part.offsets[c]), int(part.offsets[c + 1])
        if a == b:
            continue                                     # :1448
        pos = a + np.argsort(cid[a:b], kind="stable")    # :1449-1459 sort the cell by id
        id_last = int(cid[pos[0]])                       # :1460
        if id_last != 0:
            id_now += 1
            clus_len = 1                                 # :1461-1465 (the j = 0 pass below bumps it to 2: off-by-one of the C#)
        else:
            clus_len = 0
        for k in pos:                                    # :1470
            i_d = int(cid[k])
            if i_d == 0:
                merge.append(k)                          # :1475
                continue
            if i_d != id_last:                           # :1479
                if clus_len <= 3 and id_last != 0:       # :1481 cluster too small: zero it, do not advance the id
                    del_sum += 1
                    for t in range(clus_len):            # :1485-1488 walks back over clusForMerge
                        if len(merge) - 1 - t < 0:
                            raise IndexError("the C# indexes clusForMerge[-1] here (ArgumentOutOfRangeException)")
                        cid[merge[len(merge) - 1 - t]] = 0
                else:
                    id_now += 1                          # :1492
                clus_len = 1
            else:
                clus_len += 1                            # :1498
            cid[k] = id_now                              # :1500
            merge.append(k)
            id_last = i_d
    merge = np.asarray(merge, np.int64)
    cf = cluster_sum - del_sum - 1                       # :1509
    is_zero = cid[merge] == 0
    zero_list = merge[is_zero]                           # :1510 FindAll keeps list order
    kept = merge[~is_zero]                               # :1511
    amount = cluster_sum - del_sum                       # :1508 (overwritten by dbscan below)
    if len(zero_list):
        pts = part.order[zero_list]
        res = dbscan(mx[pts], my[pts], eps, min_pts, cf) # :1516 (isClassed reset :1512-1515 is what the engine assumes)
        cid[zero_list] = res.cluster_id
        amount = res.cluster_amount
    else:
        amount = cf                                      # dbscan over an empty list sets clusterAmount = cf (DBImproved.cs:112)
    final_merge = np.concatenate([kept, zero_list])      # :1517-1520
    return cid, amount, final_merge, del_sum, cluster_sum


def centroids(points_xyz, mx, my, cid_of_point, order_idx, cluster_amount):
    """Tools.GetClusList, Tools.cs:162-195: per cluster the mean of X, Y, Z and of motor_x, motor_y over the members in
    list order (LINQ Average = sequential sum / count); clusters without members are skipped (:191)."""
    ids = cid_of_point[order_idx]
    centers, centers2d, center_ids = [], [], []
    for c in range(1, cluster_amount + 1):
        mem = order_idx[ids == c]
        if len(mem) == 0:
            continue
        seq_mean = lambda v: np.cumsum(v)[-1] / len(v)   # noqa: E731  cumsum adds left to right like the C# loop
        if points_xyz is not None:
            centers.append([seq_mean(points_xyz[mem, 0]), seq_mean(points_xyz[mem, 1]), seq_mean(points_xyz[mem, 2])])
        centers2d.append([seq_mean(mx[mem]), seq_mean(my[mem])])
        center_ids.append(c)
    return (np.asarray(centers) if points_xyz is not None else None), np.asarray(centers2d).reshape(-1, 2), np.asarray(center_ids, np.int64)


def cluster_blocked(mx, my, eps: float, min_pts: int, pts_in_cell: int, dbscan, dbscan_cells, points_xyz=None, argsort=None) -> BlockedResult:
    """The whole Clustering.DoClusteringBtn_Click path (Clustering.cs:78-98 -> FrmMain.cs:1214 -> 1340 -> 1432).
    dbscan(mx, my, eps, min_pts, first_cluster_id) -> object with .cluster_id, .cluster_amount;
    dbscan_cells(mx, my, offsets, eps, min_pts) -> (object with .cluster_id (cell-local), per_cell_amount)."""
    mx = np.ascontiguousarray(mx, np.float64)
    my = np.ascontiguousarray(my, np.float64)
    part = partition_cells(mx, my, pts_in_cell, argsort)
    res, per_cell = dbscan_cells(mx[part.order], my[part.order], part.offsets, eps, min_pts)   # every StartCode work item
    cid_sorted, amount, merge, del_sum, cluster_sum = complete_work3(part, res.cluster_id, per_cell, mx, my, eps, min_pts, dbscan)
    cluster_id = np.zeros(len(mx), np.int32)
    cluster_id[part.order] = cid_sorted
    order_idx = part.order[merge]
    c3, c2, cids = centroids(points_xyz, mx, my, cluster_id, order_idx, amount)
    return BlockedResult(cluster_id, amount, order_idx, del_sum, cluster_sum, part, c3, c2, cids)


def merge_ids_by_distance(centers_xy: np.ndarray, center_ids: np.ndarray, thre: float, dbscan) -> dict:
    """Tools.MergeIDByDistance, Tools.cs:580-621: DBSCAN(thre, minPts = 2) over the centroids' (X, Y); every later member
    of a centroid cluster maps to the first member's id."""
    res = dbscan(np.ascontiguousarray(centers_xy[:, 0]), np.ascontiguousarray(centers_xy[:, 1]), thre, 2, 0)   # :591-592
    cid = res.cluster_id
    dick, seen = {}, set()
    for i in range(len(center_ids)):                     # :594
        if cid[i] != 0:
            if int(center_ids[i]) not in seen:
                seen.add(int(center_ids[i]))
                for j in range(len(center_ids)):         # :602
                    if cid[j] == cid[i] and center_ids[j] != center_ids[i]:
                        seen.add(int(center_ids[j]))
                        dick[int(center_ids[j])] = int(center_ids[i])     # Dictionary.Add (:607)
        else:
            seen.add(int(center_ids[i]))                 # :614
    return dick


def refresh_by_dictionary(cluster_id: np.ndarray, cluster_amount: int, dick: dict):
    """Tools.refreshCensAndClusByDictionary, Tools.cs:521-572 (id part): merged clusters are appended to their target,
    the merged ids disappear, the survivors are renumbered 1.. in id order.  Returns (new cluster_id, new amount)."""
    out = np.asarray(cluster_id).copy()
    for src, dst in dick.items():                        # :525-533
        out[cluster_id == src] = dst
    survivors = [c for c in range(1, cluster_amount + 1) if c not in dick]      # :534-552
    remap = np.zeros(cluster_amount + 1, np.int64)
    for new, c in enumerate(survivors, start=1):         # :553-562
        remap[c] = new
    return remap[out].astype(np.int32), len(survivors)
