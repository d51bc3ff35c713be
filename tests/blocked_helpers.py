"""Oracle-backed engines for vtkcloudpoint_b200.blocked (test infrastructure)."""
from types import SimpleNamespace

import numpy as np

import oracle_py


def oracle_dbscan(mx, my, eps, min_pts, first_cluster_id):
    cid, key, cls, amount = oracle_py.dbscan(mx, my, eps, min_pts, first_cluster_id, variant="literal" if len(mx) <= 4000 else "grid")
    return SimpleNamespace(cluster_id=cid, is_key=key, is_classed=cls, cluster_amount=amount)


def oracle_dbscan_cells(mx, my, offsets, eps, min_pts):
    cid = np.zeros(len(mx), np.int32)
    per_cell = np.zeros(len(offsets) - 1, np.int32)
    for k in range(len(offsets) - 1):
        a, b = int(offsets[k]), int(offsets[k + 1])
        c, _, _, amount = oracle_py.dbscan(mx[a:b], my[a:b], eps, min_pts, 0, variant="literal")   # one StartCode work item
        cid[a:b] = c
        per_cell[k] = amount
    return SimpleNamespace(cluster_id=cid), per_cell
