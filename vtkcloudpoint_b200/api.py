"""Array-level Python host API over the C ABI (include/vpc.h).

Two flavours per operation:
  * host (NumPy) arrays  -> the host-pointer exports, H2D/D2H inside the call
  * device (torch.cuda) tensors -> the *_dev exports, enqueued on torch's current stream

PyTorch is used for device memory and streams only; all arithmetic is in libvpc.so.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import capi


def _ptr(a: np.ndarray) -> C.c_void_p:
    return C.c_void_p(a.ctypes.data)


def _planar(xyz) -> np.ndarray:
    """(k,3) or planar (3,k) float64 -> C-contiguous planar (3,k)."""
    a = np.asarray(xyz, dtype=np.float64)
    if a.ndim != 2:
        raise ValueError("point set must be 2-D: (k,3) or planar (3,k)")
    if a.shape[0] == 3 and a.shape[1] != 3:
        return np.ascontiguousarray(a)
    if a.shape[1] == 3:
        return np.ascontiguousarray(a.T)
    raise ValueError(f"bad point-set shape {a.shape}")


@dataclass
class DbscanResult:
    cluster_id: np.ndarray   # int32, 0 = noise   (Point3D.clusterId)
    is_key: np.ndarray       # uint8              (Point3D.isKeyPoint)
    is_classed: np.ndarray   # uint8              (Point3D.isClassed)
    cluster_amount: int      # DBImproved.clusterAmount


@dataclass
class IcpResult:
    R: np.ndarray            # (3,3) row-major, Matrix(3,3).mat
    T: np.ndarray            # (3,)
    iters_done: int
    sse_last: float
    order_last: np.ndarray | None
    converged: bool | None = None

    @property
    def rmse(self) -> float:
        n = len(self.order_last) if self.order_last is not None else 1
        return float(np.sqrt(self.sse_last / n))


class Context:
    """One vpc_ctx.  device = an int: one GPU.  device = a list of ints: ONE context driving several GPUs from this process
    (vpc_create with n_devices > 1): the host-array calls dbscan() / icp_rigid() are then spread over those devices inside the
    library (csrc/group_api.cuh); everything else runs on the first one.  Raises VpcError when no CUDA device is usable."""

    def __init__(self, device=0):
        self._lib = capi.lib()
        self._h = C.c_void_p()
        devs = [int(d) for d in device] if isinstance(device, (list, tuple)) else [int(device)]
        ids = (C.c_int * len(devs))(*devs)
        rc = self._lib.vpc_create(C.byref(self._h), ids, len(devs))
        if rc != capi.VPC_OK:
            self._h = C.c_void_p()
            raise capi.VpcError(rc, "vpc_create failed (no CPU fallback exists)")
        self.device = devs[0]
        self.devices = devs

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.vpc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def _check(self, rc: int):
        if rc != capi.VPC_OK:
            raise capi.VpcError(rc, self._lib.vpc_last_error(self._h).decode())

    @property
    def launch_count(self) -> int:
        return int(self._lib.vpc_launch_count(self._h))

    def profile(self, on: bool):
        self._check(self._lib.vpc_profile_enable(self._h, 1 if on else 0))

    def profile_report(self) -> list[tuple[str, float]]:
        """[(kernel_name, ms)] for every launch since the last report (synchronises the device)."""
        buf = C.create_string_buffer(1 << 22)
        nb = self._lib.vpc_profile_report(self._h, buf, len(buf))
        if nb < 0:
            self._check(int(nb))
        out = []
        for line in buf.value.decode().splitlines():
            name, ms = line.rsplit(" ", 1)
            out.append((name, float(ms)))
        return out

    # ------------------------------------------------------------------ DBSCAN, host arrays
    def dbscan(self, mx, my, eps: float, min_pts: int, first_cluster_id: int = 0, out: DbscanResult | None = None) -> DbscanResult:
        mx = np.ascontiguousarray(mx, dtype=np.float64)
        my = np.ascontiguousarray(my, dtype=np.float64)
        if mx.shape != my.shape or mx.ndim != 1:
            raise ValueError("mx and my must be 1-D arrays of the same length")
        n = mx.shape[0]
        if out is None:
            out = DbscanResult(np.empty(n, np.int32), np.empty(n, np.uint8), np.empty(n, np.uint8), 0)
        amount = C.c_int32(0)
        self._check(self._lib.vpc_dbscan_l1_2d(self._h, _ptr(mx), _ptr(my), n, float(eps), int(min_pts), int(first_cluster_id),
                                               _ptr(out.cluster_id), _ptr(out.is_key), _ptr(out.is_classed),
                                               C.cast(C.byref(amount), C.c_void_p)))
        out.cluster_amount = int(amount.value)
        return out

    def dbscan_cells(self, mx, my, cell_offsets, eps: float, min_pts: int):
        """Batched per-cell DBSCAN (all StartCode work items of the blocked clustering in one launch).
        Returns (DbscanResult with cell-local ids and cluster_amount = sum, cluster_amount_per_cell int32[n_cells])."""
        mx = np.ascontiguousarray(mx, dtype=np.float64)
        my = np.ascontiguousarray(my, dtype=np.float64)
        off = np.ascontiguousarray(cell_offsets, dtype=np.int64)
        n, n_cells = mx.shape[0], off.shape[0] - 1
        out = DbscanResult(np.empty(n, np.int32), np.empty(n, np.uint8), np.empty(n, np.uint8), 0)
        per_cell = np.zeros(max(n_cells, 0), np.int32)
        self._check(self._lib.vpc_dbscan_l1_2d_cells(self._h, _ptr(mx), _ptr(my), n, _ptr(off), n_cells, float(eps), int(min_pts),
                                                     _ptr(out.cluster_id), _ptr(out.is_key), _ptr(out.is_classed), _ptr(per_cell)))
        out.cluster_amount = int(per_cell.sum())
        return out, per_cell

    # ------------------------------------------------------------------ DBSCAN, device tensors
    def dbscan_dev(self, mx, my, eps: float, min_pts: int, first_cluster_id: int = 0, out=None):
        """mx, my: float64 CUDA tensors.  Returns (cluster_id i32, is_key u8, is_classed u8, amount i32[1]) tensors.
        Enqueued on torch's current stream; nothing is synchronised."""
        import torch
        assert mx.is_cuda and my.is_cuda and mx.dtype == torch.float64 and my.dtype == torch.float64
        assert mx.is_contiguous() and my.is_contiguous() and mx.numel() == my.numel()
        n = mx.numel()
        if out is None:
            out = (torch.empty(n, dtype=torch.int32, device=mx.device), torch.empty(n, dtype=torch.uint8, device=mx.device),
                   torch.empty(n, dtype=torch.uint8, device=mx.device), torch.empty(1, dtype=torch.int32, device=mx.device))
        cid, key, cls, amount = out
        stream = torch.cuda.current_stream(mx.device).cuda_stream
        self._check(self._lib.vpc_dbscan_l1_2d_dev(self._h, mx.data_ptr(), my.data_ptr(), n, float(eps), int(min_pts),
                                                   int(first_cluster_id), cid.data_ptr(), key.data_ptr(), cls.data_ptr(),
                                                   amount.data_ptr(), stream))
        return out

    # ------------------------------------------------------------------ ICP, host arrays
    def closest_point_set(self, model_xyz, data_xyz, want_sqdist: bool = True):
        model = _planar(model_xyz)
        data = _planar(data_xyz)
        m, n = model.shape[1], data.shape[1]
        order = np.empty(n, np.int32)
        sq = np.empty(n, np.float64) if want_sqdist else None
        self._check(self._lib.vpc_closest_point_set(self._h, _ptr(model), m, _ptr(data), n, _ptr(order),
                                                    _ptr(sq) if sq is not None else None))
        return order, sq

    def icp_rigid(self, model_xyz, data_xyz, e: float, max_iters: int = 0, R0=None, T0=None, want_order: bool = True) -> IcpResult:
        model = _planar(model_xyz)
        data = _planar(data_xyz)
        m, n = model.shape[1], data.shape[1]
        R = np.zeros(9, np.float64) if R0 is None else np.ascontiguousarray(R0, np.float64).reshape(9).copy()
        T = np.zeros(3, np.float64) if T0 is None else np.ascontiguousarray(T0, np.float64).reshape(3).copy()
        iters = C.c_int32(0)
        sse = C.c_double(0.0)
        order = np.empty(n, np.int32) if want_order else None
        self._check(self._lib.vpc_icp_rigid(self._h, _ptr(model), m, _ptr(data), n, float(e), int(max_iters), _ptr(R), _ptr(T),
                                            C.cast(C.byref(iters), C.c_void_p), C.cast(C.byref(sse), C.c_void_p),
                                            _ptr(order) if order is not None else None))
        return IcpResult(R.reshape(3, 3), T, int(iters.value), float(sse.value), order)

    def match_within(self, truth_xyz, centers_xyz, match_distance: float):
        """MainForm.RecorrectMatchingPtsByDistance's search (FrmMain.cs:3588-3618): (matched_id int32 [-1 = unmatched], dist)."""
        truth, cen = _planar(truth_xyz), _planar(centers_xyz)
        n = cen.shape[1]
        mid = np.empty(n, np.int32)
        dist = np.empty(n, np.float64)
        self._check(self._lib.vpc_match_within(self._h, _ptr(truth), truth.shape[1], _ptr(cen), n, float(match_distance), _ptr(mid), _ptr(dist)))
        return mid, dist

    def cluster_means_dev(self, cluster_id, n_clusters: int, vals_planar):
        """Tools.GetClusList's averages (Tools.cs:187-194) on the device.  cluster_id: int32 CUDA tensor [n]; vals_planar:
        float64 CUDA tensor [n_fields, n].  Returns (means [n_fields, n_clusters + 1], counts int32 [n_clusters + 1])."""
        import torch
        assert cluster_id.is_cuda and cluster_id.dtype == torch.int32 and vals_planar.is_cuda and vals_planar.dtype == torch.float64
        assert vals_planar.is_contiguous() and cluster_id.is_contiguous() and vals_planar.shape[1] == cluster_id.numel()
        nf = vals_planar.shape[0]
        means = torch.empty((nf, n_clusters + 1), dtype=torch.float64, device=cluster_id.device)
        counts = torch.empty(n_clusters + 1, dtype=torch.int32, device=cluster_id.device)
        stream = torch.cuda.current_stream(cluster_id.device).cuda_stream
        self._check(self._lib.vpc_cluster_means_dev(self._h, cluster_id.data_ptr(), cluster_id.numel(), int(n_clusters), vals_planar.data_ptr(),
                                                    nf, means.data_ptr(), counts.data_ptr(), stream))
        return means, counts

    # ------------------------------------------------------------------ ICP, device tensors
    def icp_set_model_dev(self, model_planar):
        """model_planar: float64 CUDA tensor of shape (3, m), contiguous.  It must stay alive while queries run."""
        import torch
        assert model_planar.is_cuda and model_planar.dtype == torch.float64 and model_planar.is_contiguous()
        assert model_planar.dim() == 2 and model_planar.shape[0] == 3
        self._model_keepalive = model_planar
        stream = torch.cuda.current_stream(model_planar.device).cuda_stream
        self._check(self._lib.vpc_icp_set_model_dev(self._h, model_planar.data_ptr(), model_planar.shape[1], stream))

    def closest_point_set_dev(self, data_planar, want_sqdist: bool = True):
        import torch
        assert data_planar.is_cuda and data_planar.dtype == torch.float64 and data_planar.is_contiguous()
        n = data_planar.shape[1]
        order = torch.empty(n, dtype=torch.int32, device=data_planar.device)
        sq = torch.empty(n, dtype=torch.float64, device=data_planar.device) if want_sqdist else None
        stream = torch.cuda.current_stream(data_planar.device).cuda_stream
        self._check(self._lib.vpc_closest_point_set_dev(self._h, data_planar.data_ptr(), n, order.data_ptr(),
                                                        sq.data_ptr() if sq is not None else None, stream))
        return order, sq

    def icp_rigid_dev(self, data_planar, e: float, max_iters: int, out=None):
        """Returns (state f64[16] = R[9] T[3] sse iters converged 0, order_last i32[n]) device tensors; no sync."""
        import torch
        assert data_planar.is_cuda and data_planar.dtype == torch.float64 and data_planar.is_contiguous()
        n = data_planar.shape[1]
        if out is None:
            out = (torch.empty(16, dtype=torch.float64, device=data_planar.device),
                   torch.empty(n, dtype=torch.int32, device=data_planar.device))
        state, order = out
        stream = torch.cuda.current_stream(data_planar.device).cuda_stream
        self._check(self._lib.vpc_icp_rigid_dev(self._h, data_planar.data_ptr(), n, float(e), int(max_iters), state.data_ptr(),
                                                order.data_ptr(), stream))
        return out

    # ------------------------------------------------------------------ cluster statistics / matching / ingest
    def cluster_stats(self, cluster_id, n_clusters: int, xyz, mx, my, circles3d: bool = True, circles2d: bool = True):
        """The statistics block of CompleteWork3 (FrmMain.cs:1521-1540) through the host-pointer ABI.  Returns a dict:
        means [5, k+1] (X Y Z motor_x motor_y), counts [k+1], circle3d / circle2d [3, k+1] (cx, cy, radius) and their status [k+1]."""
        cid = np.ascontiguousarray(cluster_id, np.int32)
        pts = _planar(xyz) if len(cid) else np.zeros((3, 0))
        mx = np.ascontiguousarray(mx, np.float64)
        my = np.ascontiguousarray(my, np.float64)
        n, k1 = len(cid), int(n_clusters) + 1
        means = np.empty((5, k1), np.float64)
        counts = np.empty(k1, np.int32)
        c3 = np.empty((3, k1), np.float64) if circles3d else None
        s3 = np.empty(k1, np.int32) if circles3d else None
        c2 = np.empty((3, k1), np.float64) if circles2d else None
        s2 = np.empty(k1, np.int32) if circles2d else None
        opt = lambda a: _ptr(a) if a is not None else None   # noqa: E731
        self._check(self._lib.vpc_cluster_stats(self._h, _ptr(cid), n, int(n_clusters), _ptr(pts), _ptr(mx), _ptr(my), _ptr(means), _ptr(counts),
                                                opt(c3), opt(s3), opt(c2), opt(s2)))
        return {"means": means, "counts": counts, "circle3d": c3, "status3d": s3, "circle2d": c2, "status2d": s2}

    def nearest_truth_2d(self, truth_x, truth_y, truth_id, px, py, radius: float):
        """MainForm.refreshClusList's query (FrmMain.cs:3446-3467): id of the nearest truth within radius (ties -> highest index), else 0."""
        tx = np.ascontiguousarray(truth_x, np.float64)
        ty = np.ascontiguousarray(truth_y, np.float64)
        tid = None if truth_id is None else np.ascontiguousarray(truth_id, np.int32)
        px = np.ascontiguousarray(px, np.float64)
        py = np.ascontiguousarray(py, np.float64)
        out = np.empty(len(px), np.int32)
        self._check(self._lib.vpc_nearest_truth_2d(self._h, _ptr(tx), _ptr(ty), _ptr(tid) if tid is not None else None, len(tx), _ptr(px), _ptr(py),
                                                   len(px), float(radius), _ptr(out)))
        return out

    def ingest_text(self, text: bytes, x_angle: float, y_angle: float, xdir: int = 2, ydir: int = 1, remove_duplicates: bool = True):
        """A scan file in memory -> dict(mx, my, dist, xyz [3, rows], keep, row_status, n_kept, n_duplicates) (FrmMain.cs:975-1068)."""
        cap = text.count(b"\n") + 1
        mx, my, ds = (np.empty(cap, np.float64) for _ in range(3))
        xyz = np.empty(3 * cap, np.float64)
        keep, st = np.empty(cap, np.uint8), np.empty(cap, np.uint8)
        n_rows, n_kept, n_dup = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        ref = lambda v: C.cast(C.byref(v), C.c_void_p)   # noqa: E731
        self._check(self._lib.vpc_ingest_text(self._h, text, len(text), float(x_angle), float(y_angle), int(xdir), int(ydir), 1 if remove_duplicates else 0,
                                              cap, _ptr(mx), _ptr(my), _ptr(ds), _ptr(xyz), _ptr(keep), _ptr(st), ref(n_rows), ref(n_kept), ref(n_dup)))
        r = int(n_rows.value)
        return {"mx": mx[:r], "my": my[:r], "dist": ds[:r], "xyz": xyz[:3 * r].reshape(3, r), "keep": keep[:r], "row_status": st[:r],
                "n_kept": int(n_kept.value), "n_duplicates": int(n_dup.value)}

    def dbscan_blocked_ref(self, mx, my, eps: float, min_pts: int, pts_in_cell: int):
        """The reference's blocked clustering (getClusterFromMotor -> DoWork3 -> CompleteWork3) as one call, on the device.
        Returns dict(cluster_id, cluster_sum, del_sum, rows, cols, n_unassigned, n_shared, merge_order, merge_cid, cluster_sum_cells)."""
        mx = np.ascontiguousarray(mx, np.float64)
        my = np.ascontiguousarray(my, np.float64)
        n = len(mx)
        cid = np.zeros(n, np.int32)
        mo, mc = np.zeros(3 * n + 1, np.int64), np.zeros(3 * n + 1, np.int32)
        cs, ds, r, c, csc = (C.c_int32(0) for _ in range(5))
        un, sh, nm = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        ref = lambda v: C.cast(C.byref(v), C.c_void_p)   # noqa: E731
        self._check(self._lib.vpc_dbscan_blocked_ref_ex(self._h, _ptr(mx), _ptr(my), n, float(eps), int(min_pts), int(pts_in_cell), _ptr(cid),
                                                        ref(cs), ref(ds), ref(r), ref(c), ref(un), ref(sh), _ptr(mo), _ptr(mc), ref(nm), ref(csc)))
        k = int(nm.value)
        return {"cluster_id": cid, "cluster_sum": int(cs.value), "del_sum": int(ds.value), "rows": int(r.value), "cols": int(c.value),
                "n_unassigned": int(un.value), "n_shared": int(sh.value), "merge_order": mo[:k].copy(), "merge_cid": mc[:k].copy(),
                "cluster_sum_cells": int(csc.value)}

    def merge_ids_by_distance(self, merge_cid, xyz_entries, mx_entries, my_entries, cluster_amount: int, thre: float):
        """Clustering.MergeBtn_Click's chain (GetClusList -> MergeIDByDistance -> refreshCensAndClusByDictionary) on the clusForMerge
        list; arrays per ENTRY in list order.  Returns dict(cluster_id, cluster_amount, dict [(from, to)], centers5, center_ids, new_centers5)."""
        mcid = np.ascontiguousarray(merge_cid, np.int32)
        k = len(mcid)
        xyz = _planar(xyz_entries)
        mx = np.ascontiguousarray(mx_entries, np.float64); my = np.ascontiguousarray(my_entries, np.float64)
        cap = max(int(cluster_amount), 1)
        new_cid = np.zeros(k, np.int32)
        dfrom, dto, cids = np.zeros(cap, np.int32), np.zeros(cap, np.int32), np.zeros(cap, np.int32)
        c5, nc5 = np.zeros(5 * cap), np.zeros(5 * cap)
        amount, nd, ncen = C.c_int32(0), C.c_int32(0), C.c_int32(0)
        ref = lambda v: C.cast(C.byref(v), C.c_void_p)   # noqa: E731
        self._check(self._lib.vpc_merge_ids_by_distance(self._h, _ptr(mcid), _ptr(xyz), _ptr(mx), _ptr(my), k, int(cluster_amount), float(thre), _ptr(new_cid),
                                                        ref(amount), _ptr(dfrom), _ptr(dto), ref(nd), _ptr(c5), _ptr(cids), ref(ncen), _ptr(nc5)))
        m, a = int(ncen.value), int(amount.value)
        return {"cluster_id": new_cid, "cluster_amount": a, "dict": list(zip(dfrom[:nd.value].tolist(), dto[:nd.value].tolist())),
                "centers5": c5[:5 * m].reshape(5, m).copy(), "center_ids": cids[:m].copy(), "new_centers5": nc5[:5 * a].reshape(5, a).copy()}

    def synth_dbscan_cloud_dev(self, seed: int, grid: int, n_total: int, start: int = 0, count: int | None = None, pts_per_cluster: int = 40, pitch: float = 0.5,
                               sigma: float = 0.012, x0: float = 149.0, y0: float = 307.0, device=None):
        """synth.dbscan_cloud generated on the device (bit for bit the same doubles).  Returns (mx, my) float64 CUDA tensors."""
        import torch
        count = n_total - start if count is None else count
        dev = torch.device("cuda", self.device) if device is None else device
        mx = torch.empty(count, dtype=torch.float64, device=dev); my = torch.empty(count, dtype=torch.float64, device=dev)
        self._check(self._lib.vpc_synth_dbscan_cloud_dev(self._h, int(seed), int(grid), int(pts_per_cluster), int(n_total), float(pitch), float(sigma), float(x0), float(y0),
                                                         int(start), int(count), mx.data_ptr(), my.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
        return mx, my

    def argsort_f64(self, vals) -> np.ndarray:
        """Stable ascending permutation of a host array of doubles on the device (vpc_argsort_f64_dev): the sort of
        MainForm.getClusterFromMotor (FrmMain.cs:1229-1233)."""
        import torch
        v = torch.from_numpy(np.ascontiguousarray(vals, np.float64)).to(f"cuda:{self.device}")
        return self.argsort_f64_dev(v).cpu().numpy()

    # device-tensor forms ---------------------------------------------------------------------------------------------
    def _stream(self, t):
        import torch
        return torch.cuda.current_stream(t.device).cuda_stream

    def sort_pairs_dev(self, keys, vals=None, begin_bit: int = 0, end_bit: int = 64):
        """Stable radix sort in place.  keys: int64/uint64-viewed CUDA tensor; vals None -> returns the sorting permutation."""
        import torch
        n = keys.numel()
        identity = vals is None
        if identity:
            vals = torch.empty(n, dtype=torch.int32, device=keys.device)
        self._check(self._lib.vpc_sort_pairs_dev(self._h, keys.data_ptr(), vals.data_ptr(), n, begin_bit, end_bit, 1 if identity else 0, self._stream(keys)))
        return keys, vals

    def argsort_f64_dev(self, vals):
        import torch
        order = torch.empty(vals.numel(), dtype=torch.int32, device=vals.device)
        self._check(self._lib.vpc_argsort_f64_dev(self._h, vals.data_ptr(), vals.numel(), order.data_ptr(), self._stream(vals)))
        return order

    def cluster_groups_dev(self, cluster_id, n_clusters: int):
        import torch
        n = cluster_id.numel()
        members = torch.empty(max(n, 1), dtype=torch.int32, device=cluster_id.device)
        offsets = torch.empty(n_clusters + 2, dtype=torch.int32, device=cluster_id.device)
        self._check(self._lib.vpc_cluster_groups_dev(self._h, cluster_id.data_ptr(), n, int(n_clusters), members.data_ptr(), offsets.data_ptr(), self._stream(cluster_id)))
        return members[:n], offsets

    def cluster_means_ordered_dev(self, members, offsets, n_clusters: int, vals_planar):
        import torch
        nf, n = vals_planar.shape
        means = torch.empty((nf, n_clusters + 1), dtype=torch.float64, device=vals_planar.device)
        counts = torch.empty(n_clusters + 1, dtype=torch.int32, device=vals_planar.device)
        self._check(self._lib.vpc_cluster_means_ordered_dev(self._h, members.data_ptr(), offsets.data_ptr(), int(n_clusters), vals_planar.data_ptr(), n, nf,
                                                            means.data_ptr(), counts.data_ptr(), self._stream(vals_planar)))
        return means, counts

    def cluster_circles_dev(self, members, offsets, n_clusters: int, hx, hy):
        import torch
        k1 = n_clusters + 1
        circ = torch.empty((3, k1), dtype=torch.float64, device=hx.device)
        status = torch.empty(k1, dtype=torch.int32, device=hx.device)
        self._check(self._lib.vpc_cluster_circles_dev(self._h, members.data_ptr(), offsets.data_ptr(), int(n_clusters), hx.numel(), hx.data_ptr(), hy.data_ptr(),
                                                      circ[0].data_ptr(), circ[1].data_ptr(), circ[2].data_ptr(), status.data_ptr(), self._stream(hx)))
        return circ, status

    def radius_filter_dev(self, radius, status, n_clusters: int, threshold: float):
        import torch
        flag = torch.empty(n_clusters + 1, dtype=torch.uint8, device=radius.device)
        self._check(self._lib.vpc_radius_filter_dev(self._h, radius.data_ptr(), status.data_ptr(), int(n_clusters), float(threshold), flag.data_ptr(), self._stream(radius)))
        return flag

    def nearest_truth_2d_dev(self, truth_id, px, py, radius: float, want_extras: bool = False):
        import torch
        n = px.numel()
        out = torch.empty(n, dtype=torch.int32, device=px.device)
        idx = torch.empty(n, dtype=torch.int32, device=px.device) if want_extras else None
        dist = torch.empty(n, dtype=torch.float64, device=px.device) if want_extras else None
        self._check(self._lib.vpc_nearest_truth_2d_dev(self._h, truth_id.data_ptr() if truth_id is not None else None, px.data_ptr(), py.data_ptr(), n, float(radius),
                                                       out.data_ptr(), idx.data_ptr() if want_extras else None, dist.data_ptr() if want_extras else None, self._stream(px)))
        return (out, idx, dist) if want_extras else out

    def polar_to_xyz_dev(self, mx, my, dist, x_angle: float, y_angle: float, xdir: int = 2, ydir: int = 1):
        import torch
        n = mx.numel()
        xyz = torch.empty((3, n), dtype=torch.float64, device=mx.device)
        keep = torch.empty(n, dtype=torch.uint8, device=mx.device)
        self._check(self._lib.vpc_polar_to_xyz_dev(self._h, mx.data_ptr(), my.data_ptr(), dist.data_ptr(), n, float(x_angle), float(y_angle), int(xdir), int(ydir),
                                                   xyz.data_ptr(), keep.data_ptr(), self._stream(mx)))
        return xyz, keep

    def dedupe_xyz_dev(self, xyz_planar, live=None):
        import torch
        n = xyz_planar.shape[1]
        keep = torch.empty(n, dtype=torch.uint8, device=xyz_planar.device)
        first = torch.empty(n, dtype=torch.int32, device=xyz_planar.device)
        ndup = torch.zeros(1, dtype=torch.int32, device=xyz_planar.device)
        self._check(self._lib.vpc_dedupe_xyz_dev(self._h, xyz_planar.data_ptr(), live.data_ptr() if live is not None else None, n, keep.data_ptr(), first.data_ptr(),
                                                 ndup.data_ptr(), self._stream(xyz_planar)))
        return keep, first, ndup
