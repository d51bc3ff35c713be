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
    """One vpc_ctx = one GPU.  Raises VpcError when no CUDA device is usable."""

    def __init__(self, device: int = 0):
        self._lib = capi.lib()
        self._h = C.c_void_p()
        ids = (C.c_int * 1)(device)
        rc = self._lib.vpc_create(C.byref(self._h), ids, 1)
        if rc != capi.VPC_OK:
            self._h = C.c_void_p()
            raise capi.VpcError(rc, "vpc_create failed (no CPU fallback exists)")
        self.device = device

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
