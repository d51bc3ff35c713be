"""ctypes binding of the CPU oracle (oracle/libvpc_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from functools import lru_cache
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "libvpc_oracle.so"


def build(force: bool = False) -> Path:
    srcs = [HERE / "vpc_oracle.cpp", HERE / "vpc_oracle.h"]
    if force or not LIB.exists() or any(s.stat().st_mtime > LIB.stat().st_mtime for s in srcs):
        res = subprocess.run(["make", "-C", str(HERE), "-B", "libvpc_oracle.so"], capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("oracle build failed:\n" + res.stdout + res.stderr)
    return LIB


_p, _i64, _i32, _f64 = C.c_void_p, C.c_int64, C.c_int32, C.c_double


@lru_cache(maxsize=None)
def lib() -> C.CDLL:
    dll = C.CDLL(str(build()))
    sig = {
        "vpco_dbscan_l1_2d_literal": [_p, _p, _i64, _f64, _i32, _i32, _p, _p, _p, _p, C.c_int, _p],
        "vpco_dbscan_l1_2d_grid": [_p, _p, _i64, _f64, _i32, _i32, _p, _p, _p, _p, C.c_int],
        "vpco_closest_point_set_literal": [_p, _i64, _p, _i64, _p, _p, C.c_int],
        "vpco_closest_point_set_grid": [_p, _i64, _p, _i64, _p, _p, C.c_int],
        "vpco_rigid_step": [_p, _p, _i64, _p, _p, _p],
        "vpco_icp_rigid": [_p, _i64, _p, _i64, _f64, _i32, _p, _p, _p, _p, _p, C.c_int, C.c_int],
        "vpco_jacobi_eig": [_p, C.c_int, _p, _p, C.c_int, _f64],
        "vpco_trans_points": [_p, _i64, _p, _p, _p],
        "vpco_match_within_literal": [_p, _i64, _p, _i64, _f64, _p, _p],
    }
    for name, args in sig.items():
        fn = getattr(dll, name)
        fn.restype = C.c_int
        fn.argtypes = args
    return dll


def _ptr(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _planar(xyz) -> np.ndarray:
    a = np.asarray(xyz, dtype=np.float64)
    if a.shape[0] == 3 and (a.ndim == 2 and a.shape[1] != 3):
        return np.ascontiguousarray(a)
    return np.ascontiguousarray(a.T)


def dbscan(mx, my, eps, min_pts, first_cluster_id=0, variant="grid", dedup_loop=False, n_threads=1):
    """Returns (cluster_id, is_key, is_classed, cluster_amount[, dist_evals])."""
    mx = np.ascontiguousarray(mx, np.float64)
    my = np.ascontiguousarray(my, np.float64)
    n = mx.shape[0]
    cid = np.empty(n, np.int32)
    key = np.empty(n, np.uint8)
    cls = np.empty(n, np.uint8)
    amount = C.c_int32(0)
    ap = C.cast(C.byref(amount), C.c_void_p)
    if variant == "literal":
        evals = C.c_int64(0)
        rc = lib().vpco_dbscan_l1_2d_literal(_ptr(mx), _ptr(my), n, eps, min_pts, first_cluster_id, _ptr(cid), _ptr(key),
                                             _ptr(cls), ap, int(dedup_loop), C.cast(C.byref(evals), C.c_void_p))
    else:
        rc = lib().vpco_dbscan_l1_2d_grid(_ptr(mx), _ptr(my), n, eps, min_pts, first_cluster_id, _ptr(cid), _ptr(key),
                                          _ptr(cls), ap, n_threads)
    if rc != 0:
        raise RuntimeError(f"oracle dbscan rc={rc}")
    return cid, key, cls, int(amount.value)


def closest_point_set(model_xyz, data_xyz, variant="grid", n_threads=1):
    model, data = _planar(model_xyz), _planar(data_xyz)
    m, n = model.shape[1], data.shape[1]
    order = np.empty(n, np.int32)
    sq = np.empty(n, np.float64)
    fn = lib().vpco_closest_point_set_literal if variant == "literal" else lib().vpco_closest_point_set_grid
    rc = fn(_ptr(model), m, _ptr(data), n, _ptr(order), _ptr(sq), n_threads)
    if rc != 0:
        raise RuntimeError(f"oracle closest_point_set rc={rc}")
    return order, sq


def rigid_step(P_xyz, Y_xyz):
    P, Y = _planar(P_xyz), _planar(Y_xyz)
    R1 = np.empty(9, np.float64)
    T1 = np.empty(3, np.float64)
    sse = C.c_double(0)
    rc = lib().vpco_rigid_step(_ptr(P), _ptr(Y), P.shape[1], _ptr(R1), _ptr(T1), C.cast(C.byref(sse), C.c_void_p))
    if rc != 0:
        raise RuntimeError(f"oracle rigid_step rc={rc}")
    return R1.reshape(3, 3), T1, float(sse.value)


def icp_rigid(model_xyz, data_xyz, e, max_iters=0, R0=None, T0=None, use_grid=True, n_threads=1):
    model, data = _planar(model_xyz), _planar(data_xyz)
    m, n = model.shape[1], data.shape[1]
    R = np.zeros(9) if R0 is None else np.ascontiguousarray(R0, np.float64).reshape(9).copy()
    T = np.zeros(3) if T0 is None else np.ascontiguousarray(T0, np.float64).reshape(3).copy()
    iters = C.c_int32(0)
    sse = C.c_double(0)
    order = np.empty(n, np.int32)
    rc = lib().vpco_icp_rigid(_ptr(model), m, _ptr(data), n, e, max_iters, _ptr(R), _ptr(T), C.cast(C.byref(iters), C.c_void_p),
                              C.cast(C.byref(sse), C.c_void_p), _ptr(order), int(use_grid), n_threads)
    if rc != 0:
        raise RuntimeError(f"oracle icp_rigid rc={rc}")
    return R.reshape(3, 3), T, int(iters.value), float(sse.value), order


def jacobi_eig(a, max_it=100, eps=1e-4):
    a = np.ascontiguousarray(a, np.float64).copy()
    n = a.shape[0]
    w = np.empty(n)
    v = np.zeros((n, n))
    ok = lib().vpco_jacobi_eig(_ptr(a), n, _ptr(w), _ptr(v), max_it, eps)
    return bool(ok), w, v, a


def trans_points(src_xyz, R, T):
    src = _planar(src_xyz)
    dst = np.empty_like(src)
    R = np.ascontiguousarray(R, np.float64).reshape(9)
    T = np.ascontiguousarray(T, np.float64).reshape(3)
    lib().vpco_trans_points(_ptr(src), src.shape[1], _ptr(R), _ptr(T), _ptr(dst))
    return dst


def match_within(truth_xyz, centers_xyz, match_distance):
    truth, cen = _planar(truth_xyz), _planar(centers_xyz)
    n = cen.shape[1]
    mid = np.empty(n, np.int32)
    dist = np.empty(n, np.float64)
    rc = lib().vpco_match_within_literal(_ptr(truth), truth.shape[1], _ptr(cen), n, match_distance, _ptr(mid), _ptr(dist))
    if rc != 0:
        raise RuntimeError(f"oracle match_within rc={rc}")
    return mid, dist
