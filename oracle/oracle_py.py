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
    srcs = [HERE / "vpc_oracle.cpp", HERE / "vpc_oracle_stats.cpp", HERE / "vpc_oracle_aswritten.cpp", HERE / "vpc_oracle_blocked.cpp", HERE / "vpc_oracle.h"]
    import hashlib
    h = hashlib.sha256()
    for src in srcs + [HERE / "Makefile"]:
        h.update(src.read_bytes())
    stamp = HERE / "libvpc_oracle.so.stamp"           # contents, not file times: those do not survive a snapshot copy
    if force or not LIB.exists() or not stamp.exists() or stamp.read_text().strip() != h.hexdigest():
        res = subprocess.run(["make", "-C", str(HERE), "-B", "libvpc_oracle.so"], capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("oracle build failed:\n" + res.stdout + res.stderr)
        stamp.write_text(h.hexdigest())
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
        "vpco_jacobi_eig_as_written": [_p, C.c_int, _p, _p, C.c_int, _f64],
        "vpco_icp_as_written": [_p, _i64, _p, _i64, _f64, _i32, _p, _p, _p, _p, _p, _i32],
        "vpco_cluster_stats_literal": [_p, _i64, _i32, _p, _p, _p, _p, _p, _p, _p, _p, _p],
        "vpco_nearest_truth_2d_literal": [_p, _p, _p, _i64, _p, _p, _i64, _f64, _p],
        "vpco_polar_to_xyz": [_p, _p, _p, _i64, _f64, _f64, _i32, _i32, _p, _p],
        "vpco_dedupe_xyz_literal": [_p, _p, _i64, _p, _p, _p],
        "vpco_parse_rows": [C.c_char_p, _i64, _i64, _p, _p, _p, _p, _p],
        "vpco_blocked_literal": [_p, _p, _i64, _f64, _i32, _i32, C.c_int, C.c_int, C.c_int, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p],
        "vpco_merge_ids_literal": [_p, _p, _p, _p, _i64, _i32, _f64, _p, _p, _p, _p, _p, _p, _p, _p, _p],
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


def cluster_stats(cluster_id, n_clusters, xyz, mx, my, circles3d=True, circles2d=True):
    cid = np.ascontiguousarray(cluster_id, np.int32)
    pts = _planar(xyz) if len(cid) else np.zeros((3, 0))
    mx = np.ascontiguousarray(mx, np.float64)
    my = np.ascontiguousarray(my, np.float64)
    k1 = int(n_clusters) + 1
    means = np.empty((5, k1)); counts = np.empty(k1, np.int32)
    c3 = np.empty((3, k1)) if circles3d else None
    s3 = np.empty(k1, np.int32) if circles3d else None
    c2 = np.empty((3, k1)) if circles2d else None
    s2 = np.empty(k1, np.int32) if circles2d else None
    rc = lib().vpco_cluster_stats_literal(_ptr(cid), len(cid), int(n_clusters), _ptr(pts), _ptr(mx), _ptr(my), _ptr(means), _ptr(counts),
                                          _ptr(c3), _ptr(s3), _ptr(c2), _ptr(s2))
    if rc != 0:
        raise RuntimeError(f"oracle cluster_stats rc={rc}")
    return {"means": means, "counts": counts, "circle3d": c3, "status3d": s3, "circle2d": c2, "status2d": s2}


def nearest_truth_2d(truth_x, truth_y, truth_id, px, py, radius):
    tx = np.ascontiguousarray(truth_x, np.float64); ty = np.ascontiguousarray(truth_y, np.float64)
    tid = None if truth_id is None else np.ascontiguousarray(truth_id, np.int32)
    px = np.ascontiguousarray(px, np.float64); py = np.ascontiguousarray(py, np.float64)
    out = np.empty(len(px), np.int32)
    rc = lib().vpco_nearest_truth_2d_literal(_ptr(tx), _ptr(ty), _ptr(tid), len(tx), _ptr(px), _ptr(py), len(px), float(radius), _ptr(out))
    if rc != 0:
        raise RuntimeError(f"oracle nearest_truth_2d rc={rc}")
    return out


def polar_to_xyz(mx, my, dist, x_angle, y_angle, xdir=2, ydir=1):
    mx = np.ascontiguousarray(mx, np.float64); my = np.ascontiguousarray(my, np.float64); dist = np.ascontiguousarray(dist, np.float64)
    n = len(mx)
    xyz = np.empty((3, n)); keep = np.empty(n, np.uint8)
    rc = lib().vpco_polar_to_xyz(_ptr(mx), _ptr(my), _ptr(dist), n, float(x_angle), float(y_angle), int(xdir), int(ydir), _ptr(xyz), _ptr(keep))
    if rc != 0:
        raise RuntimeError(f"oracle polar_to_xyz rc={rc}")
    return xyz, keep


def dedupe_xyz(xyz_planar, live=None):
    xyz = np.ascontiguousarray(xyz_planar, np.float64)
    n = xyz.shape[1]
    live = None if live is None else np.ascontiguousarray(live, np.uint8)
    keep = np.empty(n, np.uint8); first = np.empty(n, np.int32); ndup = C.c_int64(0)
    rc = lib().vpco_dedupe_xyz_literal(_ptr(xyz), _ptr(live), n, _ptr(keep), _ptr(first), C.cast(C.byref(ndup), C.c_void_p))
    if rc != 0:
        raise RuntimeError(f"oracle dedupe rc={rc}")
    return keep, first, int(ndup.value)


def parse_rows(text: bytes):
    cap = text.count(b"\n") + 1
    mx, my, ds = (np.empty(cap) for _ in range(3))
    st = np.empty(cap, np.uint8); rows = C.c_int64(0)
    rc = lib().vpco_parse_rows(text, len(text), cap, _ptr(mx), _ptr(my), _ptr(ds), _ptr(st), C.cast(C.byref(rows), C.c_void_p))
    if rc != 0:
        raise RuntimeError(f"oracle parse_rows rc={rc}")
    r = int(rows.value)
    return mx[:r], my[:r], ds[:r], st[:r]


def jacobi_eig_as_written(a, max_it=100, eps=1e-4):
    """Matrix.ComputeEvJacobi exactly as written (Matrix.cs:571-668).  Returns (ok, eigenvalues, V, a_after)."""
    a = np.ascontiguousarray(a, np.float64).copy()
    n = a.shape[0]
    w = np.zeros(n); v = np.zeros((n, n))
    ok = lib().vpco_jacobi_eig_as_written(_ptr(a), n, _ptr(w), _ptr(v), max_it, eps)
    return bool(ok), w, v, a


def icp_as_written(model_xyz, data_xyz, e, max_rounds=0, max_trace=64):
    """ICP.go_hell_ICP exactly as written.  Returns dict(rc, R, T, rounds, sse[rounds], jacobi_ok[rounds]); rc == -7 means the C#
    throws IndexOutOfRangeException (ICP.cs:170-174)."""
    model, data = _planar(model_xyz), _planar(data_xyz)
    R = np.zeros(9); T = np.zeros(3)
    rounds = C.c_int32(0)
    sse = np.zeros(max_trace); jok = np.zeros(max_trace, np.int32)
    rc = lib().vpco_icp_as_written(_ptr(model), model.shape[1], _ptr(data), data.shape[1], float(e), int(max_rounds), _ptr(R), _ptr(T),
                                   C.cast(C.byref(rounds), C.c_void_p), _ptr(sse), _ptr(jok), max_trace)
    r = min(int(rounds.value), max_trace)
    return {"rc": rc, "R": R.reshape(3, 3), "T": T, "rounds": int(rounds.value), "sse": sse[:r], "jacobi_ok": jok[:r].astype(bool)}


class ReferenceThrows(RuntimeError):
    """The C# as written would throw at this point (VPCO_E_REFERENCE_THROWS)."""


def blocked(mx, my, eps, min_pts, pts_in_cell, shared_objects=False, fast=False, n_threads=1):
    """MainForm.getClusterFromMotor -> DoWork3/StartCode -> CompleteWork3 (vpc_oracle_blocked.cpp).  Returns a dict."""
    mx = np.ascontiguousarray(mx, np.float64); my = np.ascontiguousarray(my, np.float64)
    n = len(mx)
    cid = np.zeros(n, np.int32)
    morder = np.zeros(3 * n + 1, np.int64); mcid = np.zeros(3 * n + 1, np.int32)
    i32 = [C.c_int32(0) for _ in range(5)]      # cluster_sum del_sum rows cols cluster_sum_cells
    i64 = [C.c_int64(0) for _ in range(3)]      # n_unassigned n_shared n_merge
    ref = lambda v: C.cast(C.byref(v), C.c_void_p)   # noqa: E731
    rc = lib().vpco_blocked_literal(_ptr(mx), _ptr(my), n, float(eps), int(min_pts), int(pts_in_cell), int(shared_objects), int(fast), int(n_threads),
                                    _ptr(cid), ref(i32[0]), ref(i32[1]), ref(i32[2]), ref(i32[3]), ref(i64[0]), ref(i64[1]), _ptr(morder), _ptr(mcid),
                                    ref(i64[2]), ref(i32[4]))
    if rc == -7:
        raise ReferenceThrows("the C# throws here (degenerate first cell or clusForMerge[-1])")
    if rc != 0:
        raise RuntimeError(f"oracle blocked rc={rc}")
    k = int(i64[2].value)
    return {"cluster_id": cid, "cluster_sum": int(i32[0].value), "del_sum": int(i32[1].value), "rows": int(i32[2].value), "cols": int(i32[3].value),
            "n_unassigned": int(i64[0].value), "n_shared": int(i64[1].value), "merge_order": morder[:k].copy(), "merge_cid": mcid[:k].copy(),
            "cluster_sum_cells": int(i32[4].value)}


def merge_ids(merge_cid, xyz_entries, mx_entries, my_entries, cluster_amount, thre):
    """Tools.GetClusList -> MergeIDByDistance -> refreshCensAndClusByDictionary on the clusForMerge list (per-ENTRY arrays)."""
    mcid = np.ascontiguousarray(merge_cid, np.int32)
    k = len(mcid)
    xyz = _planar(xyz_entries) if k else np.zeros((3, 0))
    mx = np.ascontiguousarray(mx_entries, np.float64); my = np.ascontiguousarray(my_entries, np.float64)
    new_cid = np.zeros(k, np.int32)
    cap = max(int(cluster_amount), 1)
    dfrom = np.zeros(cap, np.int32); dto = np.zeros(cap, np.int32)
    c5 = np.zeros(5 * cap); cids = np.zeros(cap, np.int32); nc5 = np.zeros(5 * cap)
    amount, nd, ncen = C.c_int32(0), C.c_int32(0), C.c_int32(0)
    ref = lambda v: C.cast(C.byref(v), C.c_void_p)   # noqa: E731
    rc = lib().vpco_merge_ids_literal(_ptr(mcid), _ptr(xyz), _ptr(mx), _ptr(my), k, int(cluster_amount), float(thre), _ptr(new_cid), ref(amount),
                                      _ptr(dfrom), _ptr(dto), ref(nd), _ptr(c5), _ptr(cids), ref(ncen), _ptr(nc5))
    if rc == -7:
        raise ReferenceThrows("the C# throws here (id out of range or Average over an empty cluster)")
    if rc != 0:
        raise RuntimeError(f"oracle merge_ids rc={rc}")
    m, a = int(ncen.value), int(amount.value)
    return {"cluster_id": new_cid, "cluster_amount": a, "dict": list(zip(dfrom[:nd.value].tolist(), dto[:nd.value].tolist())),
            "centers5": c5[:5 * m].reshape(5, m).copy(), "center_ids": cids[:m].copy(), "new_centers5": nc5[:5 * a].reshape(5, a).copy()}
