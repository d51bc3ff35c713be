"""Helpers for the peer-memory multi-GPU tests: slab cutting of a cloud and the phase-by-phase (lockstep) emulation of `world`
ranks on ONE GPU -- every phase is enqueued for all ranks before the next one, so a kernel that waits always finds its flag set
(B200_PROFILING.md: ranks that wait on each other must not run as separate launches on one GPU unless ordered like this)."""
import numpy as np
import torch

from vtkcloudpoint_b200 import Context
from vtkcloudpoint_b200.peer import IcpDistPlan, PeerComm, SlabPeerPlan, slab_heap_bytes


def cut_slabs(fx, fy, world):
    """Quantile cut of u = x + y into `world` slabs; non-finite points are dealt round-robin.  Returns (order, counts, splitters):
    order = permutation that groups the cloud by rank (global index = position in that order)."""
    fu = fx + fy
    ok = np.isfinite(fx) & np.isfinite(fy) & np.isfinite(fu) & np.isfinite(fx - fy)
    qs = np.quantile(fu[ok], [j / world for j in range(1, world)]) if world > 1 else np.empty(0)
    band = np.searchsorted(qs, np.where(ok, fu, 0.0), side="right")
    bad = np.flatnonzero(~ok)
    band[bad] = np.arange(len(bad)) % world
    order = np.argsort(band, kind="stable")
    counts = np.bincount(band, minlength=world)
    return order, counts, qs


def run_slabs_lockstep(fx, fy, world, eps, min_pts, first_cluster_id=0, steps=2, cap_frac=1.0, device=0):
    """Exact DBSCAN of the whole cloud through `world` emulated ranks on one GPU.  Returns (cid, key, cls, amount, status list)
    in the ORIGINAL point order."""
    dev = torch.device("cuda", device)
    order, counts, qs = cut_slabs(fx, fy, world)
    sx, sy = fx[order], fy[order]
    fin = np.isfinite(sx + sy) & np.isfinite(sx - sy)
    bound = float(np.abs(sx[fin] + sy[fin]).max() + np.abs(sx[fin] - sy[fin]).max()) if fin.any() else 1.0
    n_max = int(counts.max())
    cap_h, cap_p = max(1024, int(n_max * cap_frac)), max(1024, int(2 * n_max * cap_frac))
    ctxs = [Context(device) for _ in range(world)]
    comms = PeerComm.local_group(ctxs, slab_heap_bytes(ctxs[0]._lib, world, n_max, cap_h, cap_p))
    plans = [SlabPeerPlan(c, counts.tolist(), qs.tolist(), eps, min_pts, bound, cap_h, cap_p, dev) for c in comms]
    starts = np.concatenate([[0], np.cumsum(counts)])
    for r, p in enumerate(plans):
        p.x.copy_(torch.from_numpy(sx[starts[r]:starts[r + 1]].copy()))
        p.y.copy_(torch.from_numpy(sy[starts[r]:starts[r + 1]].copy()))
    for _ in range(steps):                       # several steps: epochs, parity buffers and re-armed counters are exercised
        for ph in range(5):
            for p in plans:
                p.step_phase(ph, first_cluster_id)
    torch.cuda.synchronize()
    n = len(fx)
    cid, key, cls = np.zeros(n, np.int32), np.zeros(n, np.uint8), np.zeros(n, np.uint8)
    status = []
    for r, p in enumerate(plans):
        idx = order[starts[r]:starts[r + 1]]
        cid[idx] = p.cluster_id.cpu().numpy(); key[idx] = p.is_key.cpu().numpy(); cls[idx] = p.is_classed.cpu().numpy()
        status.append(p.status.cpu().numpy().copy())
    errs = [c.error_bits() for c in comms]
    for p in plans:
        p.close()
    for c in comms:
        c.close()
    for c in ctxs:
        c.close()
    # ids are ranks of minimum GLOBAL (slab-order) core indices: renumber to the original order's minima for the comparison
    return cid, key, cls, int(status[0][0]), status, errs, order


def canon(cid):
    """Cluster partition canonicalised by the minimum point index (BASELINE.json: labels canonicalised by the minimum point index)."""
    out = np.zeros_like(cid)
    nz = np.flatnonzero(cid > 0)
    if len(nz) == 0:
        return out
    first = {}
    for i in nz:
        first.setdefault(int(cid[i]), i)
    ranks = {c: k + 1 for k, (c, _) in enumerate(sorted(first.items(), key=lambda kv: kv[1]))}
    out[nz] = [ranks[int(c)] for c in cid[nz]]
    return out


def run_icp_lockstep(model, data, world, mode, e, max_iters, device=0):
    """ICP through `world` emulated ranks on one GPU (mode 0: target sharded, 1: source sharded).  Returns per-rank (state, order)."""
    dev = torch.device("cuda", device)
    m, n = model.shape[1], data.shape[1]
    ctxs = [Context(device) for _ in range(world)]
    comms = PeerComm.local_group(ctxs, IcpDistPlan.heap_bytes(ctxs[0]._lib, world, n))
    td = torch.from_numpy(np.ascontiguousarray(data)).to(dev)
    plans, keep = [], []
    for r, (c, cm) in enumerate(zip(ctxs, comms)):
        a, b = (m * r // world, m * (r + 1) // world) if mode == 0 else (0, m)
        tm = torch.from_numpy(np.ascontiguousarray(model[:, a:b])).to(dev)
        keep.append(tm)
        c.icp_set_model_dev(tm)
        plans.append(IcpDistPlan(cm, mode, td, a))
    for p in plans:
        p.begin()
    for _ in range(max_iters):
        for ph in (0, 1, 2):
            if mode == 1 and ph == 1:
                continue
            for p in plans:
                p.round_phase(ph, e, max_iters)
    outs = []
    for p in plans:
        st, order = p.export()
        outs.append((st.cpu().numpy().copy(), order.cpu().numpy().copy()))
    torch.cuda.synchronize()
    errs = [c.error_bits() for c in comms]
    for p in plans:
        p.close()
    for c in comms:
        c.close()
    for c in ctxs:
        c.close()
    return outs, errs
