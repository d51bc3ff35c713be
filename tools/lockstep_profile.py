"""Per-kernel CUDA-event times of the peer-memory paths with the ranks EMULATED on one GPU (phase by phase, no waiting, no NVLink hop):
the pure work of every kernel, to be compared with the multi-GPU per-launch times (which include waiting for the slowest rank).
Usage: python tools/lockstep_profile.py [world]"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from vtkcloudpoint_b200 import Context, synth  # noqa: E402
from vtkcloudpoint_b200.peer import IcpDistPlan, PeerComm, SlabPeerPlan, slab_heap_bytes  # noqa: E402
from peer_helpers import cut_slabs  # noqa: E402

W = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda", 0)


def report(ctxs, title):
    agg = {}
    for c in ctxs:
        for name, ms in c.profile_report():
            agg.setdefault(name, []).append(ms)
    print(title + ": " + "  ".join(f"{k}={1e3 * sum(v) / len(v):.1f}" for k, v in agg.items()), flush=True)


# ---- slab DBSCAN, 1M points per rank
n = 1_000_000 * W
fx, fy = synth.dbscan_cloud(0xC2, int(round((n * 0.784 / 40) ** 0.5)), n_total=n)
order, counts, qs = cut_slabs(fx, fy, W)
sx, sy = fx[order], fy[order]
starts = np.concatenate([[0], np.cumsum(counts)])
ctxs = [Context(0) for _ in range(W)]
cap_h, cap_p = 20000, 40000
comms = PeerComm.local_group(ctxs, slab_heap_bytes(ctxs[0]._lib, W, int(counts.max()), cap_h, cap_p))
bound = float(np.abs(sx + sy).max() + np.abs(sx - sy).max())
plans = [SlabPeerPlan(c, counts.tolist(), qs.tolist(), 0.07, 7, bound, cap_h, cap_p, dev) for c in comms]
for r, p in enumerate(plans):
    p.x.copy_(torch.from_numpy(sx[starts[r]:starts[r + 1]].copy())); p.y.copy_(torch.from_numpy(sy[starts[r]:starts[r + 1]].copy()))
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for it in range(4):
    if it == 2:
        for c in ctxs:
            c.profile(True)
    for ph in range(5):
        for p in plans:
            if ph == 0:
                flush.zero_()
            p.step_phase(ph, 0)
torch.cuda.synchronize()
report(ctxs, f"slab step, {W} emulated ranks x 1M points, per launch [us]")
print("status rank 0:", plans[0].status.cpu().numpy()[:8], "errors", [c.error_bits() for c in comms])
for c in ctxs:
    c.profile(False)
for p in plans:
    p.close()
for c in comms:
    c.close()

# ---- ICP, both splits
for mode, m_tot in ((0, 1_000_000 * W), (1, 1_000_000)):
    model, data, _, _ = synth.icp_clouds(0xC3, m_tot, 100_000, box=100.0 * (m_tot / 1e6) ** (1.0 / 3.0))
    comms = PeerComm.local_group(ctxs, IcpDistPlan.heap_bytes(ctxs[0]._lib, W, 100_000))
    td = torch.from_numpy(data).to(dev)
    plans, keep = [], []
    for r, (c, cm) in enumerate(zip(ctxs, comms)):
        a, b = (m_tot * r // W, m_tot * (r + 1) // W) if mode == 0 else (0, m_tot)
        tm = torch.from_numpy(np.ascontiguousarray(model[:, a:b])).to(dev); keep.append(tm)
        c.icp_set_model_dev(tm)
        plans.append(IcpDistPlan(cm, mode, td, a))
    for p in plans:
        p.begin()
    for it in range(10):
        if it == 4:
            torch.cuda.synchronize()
            for c in ctxs:
                c.profile_report(); c.profile(True)
        for ph in (0, 1, 2):
            if mode == 1 and ph == 1:
                continue
            for p in plans:
                p.round_phase(ph, -1.0, 10)
    torch.cuda.synchronize()
    report(ctxs, f"ICP {'target' if mode == 0 else 'source'} sharded, {W} emulated ranks, 100k x {m_tot}, per launch [us]")
    for c in ctxs:
        c.profile(False)
    for p in plans:
        p.close()
    for c in comms:
        c.close()
