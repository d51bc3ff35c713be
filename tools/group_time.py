"""One process, N GPUs through vpc_create(n_devices = N): timing of the host-pointer calls (pageable NumPy arrays) and a phase trace.
Usage: python tools/group_time.py [n_gpus] [points]"""
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle"))
from vtkcloudpoint_b200 import Context, DbscanResult, synth  # noqa: E402

W = int(sys.argv[1]) if len(sys.argv) > 1 else 2
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000 * W
grid = int(round((n * 0.784 / 40) ** 0.5))
mx, my = synth.dbscan_cloud(0xC2, grid, n_total=n)
for devs in ([0], list(range(W))):
    ctx = Context(devs if len(devs) > 1 else devs[0])
    res = DbscanResult(np.empty(n, np.int32), np.empty(n, np.uint8), np.empty(n, np.uint8), 0)
    for _ in range(3):
        ctx.dbscan(mx, my, 0.07, 7, 0, out=res)
    if len(devs) > 1:
        os.environ["VPC_GROUP_TRACE"] = "1"
    t0 = time.perf_counter()
    reps = 10
    for _ in range(reps):
        ctx.dbscan(mx, my, 0.07, 7, 0, out=res)
    dt = (time.perf_counter() - t0) / reps
    print(f"devices {devs}: {n} points, pageable arrays: {dt*1e3:.3f} ms per call ({n/dt/1e6:.1f} Mpts/s), clusters {res.cluster_amount}", flush=True)
    if len(devs) == 1:
        ref = (res.cluster_id.copy(), res.is_key.copy(), res.cluster_amount)
    else:
        ok = np.array_equal(ref[0], res.cluster_id) and np.array_equal(ref[1], res.is_key) and ref[2] == res.cluster_amount
        print("multi-device result equals the single-device result:", "OK" if ok else "MISMATCH", flush=True)
    ctx.close()
