"""Timings of the steps AROUND the hot path on cuda:0, device-resident, CUDA events, median of 5 after 2
warm-ups: stable radix sort, cluster grouping, ordered centroids, bounding circles, radius filter, 2-D nearest truth, polar->XYZ,
duplicate removal, text ingest.  Algorithmic bytes = what the step must read and write once; GB/s = those bytes / time.
Usage: python tools/time_around.py [n_points]   (default 10M, the C5 size)"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from vtkcloudpoint_b200 import Context, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
dev = torch.device("cuda", 0)
ctx = Context(0)
grid = int(round((n * 0.784 / 40) ** 0.5))
xs, ys = [], []
for s in range(0, n, 5_000_000):
    mx, my = synth.dbscan_cloud(0xC5, grid, n_total=n, start=s, count=min(5_000_000, n - s))
    xs.append(mx); ys.append(my)
mx, my = np.concatenate(xs), np.concatenate(ys)
dist = 41.91 + 0.004 * (synth.uniform(0xC5, 40, np.arange(n, dtype=np.uint64)) - 0.5)
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)   # noqa: E731
d_mx, d_my, d_dist = t(mx), t(my), t(dist)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(name, fn, algo_bytes, reps=5):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    print(f"{name:42s} {ms:9.3f} ms   {algo_bytes / ms / 1e6:8.1f} GB/s algorithmic   ({algo_bytes / 1e6:.0f} MB)", flush=True)
    return out


print(f"n = {n} points, {grid * grid} planted clusters", flush=True)
cid, key, cls, amount = ctx.dbscan_dev(d_mx, d_my, 0.07, 7, 0)
k = int(amount.item())
xyz, keep = timed("polar -> XYZ + distance gate", lambda: ctx.polar_to_xyz_dev(d_mx, d_my, d_dist, 149.0, 307.0), n * (24 + 24 + 1))
keys64 = (torch.rand(n, device=dev, dtype=torch.float64) * 2 ** 40).to(torch.int64)
timed("stable radix sort, 40-bit keys (5 passes)", lambda: ctx.sort_pairs_dev(keys64.clone(), None, 0, 40), n * 12 * 2)
timed("argsort of doubles (8 passes)", lambda: ctx.argsort_f64_dev(d_mx), n * (8 + 4))
members, offsets = timed(f"group points by cluster id ({k} clusters)", lambda: ctx.cluster_groups_dev(cid, k), n * (4 + 4))
vals5 = torch.stack([xyz[0], xyz[1], xyz[2], d_mx, d_my]).contiguous()
timed("ordered centroids (5 fields)", lambda: ctx.cluster_means_ordered_dev(members, offsets, k, vals5), n * (4 + 40))
circ, status = timed("minimal bounding circles (hull + search)", lambda: ctx.cluster_circles_dev(members, offsets, k, vals5[0], vals5[1]), n * (4 + 16))
timed("radius filter", lambda: ctx.radius_filter_dev(circ[2].contiguous(), status, k, 0.088), (k + 1) * 13)
gx, gy = np.meshgrid(np.arange(grid), np.arange(grid), indexing="ij")
truth = torch.stack([t(149.0 + 0.5 * gx.ravel()), t(307.0 + 0.5 * gy.ravel()), torch.zeros(grid * grid, dtype=torch.float64, device=dev)]).contiguous()
ctx.icp_set_model_dev(truth)
tid = torch.arange(1, grid * grid + 1, dtype=torch.int32, device=dev)
timed(f"2-D nearest truth ({grid * grid} truths, radius 0.088)", lambda: ctx.nearest_truth_2d_dev(tid, d_mx, d_my, 0.088), n * (16 + 4) + grid * grid * 24)
timed("duplicate removal (hash set on X, Y, Z)", lambda: ctx.dedupe_xyz_dev(xyz, keep), n * (24 + 1 + 1))
rows = min(n, 2_000_000)
text = ("motor_x\tmotor_y\tDistance\n" + "".join(f"{a:.3f}\t{b:.3f}\t{c:.3f}\n" for a, b, c in zip(mx[:rows], my[:rows], dist[:rows]))).encode()
import time
ctx.ingest_text(text, 149.0, 307.0)                       # first call sizes the arenas
t0 = time.perf_counter(); r = ctx.ingest_text(text, 149.0, 307.0); dt = time.perf_counter() - t0
print(f"{'text ingest incl. H2D/D2H (' + str(rows) + ' rows)':42s} {dt * 1e3:9.3f} ms   {len(text) / dt / 1e9:8.2f} GB/s of text   ({len(text) / 1e6:.0f} MB, {r['n_kept']} kept, {r['n_duplicates']} duplicates; pageable host buffers)")
ctx.profile(True)
ctx.ingest_text(text, 149.0, 307.0)
rep = ctx.profile_report(); ctx.profile(False)
print("   kernels: " + "  ".join(f"{k.replace('k_in_', '').replace('k_scan_exclusive', 'scan')}={v * 1e3:.0f}us" for k, v in rep), flush=True)
ctx.close()
