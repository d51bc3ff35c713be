"""A small pass over every host-pointer entry point, meant to run under compute-sanitizer:
   compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitize_run.py
(no torch: NumPy arrays through the C ABI only, so that the tool sees libvpc's kernels and nothing else)."""
import os
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from vtkcloudpoint_b200 import Context, synth  # noqa: E402

rng = np.random.default_rng(3)
ctx = Context(0)
mx, my = synth.dbscan_cloud(0xC1, 14, n_total=10_000, decimals=3)
r = ctx.dbscan(mx, my, 0.07, 7)
print("dbscan", r.cluster_amount)
os.environ["VPC_DB_BAND_MIN"] = "1"
rb = ctx.dbscan(mx, my, 0.07, 7)
os.environ.pop("VPC_DB_BAND_MIN")
assert np.array_equal(r.cluster_id, rb.cluster_id)
x = rng.uniform(0, 1, 3000); y = rng.uniform(0, 1, 3000); x[5] = np.nan; y[9] = np.inf
for eps, mp in ((0.05, 4), (0.05, 0), (-1.0, 2), (0.0, 1), (5.0, 3)):
    ctx.dbscan(x, y, eps, mp, 3)
off = np.array([0, 700, 700, 1900, 3000], np.int64)
ctx.dbscan_cells(x, y, off, 0.05, 3)
print("blocked", ctx.dbscan_blocked_ref(mx, my, 0.07, 7, 200)["cluster_sum"])
model, data, _, _ = synth.icp_clouds(0xC3, 6000, 900)
print("icp", ctx.icp_rigid(model, data, -1.0, 4).iters_done)
ctx.closest_point_set(model, data)
ctx.match_within(model, data, 0.5)
xyz = np.stack([40 * (mx - 149), 40 * (my - 307), np.zeros_like(mx)])
st = ctx.cluster_stats(r.cluster_id, r.cluster_amount, xyz, mx, my)
print("stats", int((st["status3d"] == 1).sum()))
big = np.ones(4000, np.int32); bx = rng.normal(size=4000); by = rng.normal(size=4000)
ctx.cluster_stats(big, 1, np.stack([bx, by, bx]), by, bx)
ctx.nearest_truth_2d(st["means"][3, 1:], st["means"][4, 1:], None, mx, my, 0.088)
text = ("h\n" + "".join(f"{a:.3f}\t{b:.3f}\t{41.9:.3f}\n" for a, b in zip(mx[:3000], my[:3000])) + "bad\t1\n1.5\t2.5\t0").encode()
print("ingest", ctx.ingest_text(text, 149.0, 307.0)["n_kept"])
ctx.close()
print("sanitize run ok")
