"""Single-GPU DBSCAN throughput vs cloud size (C2 recipe at constant density), device-resident, CUDA events."""
import sys
from pathlib import Path
import numpy as np
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from vtkcloudpoint_b200 import Context, synth

ctx = Context(0)
dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for n in [int(a) for a in sys.argv[1:] if a.isdigit()] or [1_000_000, 4_000_000, 16_000_000]:
    grid = int(round((n * 0.784 / 40) ** 0.5))
    xs, ys = [], []
    for s in range(0, n, 4_000_000):
        mx, my = synth.dbscan_cloud(0xC4, grid, n_total=n, start=s, count=min(4_000_000, n - s))
        xs.append(torch.from_numpy(mx).to(dev)); ys.append(torch.from_numpy(my).to(dev))
    dx, dy = torch.cat(xs), torch.cat(ys)
    del xs, ys
    out = None
    for _ in range(3):
        out = ctx.dbscan_dev(dx, dy, 0.07, 7, 0, out=out)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ctx.dbscan_dev(dx, dy, 0.07, 7, 0, out=out); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = sorted(ts)[len(ts) // 2]
    print(f"n={n}: {t:.3f} ms  {n / t / 1e3:.1f} Mpts/s  clusters={int(out[3].item())}  ({21 * n / t / 1e6 / 6549.4 * 100:.2f}% of HBM roofline at 21 B/pt)", flush=True)
    ctx.profile(True)
    ctx.dbscan_dev(dx, dy, 0.07, 7, 0, out=out)
    rep = ctx.profile_report()
    ctx.profile(False)
    print("   " + "  ".join(f"{k.replace('k_db_', '').replace('k_scan_exclusive', 'scan')}={v * 1e3:.0f}us" for k, v in rep), flush=True)
    del dx, dy, out
    torch.cuda.empty_cache()
ctx.close()
