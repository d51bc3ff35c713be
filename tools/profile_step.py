"""One profiled step for ncu: warm up, then run a single DBSCAN (C2) and/or ICP (C3, few rounds) step between
cudaProfilerStart/Stop.  Usage: ncu --profile-from-start off ... python tools/profile_step.py [dbscan|icp] [n]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from vtkcloudpoint_b200 import Context, synth  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "dbscan"
ctx = Context(0)
dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
if what == "dbscan":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
    grid = int(round((n * 0.784 / 40) ** 0.5))
    mx, my = synth.dbscan_cloud(0xC2, grid, n_total=n)
    dx, dy = torch.from_numpy(mx).to(dev), torch.from_numpy(my).to(dev)
    out = None
    for _ in range(3):
        out = ctx.dbscan_dev(dx, dy, 0.07, 7, 0, out=out)
    flush.zero_()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    ctx.dbscan_dev(dx, dy, 0.07, 7, 0, out=out)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("clusters", int(out[3].item()) , "launches", ctx.launch_count)
else:
    model, data, _, _ = synth.icp_clouds(0xC3, 1_000_000, 100_000)
    dm, dd = torch.from_numpy(model).to(dev), torch.from_numpy(data).to(dev)
    ctx.icp_set_model_dev(dm)
    outs = ctx.icp_rigid_dev(dd, -1.0, 5)
    flush.zero_()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    ctx.icp_set_model_dev(dm)
    ctx.icp_rigid_dev(dd, -1.0, 3, out=outs)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("state", outs[0].cpu().numpy()[12:15], "launches", ctx.launch_count)
ctx.close()
