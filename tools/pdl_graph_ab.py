"""Does programmatic dependent launch survive stream capture?  One GPU, the C2 DBSCAN step issued eagerly and replayed as a CUDA graph;
run once with VPC_PDL=1 and once with VPC_PDL=0.  Usage: python tools/pdl_graph_ab.py [n]"""
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from vtkcloudpoint_b200 import Context, synth  # noqa: E402
from vtkcloudpoint_b200.peer import GraphedStep  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
ctx = Context(0)
dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
grid = int(round((n * 0.784 / 40) ** 0.5))
mx, my = synth.dbscan_cloud(0xC2, grid, n_total=n)
dx, dy = torch.from_numpy(mx).to(dev), torch.from_numpy(my).to(dev)
out = ctx.dbscan_dev(dx, dy, 0.07, 7, 0)


def timed(fn, reps=40):
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / reps


eager = lambda: ctx.dbscan_dev(dx, dy, 0.07, 7, 0, out=out)  # noqa: E731
for _ in range(5):
    eager()
t_e = timed(eager)
g = GraphedStep(eager, dev)
for _ in range(5):
    g.replay()
t_g = timed(g.replay)
print(f"VPC_PDL={os.environ.get('VPC_PDL', '1')} n={n}: eager {t_e * 1e3:.1f} us/step, graph replay {t_g * 1e3:.1f} us/step, clusters {int(out[3].item())}")
ctx.close()
