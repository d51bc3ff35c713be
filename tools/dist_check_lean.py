"""Multi-GPU parity + timing of the sync-free pre-cut slab path (torchrun, one rank per GPU).
Usage: torchrun --nproc-per-node N tools/dist_check_lean.py [points_per_gpu]"""
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle"))
import oracle_py  # noqa: E402  (checker)
from vtkcloudpoint_b200 import Context, synth  # noqa: E402
from vtkcloudpoint_b200.distributed import calibrated_lean_plan, dbscan_slabs_lean  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
per = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
n = per * world
grid = int(round((n * 0.784 / 40) ** 0.5))
fx, fy = synth.dbscan_cloud(0xC2, grid, n_total=n)
fu = fx + fy
qs = np.quantile(fu, [j / world for j in range(1, world)]) if world > 1 else np.empty(0)
band = np.searchsorted(qs, fu, side="right")
order = np.argsort(band, kind="stable")            # global index = position in slab order
fx, fy, band = fx[order], fy[order], band[order]
a, b = int(np.searchsorted(band, rank, "left")), int(np.searchsorted(band, rank, "right"))
ctx = Context(local)
tx, ty = torch.from_numpy(fx[a:b].copy()).to(dev), torch.from_numpy(fy[a:b].copy()).to(dev)
plan = calibrated_lean_plan(ctx, tx, ty, a, list(qs), 0.07, float(np.abs(fu).max() + np.abs(fx - fy).max()), 7, dev)
if rank == 0:
    print(f"calibrated capacities: halo {plan.cap}, pairs {plan.cap_pairs}, heads {plan.cap_heads}", flush=True)
for it in range(6):
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    cid, key, cls, amount, overflow = dbscan_slabs_lean(plan, tx, ty, a, 7, 0)
    torch.cuda.synchronize(); dist.barrier(); dt = time.perf_counter() - t0
    if rank == 0:
        print(f"iter {it}: {n} pts on {world} GPUs in {dt*1e3:.3f} ms ({n/dt/1e6:.1f} Mpts/s) clusters={int(amount.item())} overflow={int(overflow.item())}", flush=True)
if n <= 20_000_000:
    ocid, okey, ocls, oamount = oracle_py.dbscan(fx, fy, 0.07, 7, 0, variant="grid", n_threads=max(1, (os.cpu_count() or 8) // world))
    ok = (int(amount.item()) == oamount and int(overflow.item()) == 0 and np.array_equal(cid.cpu().numpy(), ocid[a:b])
          and np.array_equal(key.cpu().numpy(), okey[a:b]) and np.array_equal(cls.cpu().numpy(), ocls[a:b]))
    print(f"rank {rank}: lean slab path vs oracle on the whole cloud: {'OK' if ok else 'MISMATCH'} (clusters {int(amount.item())} vs {oamount})", flush=True)
    assert ok
else:
    # full size (config C4): every rank clusters the WHOLE slab-ordered cloud on its own GPU and compares its slab, bit for bit
    del tx, ty
    wx, wy = torch.from_numpy(fx).to(dev), torch.from_numpy(fy).to(dev)
    scid, skey, scls, samount = ctx.dbscan_dev(wx, wy, 0.07, 7, 0)
    ok = (int(samount.item()) == int(amount.item()) and int(overflow.item()) == 0 and bool((scid[a:b] == cid).all())
          and bool((skey[a:b] == key).all()) and bool((scls[a:b] == cls).all()))
    print(f"rank {rank}: {world}-GPU lean slab path vs single-GPU clustering of the whole {n}-point cloud: {'OK' if ok else 'MISMATCH'} "
          f"(clusters {int(amount.item())} vs {int(samount.item())})", flush=True)
    assert ok
ctx.close()
dist.destroy_process_group()
