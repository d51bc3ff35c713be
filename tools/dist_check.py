"""Multi-GPU parity check (run under torchrun, one rank per GPU): dbscan_slabs over N GPUs must equal the
oracle on the whole cloud.  Usage: torchrun --nproc-per-node N tools/dist_check.py [n_points]"""
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
from vtkcloudpoint_b200.distributed import GpuBackend, dbscan_slabs  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
grid = int(round((n * 0.784 / 40) ** 0.5))
a, b = n * rank // world, n * (rank + 1) // world
mx, my = synth.dbscan_cloud(0xC4, grid, n_total=n, start=a, count=b - a)
ctx = Context(local)
be = GpuBackend(ctx)
tx, ty = torch.from_numpy(mx).to(dev), torch.from_numpy(my).to(dev)
for it in range(3):
    stats = {}
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    cid, key, cls, amount = dbscan_slabs(be, tx, ty, a, 0.07, 7, 0, stats=stats)
    torch.cuda.synchronize(); dist.barrier(); dt = time.perf_counter() - t0
    if rank == 0:
        print(f"iter {it}: {n} pts on {world} GPUs in {dt*1e3:.2f} ms  ({n/dt/1e6:.1f} Mpts/s) clusters={amount} stats={stats}", flush=True)
if n <= 20_000_000:
    fx, fy = synth.dbscan_cloud(0xC4, grid, n_total=n)
    ocid, okey, ocls, oamount = oracle_py.dbscan(fx, fy, 0.07, 7, 0, variant="grid", n_threads=max(1, (os.cpu_count() or 8) // world))
    ok = (amount == oamount and np.array_equal(cid.cpu().numpy(), ocid[a:b]) and np.array_equal(key.cpu().numpy(), okey[a:b])
          and np.array_equal(cls.cpu().numpy(), ocls[a:b]))
    print(f"rank {rank}: parity vs oracle on the whole cloud: {'OK' if ok else 'MISMATCH'} (clusters {amount} vs {oamount})", flush=True)
    assert ok
else:
    # full size (config C4): the oracle cannot finish this, so every rank clusters the WHOLE cloud on its own GPU with the
    # single-GPU entry point and compares its chunk of the distributed result with it, bit for bit
    del tx, ty
    xs, ys = [], []
    for s0 in range(0, n, 5_000_000):
        fx, fy = synth.dbscan_cloud(0xC4, grid, n_total=n, start=s0, count=min(5_000_000, n - s0))
        xs.append(torch.from_numpy(fx).to(dev)); ys.append(torch.from_numpy(fy).to(dev))
    wx, wy = torch.cat(xs), torch.cat(ys)
    del xs, ys
    scid, skey, scls, samount = ctx.dbscan_dev(wx, wy, 0.07, 7, 0)
    ok = (int(samount.item()) == amount and bool((scid[a:b] == cid).all()) and bool((skey[a:b] == key).all()) and bool((scls[a:b] == cls).all()))
    print(f"rank {rank}: {world}-GPU result vs single-GPU clustering of the whole {n}-point cloud: {'OK' if ok else 'MISMATCH'} (clusters {amount} vs {int(samount.item())})", flush=True)
    assert ok
ctx.close()
dist.destroy_process_group()
