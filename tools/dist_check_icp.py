"""Multi-GPU ICP parity check (torchrun, one rank per GPU): sharded-model ICP over N GPUs must equal the oracle
on the whole model.  Usage: torchrun --nproc-per-node N tools/dist_check_icp.py [m] [n] [iters]"""
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
from vtkcloudpoint_b200.distributed import GpuIcpBackend, icp_rigid_sharded  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
m = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
n = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 10
model, data, R, T = synth.icp_clouds(0xC3, m, n)
a, b = m * rank // world, m * (rank + 1) // world
ctx = Context(local)
be = GpuIcpBackend(ctx)
tm, td = torch.from_numpy(model[:, a:b].copy()).to(dev), torch.from_numpy(data).to(dev)
for it in range(3):
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    state, order = icp_rigid_sharded(be, tm, a, td, -1.0, iters)
    torch.cuda.synchronize(); dist.barrier(); dt = time.perf_counter() - t0
    if rank == 0:
        print(f"iter {it}: {iters} rounds, {n} x {m} on {world} GPUs: {dt*1e3:.2f} ms ({iters/dt:.0f} iters/s incl. model grid build)", flush=True)
st = state.cpu().numpy()
Ro, To, itd, sse, oo = oracle_py.icp_rigid(model, data, -1.0, iters, n_threads=max(1, (os.cpu_count() or 8) // world))
ok = (int(st[13]) == itd and np.array_equal(order.cpu().numpy(), oo) and np.abs(st[:9].reshape(3, 3) - Ro).max() < 1e-6
      and np.abs(st[9:12] - To).max() < 1e-6 and abs(st[12] - sse) <= 1e-6 * sse)
print(f"rank {rank}: sharded ICP parity vs oracle: {'OK' if ok else 'MISMATCH'} (sse {st[12]:.9g} vs {sse:.9g})", flush=True)
assert ok
ctx.close()
dist.destroy_process_group()
