"""Multi-GPU run of the C5 pipeline (run under torchrun, one rank per GPU): DBSCAN across the GPUs -> cluster-sharded centroids and
bounding circles -> radius filter -> sharded ICP of the centres to the truth pattern.  Up to 2M points the result is compared with the
oracle pipeline on the whole cloud; above that the size-independent properties are checked (every rank agrees, cluster count, the
ICP recovers the planted rigid motion).  Usage: torchrun --nproc-per-node N tools/dist_check_pipeline.py [n_points]"""
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "oracle"), str(ROOT / "tests")):
    sys.path.insert(0, p)
import pipeline_ref  # noqa: E402  (checker)
from vtkcloudpoint_b200 import Context  # noqa: E402
from vtkcloudpoint_b200.pipeline import GpuPipelineBackend, run_pipeline  # noqa: E402

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
grid = int(round((n * 0.784 / 40) ** 0.5))
mx, my, xyz = pipeline_ref.scene(0xC5, grid, n)                     # every rank generates the scene, keeps its chunk
a, b = n * rank // world, n * (rank + 1) // world
# truth = the planted centres' XYZ by the import formulas, moved rigidly (so the expected R, T are known at any size)
gx, gy = np.meshgrid(np.arange(grid), np.arange(grid), indexing="ij")
import oracle_py  # noqa: E402
cxyz, _ = oracle_py.polar_to_xyz(149.0 + 0.5 * gx.ravel(), 307.0 + 0.5 * gy.ravel(), np.full(grid * grid, 41.91), 149.0, 307.0)
truth = pipeline_ref.truth_for(cxyz[:2])
ctx = Context(local)
be = GpuPipelineBackend(ctx)
t = lambda v: torch.from_numpy(np.ascontiguousarray(v)).to(dev)    # noqa: E731
tx, ty, txyz, ttruth = t(mx[a:b]), t(my[a:b]), t(xyz[:, a:b]), t(truth)
kw = dict(eps=0.07, min_pts=7, radius_threshold=0.088, icp_e=1e-9, icp_max_iters=10, match_distance=0.05)
for it in range(3):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    res = run_pipeline(be, tx, ty, txyz, a, ttruth, **kw)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    dt = time.perf_counter() - t0
    if rank == 0:
        st = res.icp_state.cpu().numpy()
        print(f"iter {it}: {n} pts on {world} GPU(s): {dt * 1e3:.1f} ms ({n / dt / 1e6:.1f} Mpts/s)  clusters={res.cluster_amount} "
              f"kept={res.kept_ids.numel()} filtered={int(res.filtered.sum())} icp_iters={int(st[13])} rmse={np.sqrt(st[12] / max(res.kept_ids.numel(), 1)):.3e} "
              f"matched={(res.matched >= 0).float().mean().item():.4f}", flush=True)
if n <= 2_000_000:
    ref = pipeline_ref.run(mx, my, xyz, truth, kw["eps"], kw["min_pts"], kw["radius_threshold"], kw["icp_e"], kw["icp_max_iters"], kw["match_distance"])
    pipeline_ref.check(res, ref, a, b)
    print(f"rank {rank}: pipeline vs oracle on the whole cloud: OK (clusters {ref['amount']}, kept {len(ref['kept'])}, filtered {int(ref['filtered'].sum())})", flush=True)
else:
    st = res.icp_state.cpu().numpy()
    th = np.deg2rad(0.4)
    assert abs(st[0] - np.cos(th)) < 1e-4 and abs(st[3] - np.sin(th)) < 1e-4 and abs(st[9] - 0.011) < 5e-3 and abs(st[10] + 0.007) < 5e-3, st[:12]
    assert (res.matched >= 0).float().mean().item() > 0.98
    if world > 1:   # replicated results must be identical on every rank
        chk = torch.stack([res.centres[:, 1:].nan_to_num().sum(), res.circle[2, 1:].sum(), res.icp_state.sum(), res.kept_ids.sum().double()])
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert torch.equal(lo, hi)
    print(f"rank {rank}: properties OK (planted motion recovered, ranks agree)", flush=True)
ctx.close()
if world > 1:
    dist.destroy_process_group()
