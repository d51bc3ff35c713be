"""Multi-GPU parity + timing of the PEER-MEMORY paths (one process per GPU, torchrun): slab DBSCAN of one pre-cut cloud and ICP with
the target / the source sharded, each compared with the CPU oracle on the whole problem (or, at full size, with the single-GPU entry
point on the whole cloud).  Exits non-zero on any mismatch.
Usage: torchrun --nproc-per-node N tools/dist_check_peer.py [points_per_gpu] [icp_m] [icp_n] [icp_iters]"""
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "oracle")); sys.path.insert(0, str(ROOT / "tests"))
import oracle_py  # noqa: E402  (checker)
from vtkcloudpoint_b200 import Context, synth  # noqa: E402
from vtkcloudpoint_b200.peer import GraphedStep, IcpDistPlan, PeerComm, calibrated_slab_plan  # noqa: E402
from peer_helpers import cut_slabs  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
per = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
icp_m = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
icp_n = int(sys.argv[3]) if len(sys.argv) > 3 else 100_000
icp_iters = int(sys.argv[4]) if len(sys.argv) > 4 else 10
ok_all = True
ctx = Context(local)
threads = max(1, (os.cpu_count() or 8) // world)


def say(msg):
    print(f"[rank {rank}] {msg}", flush=True)


def timed(fn, reps=10):
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ---- DBSCAN over slabs ----------------------------------------------------------------------------------------------
n = per * world
grid = int(round((n * 0.784 / 40) ** 0.5))
fx, fy = synth.dbscan_cloud(0xC2, grid, n_total=n)
if n <= 4_000_000:                       # a few non-finite owned points (noise, DBImproved.cs:41)
    fx[::100_003] = np.nan; fy[7::200_003] = np.inf
order, counts, qs = cut_slabs(fx, fy, world)
sx, sy = fx[order], fy[order]
starts = np.concatenate([[0], np.cumsum(counts)])
a, b = int(starts[rank]), int(starts[rank + 1])
fin = np.isfinite(sx + sy) & np.isfinite(sx - sy)
bound = float(np.abs(sx[fin] + sy[fin]).max() + np.abs(sx[fin] - sy[fin]).max())
tx, ty = torch.from_numpy(sx[a:b].copy()).to(dev), torch.from_numpy(sy[a:b].copy()).to(dev)
comm, plan = calibrated_slab_plan(ctx, tx, ty, counts.tolist(), qs.tolist(), 0.07, 7, bound, dev)
if rank == 0:
    say(f"slab plan: {world} ranks, capacities halo {plan.cap_halo} pairs {plan.cap_pairs}")
ms_eager = timed(lambda: plan.step(0), 5)
g = GraphedStep(lambda: plan.step(0), dev)
ms_graph = timed(g.replay, 20)
cid, key, cls, status = g.replay()
torch.cuda.synchronize()
st = status.cpu().numpy()
amount = int(st[0])
if rank == 0:
    say(f"DBSCAN {n} pts on {world} GPUs: {ms_eager:.3f} ms/step eager, {ms_graph:.3f} ms/step replayed ({n / ms_graph / 1e3:.1f} Mpts/s); clusters {amount}, error bits {int(st[1])}")
if n <= 20_000_000:
    ocid, okey, ocls, oamount = oracle_py.dbscan(sx, sy, 0.07, 7, 0, variant="grid", n_threads=threads)
    ok = (amount == oamount and int(st[1]) == 0 and np.array_equal(cid.cpu().numpy(), ocid[a:b]) and np.array_equal(key.cpu().numpy(), okey[a:b])
          and np.array_equal(cls.cpu().numpy(), ocls[a:b]))
    say(f"peer slab DBSCAN vs oracle on the whole cloud: {'OK' if ok else 'MISMATCH'} (clusters {amount} vs {oamount})")
else:
    wx, wy = torch.from_numpy(sx).to(dev), torch.from_numpy(sy).to(dev)
    scid, skey, scls, samount = ctx.dbscan_dev(wx, wy, 0.07, 7, 0)
    ok = (int(samount.item()) == amount and int(st[1]) == 0 and bool((scid[a:b] == cid).all()) and bool((skey[a:b] == key).all()) and bool((scls[a:b] == cls).all()))
    say(f"peer slab DBSCAN vs single-GPU clustering of the whole {n}-point cloud: {'OK' if ok else 'MISMATCH'}")
    del wx, wy, scid, skey, scls
ok_all &= ok
del g
plan.close(); comm.close()

# ---- ICP: target sharded (weak: the model is icp_m per GPU) and source sharded (strong: the C3 problem itself) ---------------
for mode, m_tot in ((0, icp_m * world), (1, icp_m)):
    model, data, _, _ = synth.icp_clouds(0xC3, m_tot, icp_n, box=100.0 * (m_tot / 1e6) ** (1.0 / 3.0))
    a_, b_ = (m_tot * rank // world, m_tot * (rank + 1) // world) if mode == 0 else (0, m_tot)
    tm = torch.from_numpy(np.ascontiguousarray(model[:, a_:b_])).to(dev)
    td = torch.from_numpy(data).to(dev)
    icomm = PeerComm.connected(ctx, IcpDistPlan.heap_bytes(ctx._lib, world, icp_n), dev)
    ctx.icp_set_model_dev(tm)
    ip = IcpDistPlan(icomm, mode, td, a_)
    ms_eager = timed(lambda: ip.run(-1.0, icp_iters), 3)
    if rank == 0:                        # per-kernel CUDA-event times of one eager run (waiting kernels include the wait)
        ctx.profile(True); ip.run(-1.0, icp_iters); rep = ctx.profile_report(); ctx.profile(False)
        agg = {}
        for kname, ms in rep:
            agg.setdefault(kname, []).append(ms)
        say("  per launch [us]: " + "  ".join(f"{k}={1e3 * sum(v) / len(v):.1f}" for k, v in agg.items()))
    else:
        ip.run(-1.0, icp_iters)
    ig = GraphedStep(lambda: ip.run(-1.0, icp_iters), dev)
    ms_graph = timed(ig.replay, 5)
    state, order = ig.replay()
    torch.cuda.synchronize()
    stv = state.cpu().numpy()
    name = "target sharded" if mode == 0 else "source sharded"
    if rank == 0:
        say(f"ICP {name}: {icp_n} x {m_tot} on {world} GPUs, {icp_iters} rounds: {ms_eager / icp_iters * 1e3:.1f} us/round eager, {ms_graph / icp_iters * 1e3:.1f} us/round replayed "
            f"({icp_iters / ms_graph * 1e3:.0f} iters/s)")
    Ro, To, itd, sse, oo = oracle_py.icp_rigid(model, data, -1.0, icp_iters, n_threads=threads)
    ok = (int(stv[13]) == itd and np.array_equal(order.cpu().numpy(), oo) and np.abs(stv[:9].reshape(3, 3) - Ro).max() < 1e-6
          and np.abs(stv[9:12] - To).max() < 1e-6 and abs(stv[12] - sse) <= 1e-6 * sse)
    say(f"ICP {name} vs oracle: {'OK' if ok else 'MISMATCH'} (sse {stv[12]:.9g} vs {sse:.9g})")
    ok_all &= ok
    del ig
    if m_tot <= 2_000_000:               # non-finite corner cases (ICP.cs:233-244): a NaN data point, a NaN model point in the last shard
        data2, model2 = data.copy(), model.copy()
        data2[1, 17] = np.nan; model2[2, m_tot - 5] = np.nan
        tm.copy_(torch.from_numpy(np.ascontiguousarray(model2[:, a_:b_]))); td.copy_(torch.from_numpy(data2))
        ctx.icp_set_model_dev(tm)
        ip.begin(); _, o1 = ip.rounds(-1.0, 1)
        torch.cuda.synchronize()
        oo1, _ = oracle_py.closest_point_set(model2, data2, "grid", n_threads=threads)
        ok = np.array_equal(o1.cpu().numpy(), oo1)
        say(f"ICP {name}, non-finite points, vs oracle: {'OK' if ok else 'MISMATCH'}")
        ok_all &= ok
    ip.close(); icomm.close()

flag = torch.tensor([0 if ok_all else 1], device=dev)
dist.all_reduce(flag)
ctx.close()
dist.destroy_process_group()
sys.exit(1 if int(flag.item()) else 0)
