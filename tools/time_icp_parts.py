"""Times the pieces of an ICP round on cuda:0 (CUDA events): NN-only kernel vs the full fused round."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from vtkcloudpoint_b200 import Context, synth

ctx = Context(0)
dev = torch.device("cuda", 0)
model, data, _, _ = synth.icp_clouds(0xC3, 1_000_000, 100_000)
dm, dd = torch.from_numpy(model).to(dev), torch.from_numpy(data).to(dev)
ctx.icp_set_model_dev(dm)
state, order = ctx.icp_rigid_dev(dd, -1.0, 50)
torch.cuda.synchronize()
# data moved by the converged transform: the steady-state query set
st = state.cpu().numpy()
R = torch.tensor(st[:9].reshape(3, 3), device=dev); T = torch.tensor(st[9:12], device=dev)
moved = (R @ dd + T[:, None]).contiguous()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, q in (("NN only, raw data (round 1 geometry)", dd), ("NN only, converged geometry", moved)):
    for _ in range(3):
        ctx.closest_point_set_dev(q)
    e0.record()
    for _ in range(20):
        ctx.closest_point_set_dev(q, want_sqdist=False)
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per launch")
for iters in (1, 10, 50, 200):
    out = ctx.icp_rigid_dev(dd, -1.0, iters)
    torch.cuda.synchronize()
    e0.record(); ctx.icp_rigid_dev(dd, -1.0, iters, out=out); e1.record(); torch.cuda.synchronize()
    print(f"icp_rigid_dev {iters} rounds: {e0.elapsed_time(e1) * 1e3:.1f} us total, {e0.elapsed_time(e1) / iters * 1e3:.1f} us per round")
ctx.close()
# GPU-side duration of the NN-only kernel (CUDA events around each launch, via the library's profile hook)
ctx2 = Context(0)
ctx2.icp_set_model_dev(dm)
for name, q in (("raw", dd), ("converged", moved)):
    ctx2.closest_point_set_dev(q)
    ctx2.profile(True)
    for _ in range(10):
        ctx2.closest_point_set_dev(q, want_sqdist=False)
    rep = ctx2.profile_report()
    ctx2.profile(False)
    ts = sorted(ms for _, ms in rep)
    print(f"k_icp_closest ({name}): median {ts[len(ts)//2]*1e3:.1f} us, min {ts[0]*1e3:.1f} us on the GPU")
ctx2.close()
