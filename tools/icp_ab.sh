#!/bin/bash
# A/B of the ICP round on the GPU box: rebuilds libvpc.so with different -D flags and times the round (tools/time_icp_parts.py).
for flags in "-DVPC_ICP_ITER_BLOCK=256" "-DVPC_ICP_ITER_BLOCK=128" "-DVPC_ICP_ITER_BLOCK=64"; do
  echo "== $flags"
  VPC_NVCC_EXTRA="$flags" python -c "from vtkcloudpoint_b200 import _build; _build.build_lib(force=True)" || exit 1
  python tools/time_icp_parts.py 2>&1 | grep -E "icp_rigid_dev (50|200)|NN only, conv"
done
python -c "from vtkcloudpoint_b200 import _build; _build.build_lib(force=True)"
python - <<'PY'
# the sharded steps on one GPU: per-kernel times of NN / accumulate / solve
import sys; sys.path.insert(0, '.')
import torch
from vtkcloudpoint_b200 import Context, synth
from vtkcloudpoint_b200.distributed import GpuIcpBackend, icp_rigid_sharded
ctx = Context(0); dev = torch.device('cuda', 0)
model, data, _, _ = synth.icp_clouds(0xC3, 1_000_000, 100_000)
dm, dd = torch.from_numpy(model).to(dev), torch.from_numpy(data).to(dev)
be = GpuIcpBackend(ctx)
icp_rigid_sharded(be, dm, 0, dd, -1.0, 10)
ctx.profile(True)
icp_rigid_sharded(be, dm, 0, dd, -1.0, 20)
rep = ctx.profile_report(); ctx.profile(False)
agg = {}
for k, v in rep: agg.setdefault(k, []).append(v)
print({k: round(sorted(v)[len(v)//2] * 1e3, 1) for k, v in agg.items()}, "us (median per launch)")
PY
