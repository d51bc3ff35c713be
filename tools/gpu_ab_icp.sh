#!/bin/bash
# A/B of build variants on the N=1 bench's ICP leg
out=gpurun_out; mkdir -p $out
for f in vtkcloudpoint_b200/ab/libvpc_*.so; do
  n=$(basename $f .so); n=${n#libvpc_}
  VPC_LIB=$PWD/$f timeout 300 python bench.py --no-cpu --no-blocked --steps 10 > $out/abi_$n.json 2> $out/abi_$n.err || echo "$n failed"
  python - "$n" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/abi_{n}.json').read().strip().splitlines()[-1])
    s=d['secondary']
    print(f"{n:10s} icp us/iter {s['ms_per_iter']*1e3:.2f}  iters/s {s['value']:.0f}  e2e {s['e2e_iters_per_s']:.0f}  {s['kernel_ms_per_launch']}")
except Exception as e:
    print(n,'ERR',e)
PY
done
