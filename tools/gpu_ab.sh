#!/bin/bash
# A/B of kernel build variants on the N=1 bench (DBSCAN only)
out=gpurun_out; mkdir -p $out
for f in vtkcloudpoint_b200/ab/libvpc_*.so; do
  n=$(basename $f .so); n=${n#libvpc_}
  VPC_LIB=$PWD/$f timeout 300 python bench.py --no-cpu --no-icp --steps 30 > $out/ab_$n.json 2> $out/ab_$n.err || echo "$n failed"
  python - "$n" <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.load(open(f'gpurun_out/ab_{n}.json'))
    k=d['kernel_ms_per_step']
    print(f"{n:10s} ms/step {d['ms_per_step']:.4f}  " + ' '.join(f"{a.replace('k_db_','').replace('k_scan_exclusive','scan')}={b*1e3:.0f}" for a,b in k.items()))
except Exception as e:
    print(n,'ERR',e)
PY
done
