#!/bin/bash
N=${1:-4}
out=gpurun_out; mkdir -p $out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --no-c4 --no-icp > $out/r2n_bench$N.json 2> $out/r2n_bench$N.err; echo "bench rc=$?"
python - $N <<'PY'
import json,sys
N=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/r2n_bench{N}.json').read().strip().splitlines()[-1])
    print(f'N={N} ms/step',d['ms_per_step'],'value',d['value'], 'e2e', d['e2e']['value'], 'pinned', d['e2e']['page_locked']['value'])
    print({k:v for k,v in d['parity'].items() if not k.endswith('how')})
    print(d['kernel_ms_per_step'])
except Exception as e: print('ERR',e)
PY
grep -v "^W1018\|^\*\*\*\|OMP_NUM\|^$" $out/r2n_bench$N.err | tail -5
