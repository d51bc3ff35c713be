#!/bin/bash
# bench only at N GPUs (parity legs inside), JSON line kept
N=${1:-8}
out=gpurun_out; mkdir -p $out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --c5 > $out/benchn_bench$N.json 2> $out/benchn_bench$N.err; echo "bench rc=$?"
python - $N <<'PY'
import json,sys
N=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/benchn_bench{N}.json').read().strip().splitlines()[-1])
    print(f'N={N} ms/step',d['ms_per_step'],'value',d['value'])
    print('e2e',d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e'].get('page_locked'))
    print({k:v for k,v in d['parity'].items() if not k.endswith('how')})
    sec=d['secondary']
    print({k:(v['ms_per_iter'] if isinstance(v,dict) else v) for k,v in sec.items() if k in('target_sharded_weak','source_sharded_strong')})
    print(d['kernel_ms_per_step']); print('c4',d.get('secondary_c4',{}).get('ms_per_step')); print('c5',d.get('secondary_c5',{}).get('ms_per_pass')); print(d['nvlink'])
except Exception as e: print('ERR',e)
PY
grep -v "^W1018\|^W1019\|^\*\*\*\|OMP_NUM\|^$" $out/benchn_bench$N.err | tail -8
