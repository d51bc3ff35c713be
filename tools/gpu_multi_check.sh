#!/bin/bash
N=${1:-4}
out=gpurun_out; mkdir -p $out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/dist_check_peer.py 1000000 1000000 100000 10 > $out/multi_peer$N.log 2>&1; echo "dist_check_peer rc=$?"
grep "rank 0" $out/multi_peer$N.log | tail -14; grep -c OK $out/multi_peer$N.log; grep MISMATCH $out/multi_peer$N.log | head -3
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --c5 > $out/multi_bench$N.json 2> $out/multi_bench$N.err; echo "bench rc=$?"
python - $N <<'PY'
import json,sys
N=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/multi_bench{N}.json').read().strip().splitlines()[-1])
    print(f'N={N} ms/step',d['ms_per_step'],'value',d['value'])
    print('e2e',d['e2e'])
    print({k:v for k,v in d['parity'].items() if not k.endswith('how')})
    sec=d['secondary']
    print({k:(v['ms_per_iter'] if isinstance(v,dict) else v) for k,v in sec.items() if k in('target_sharded_weak','source_sharded_strong')})
    print(d['kernel_ms_per_step']); print('c4',d.get('secondary_c4')); print('c5',d.get('secondary_c5')); print(d['nvlink'])
except Exception as e: print('ERR',e)
PY
grep -v "^W1018\|^\*\*\*\|OMP_NUM\|^$" $out/multi_bench$N.err | tail -8
