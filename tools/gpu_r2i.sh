#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_group_gpu.py tests/test_multi_gpu.py -x -q > $out/r2i_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $out/r2i_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 > $out/r2i_bench2.json 2> $out/r2i_bench2.err; echo "bench2 rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2i_bench2.json').read().strip().splitlines()[-1])
    print('N=2 ms/step',d['ms_per_step'],'value',d['value'])
    print('e2e',d['e2e'])
    print({k:v for k,v in d['parity'].items() if not k.endswith('how')})
    sec=d['secondary']
    print({k:(v['ms_per_iter'] if isinstance(v,dict) else v) for k,v in sec.items() if k in('target_sharded_weak','source_sharded_strong')})
    print(d['kernel_ms_per_step']); print('c4',d.get('secondary_c4'))
except Exception as e: print('ERR',e)
PY
grep -v "^W1018\|^\*\*\*\|OMP_NUM\|^$" $out/r2i_bench2.err | tail -8
