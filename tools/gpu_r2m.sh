#!/bin/bash
out=gpurun_out; mkdir -p $out
CUDA_VISIBLE_DEVICES=0 timeout 600 python -m pytest tests/test_peer_lockstep_gpu.py tests/test_group_gpu.py -x -q > $out/r2m_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/r2m_pytest.log
CUDA_VISIBLE_DEVICES=0 timeout 600 python tools/lockstep_profile.py 4 2>&1 | head -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --no-c4 > $out/r2m_bench2.json 2> $out/r2m_bench2.err; echo "bench2 rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r2m_bench2.json').read().strip().splitlines()[-1])
    print('N=2 ms/step',d['ms_per_step'],'value',d['value'])
    print('e2e',d['e2e']['value'], d['e2e']['ms_per_step'])
    print({k:v for k,v in d['parity'].items() if not k.endswith('how')})
    print(d['kernel_ms_per_step']); print(d['roofline'])
except Exception as e: print('ERR',e)
PY
grep -v "^W1018\|^\*\*\*\|OMP_NUM\|^$" $out/r2m_bench2.err | tail -5
