#!/bin/bash
# round 2, call C (2 GPUs): 1-GPU parity + bench, then the 2-GPU peer check + bench
out=gpurun_out; mkdir -p $out
CUDA_VISIBLE_DEVICES=0 timeout 1200 python -m pytest tests/test_dbscan_gpu.py tests/test_golden_gpu.py tests/test_peer_lockstep_gpu.py tests/test_blocked_gpu.py -x -q > $out/r2c_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $out/r2c_pytest.log
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --no-cpu > $out/r2c_bench1.json 2> $out/r2c_bench1.err; echo "bench1 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c_bench1.json'))
print('N=1 ms/step',d['ms_per_step'],'value',d['value'],'e2e',d['e2e']['value'],'icp',d['secondary']['value'])
print(d['kernel_ms_per_step'])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 > $out/r2c_bench2.json 2> $out/r2c_bench2.err; echo "bench2 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2c_bench2.json'))
print('N=2 ms/step',d['ms_per_step'],'value',d['value'],'e2e',d['e2e']['value'],'parity',{k:v for k,v in d['parity'].items() if not k.endswith('how')})
print({k:(v['ms_per_iter'] if isinstance(v,dict) else v) for k,v in d['secondary'].items() if k.endswith('weak') or k.endswith('strong')})
print(d['kernel_ms_per_step'])
PY
tail -3 $out/r2c_bench2.err
