#!/bin/bash
# round 2, call D (2 GPUs): the whole GPU suite incl. group mode, pageable staging and real multi-GPU runs; then benches
out=gpurun_out; mkdir -p $out
timeout 1500 python -m pytest tests/test_group_gpu.py tests/test_peer_lockstep_gpu.py -x -q > $out/r2d_group.log 2>&1; echo "group rc=$?"; tail -12 $out/r2d_group.log
timeout 2400 python -m pytest tests -m gpu -q --deselect tests/test_group_gpu.py --deselect tests/test_peer_lockstep_gpu.py > $out/r2d_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 $out/r2d_pytest.log
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --no-cpu > $out/r2d_bench1.json 2> $out/r2d_bench1.err; echo "bench1 rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 > $out/r2d_bench2.json 2> $out/r2d_bench2.err; echo "bench2 rc=$?"
python - <<'PY'
import json
for n in (1,2):
    try:
        d=json.load(open(f'gpurun_out/r2d_bench{n}.json'))
        print(f'N={n} ms/step',d['ms_per_step'],'value',d['value'],'e2e',d['e2e'])
        if 'parity' in d: print({k:v for k,v in d['parity'].items() if not k.endswith('how')})
        sec=d['secondary']
        print({k:(v['ms_per_iter'] if isinstance(v,dict) else v) for k,v in sec.items() if k in('target_sharded_weak','source_sharded_strong','ms_per_iter','value')})
        print(d['kernel_ms_per_step'])
    except Exception as e: print(n,'ERR',e)
PY
tail -3 $out/r2d_bench2.err
