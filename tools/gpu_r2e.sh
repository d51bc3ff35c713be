#!/bin/bash
# round 2, call E (1 GPU): group mode (emulated ranks), pageable staging, whole suite, N=1 bench
out=gpurun_out; mkdir -p $out
timeout 1500 python -m pytest tests/test_group_gpu.py tests/test_peer_lockstep_gpu.py -x -q > $out/r2e_group.log 2>&1; echo "group rc=$?"; tail -12 $out/r2e_group.log
timeout 2400 python -m pytest tests -m gpu -q --deselect tests/test_group_gpu.py --deselect tests/test_peer_lockstep_gpu.py > $out/r2e_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 $out/r2e_pytest.log
timeout 600 python bench.py --no-cpu > $out/r2e_bench1.json 2> $out/r2e_bench1.err; echo "bench1 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2e_bench1.json'))
print('N=1 ms/step',d['ms_per_step'],'value',d['value'],'e2e',d['e2e'])
PY
for t in 0 2 4 8; do VPC_COPY_THREADS=$t timeout 300 python bench.py --no-cpu --no-icp --steps 10 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('copy threads $t: pageable e2e', round(d['e2e']['value'],1), 'Mpts/s', round(d['e2e']['ms_per_step'],3),'ms; pinned', round(d['e2e']['page_locked']['value'],1))"; done
nproc
