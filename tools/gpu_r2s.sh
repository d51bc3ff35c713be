#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_peer_lockstep_gpu.py tests/test_group_gpu.py tests/test_icp_gpu.py -x -q > $out/r2s_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $out/r2s_pytest.log
timeout 300 python tools/lockstep_profile.py 4 > $out/r2s_lockstep.log 2>&1; echo "lockstep rc=$?"; tail -12 $out/r2s_lockstep.log | head -3
for v in 1 0; do
VPC_PDL=$v timeout 300 python bench.py --no-c4 --no-blocked > $out/r2s_bench_pdl$v.json 2> $out/r2s_bench_pdl$v.err; echo "bench pdl=$v rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2s_bench_pdl$v.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['secondary'])"
done
