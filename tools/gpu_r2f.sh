#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 600 python -m pytest tests/test_group_gpu.py tests/test_dbscan_gpu.py tests/test_icp_gpu.py -x -q > $out/r2f_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $out/r2f_pytest.log
for t in 0 1 2 4 6 8; do VPC_COPY_THREADS=$t timeout 300 python bench.py --no-cpu --no-icp --steps 20 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('copy threads $t: pageable e2e', round(d['e2e']['value'],1), 'Mpts/s', round(d['e2e']['ms_per_step'],3),'ms; pinned', round(d['e2e']['page_locked']['value'],1))"; done
