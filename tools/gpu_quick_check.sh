#!/bin/bash
# short validation: parity suites without the full-size cases + a DBSCAN-only bench line
out=gpurun_out; mkdir -p $out
timeout 200 python -m pytest tests/test_dbscan_gpu.py tests/test_golden_gpu.py tests/test_peer_lockstep_gpu.py tests/test_group_gpu.py tests/test_blocked_gpu.py tests/test_host_mirror_gpu.py -x -q 2>&1 | tail -2
timeout 100 python bench.py --no-blocked --no-cpu --no-icp 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['gpu_launches'], d['kernel_ms_per_step'])"
