#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_blocked_gpu.py tests/test_host_mirror_gpu.py tests/test_peer_lockstep_gpu.py tests/test_group_gpu.py -x -q > $out/r2h_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 $out/r2h_pytest.log
