#!/bin/bash
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_peer_lockstep_gpu.py tests/test_group_gpu.py -x -q > $out/r2q_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $out/r2q_pytest.log
timeout 300 python tools/lockstep_profile.py 4 > $out/r2q_lockstep.log 2>&1; echo "lockstep rc=$?"; tail -12 $out/r2q_lockstep.log | head -3
