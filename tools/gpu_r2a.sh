#!/bin/bash
# round 2, call A (1 GPU): the new peer-memory kernels in lockstep emulation, then the whole GPU suite
out=gpurun_out; mkdir -p $out
timeout 900 python -m pytest tests/test_peer_lockstep_gpu.py -x -q > $out/r2a_lockstep.log 2>&1; echo "lockstep rc=$?"
tail -15 $out/r2a_lockstep.log
timeout 2400 python -m pytest tests -m gpu -q --deselect tests/test_peer_lockstep_gpu.py > $out/r2a_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 $out/r2a_pytest.log
nvidia-smi --query-gpu=name,memory.used --format=csv
