#!/bin/bash
N=${1:-4}
out=gpurun_out; mkdir -p $out
CUDA_VISIBLE_DEVICES=0 timeout 600 python -m pytest tests/test_peer_lockstep_gpu.py -x -q > $out/r2k_lockstep.log 2>&1; echo "lockstep rc=$?"; tail -3 $out/r2k_lockstep.log
timeout 600 python -m pytest tests/test_group_gpu.py -x -q > $out/r2k_group.log 2>&1; echo "group rc=$?"; tail -3 $out/r2k_group.log
bash tools/gpu_r2j.sh $N
