#!/bin/bash
# 2 GPUs: the whole GPU suite (incl. tests/test_multi_gpu.py) + bench
out=gpurun_out; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q > $out/r2u_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $out/r2u_pytest.log
bash tools/gpu_r2t.sh 2
