#!/bin/bash
# 1 GPU: full parity suite + emulated-rank kernel times after the slab-step fusions
out=gpurun_out; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q > $out/r2p_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $out/r2p_pytest.log
timeout 300 python tools/lockstep_profile.py 4 > $out/r2p_lockstep.log 2>&1; echo "lockstep rc=$?"; tail -12 $out/r2p_lockstep.log
timeout 300 python bench.py --no-c4 --no-blocked > $out/r2p_bench1.json 2> $out/r2p_bench1.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2p_bench1.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['roofline'], d['kernel_ms_per_step'])"
