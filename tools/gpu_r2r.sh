#!/bin/bash
# PDL A/B on one GPU: bench twice (VPC_PDL=1 / 0), parity suite with PDL on
out=gpurun_out; mkdir -p $out
for v in 1 0; do
VPC_PDL=$v timeout 300 python bench.py --no-c4 --no-blocked > $out/r2r_bench_pdl$v.json 2> $out/r2r_bench_pdl$v.err; echo "bench pdl=$v rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r2r_bench_pdl$v.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['secondary']['icp_iters_per_s'] if 'secondary' in d and 'icp_iters_per_s' in d['secondary'] else '')"
done
timeout 1500 python -m pytest tests -m gpu -x -q > $out/r2r_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $out/r2r_pytest.log
