#!/bin/bash
# last check of the round on one GPU: whole GPU suite, smoke(), bench line
out=gpurun_out; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q > $out/final_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/final_pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $out/final_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $out/final_smoke.log
timeout 600 python bench.py > $out/final_bench.json 2> $out/final_bench.err; echo "bench rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/final_bench.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['page_locked']['value'], d['roofline']['frac'], d['secondary']['value'])"
