#!/bin/bash
# N = 1 bench line + the ncu launch list of one DBSCAN step (short form of gpu_round.sh)
tag=${1:-rX}
out=gpurun_out; mkdir -p $out
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
python tools/profile_step.py dbscan > $out/${tag}_plain_dbscan.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $out/${tag}_launches_dbscan.csv \
    python tools/profile_step.py dbscan > $out/${tag}_ncu_l_dbscan.log 2>&1; echo "ncu rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/${tag}_bench.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['page_locked']['value'], d['roofline'], d['secondary']['value'], d['kernel_ms_per_step'])"
