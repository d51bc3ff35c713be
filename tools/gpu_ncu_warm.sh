#!/bin/bash
out=gpurun_out; mkdir -p $out
python tools/profile_step.py dbscan > $out/r2o_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --cache-control none --import-source on --profile-from-start off -k regex:'k_db_count|k_db_union' -f -o $out/r02_hot_warm python tools/profile_step.py dbscan > $out/r2o_ncu.log 2>&1; echo "ncu rc=$?"
ls -la $out/r02_hot_warm.ncu-rep
