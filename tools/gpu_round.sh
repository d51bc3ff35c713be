#!/bin/bash
# One GPU-box pass: parity tests, size sweep, bench, ncu launch list + full capture (each only after the plain run exits 0).
# Usage (from the repo root, through gpurun): bash tools/gpu_round.sh <tag>
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/${tag}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $out/${tag}_pytest.log
python tools/scale_single.py 1000000 10000000 100000000 > $out/${tag}_scale.log 2>&1; echo "scale rc=$?"
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_ref.json 2> $out/${tag}_bench_ref.err; echo "bench ref rc=$?"
for what in dbscan icp; do
  python tools/profile_step.py $what > $out/${tag}_plain_$what.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $out/${tag}_launches_$what.csv \
      python tools/profile_step.py $what > $out/${tag}_ncu_l_$what.log 2>&1 && \
  ncu --set full --clock-control none --import-source on --profile-from-start off -f -o $out/${tag}_$what \
      python tools/profile_step.py $what > $out/${tag}_ncu_f_$what.log 2>&1
  echo "ncu $what rc=$?"
done
tail -3 $out/${tag}_pytest.log; cat $out/${tag}_scale.log; cat $out/${tag}_bench.json
