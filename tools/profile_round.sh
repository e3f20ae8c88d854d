#!/bin/bash
# Runs ON THE GPU BOX (under gpurun): contract bench, ncu launch list of the same command, and one full-set capture of
# the dominant kernel per mode.  Results land in gpurun_out/; tools/profile_digest.sh turns them into profiles/*.txt here.
tag=${1:-r01}
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_$tag.json 2>> gpurun_out/bench_$tag.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 5 --warmup 3 --e2e-steps 1 --no-cpu > /dev/null 2>&1
for m in M2 M1 M0; do
  ncu --set full --clock-control none --import-source on -k regex:k_march -s 3 -c 1 -f -o gpurun_out/prof_${tag}_$m \
      python bench.py --mode $m --steps 3 --warmup 3 --no-e2e --no-cpu --quick > /dev/null 2>&1
done
