#!/bin/bash
# Runs HERE (no GPU): digest gpurun_out/*.ncu-rep of a round into committed text summaries under profiles/.
tag=${1:-r01}
mkdir -p profiles
for m in M2 M1 M0; do
  [ -f gpurun_out/prof_${tag}_$m.ncu-rep ] && python tools/ncu_summary.py gpurun_out/prof_${tag}_$m.ncu-rep 132710400 > profiles/${tag}_ncu_g2_$m.txt
done
[ -f gpurun_out/launches_$tag.csv ] && cp gpurun_out/launches_$tag.csv profiles/${tag}_launches.csv
[ -f gpurun_out/bench_$tag.json ] && cp gpurun_out/bench_$tag.json profiles/${tag}_bench.json
[ -f gpurun_out/bench_ref_$tag.json ] && cp gpurun_out/bench_ref_$tag.json profiles/${tag}_bench_reference.json
ls -la profiles
