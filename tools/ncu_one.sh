#!/bin/bash
# usage: tools/ncu_one.sh <tag> <quick_bench/bench args...>   -> gpurun_out/prof_<tag>.ncu-rep (one k_march launch, full set)
tag=$1; shift
ncu --set full --clock-control none --import-source on -k regex:k_march -s 3 -c 1 -f -o gpurun_out/prof_$tag "$@" > /dev/null 2>&1
