#!/bin/bash
# usage: build_variant.sh <name> <extra nvcc -D flags...>   -> ../libcvsteer_b200_<name>.so (tuning A/B builds, not shipped)
set -e
name=$1; shift
mkdir -p build/var_$name
FL="-std=c++20 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo --expt-relaxed-constexpr -Xcompiler -fPIC,-fvisibility=hidden"
nvcc $FL "$@" -c march_g2.cu -o build/var_$name/march_g2.o &
nvcc $FL "$@" -c march_g4.cu -o build/var_$name/march_g4.o &
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -cudart static -o ../libcvsteer_b200_$name.so build/capi.o build/kernels.o build/taps.o build/var_$name/march_g2.o build/var_$name/march_g4.o build/march_g2_lines.o
echo built $name
