#!/bin/bash
# usage: build_variant.sh <name> <extra nvcc -D flags...>   -> ../variants/libcvsteer_b200_<name>.so (tuning A/B builds, not shipped)
# Recompiles only the translation units named in $UNITS (default: the two G4 ones) and links the rest from build/.
set -e
name=$1; shift
UNITS=${UNITS:-"march_g4 march_g4_steer"}
mkdir -p build/var_$name ../variants
FL="-std=c++20 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo --expt-relaxed-constexpr -Xcompiler -fPIC,-fvisibility=hidden"
objs=""
for o in capi bands kernels march_g2 march_g2_lines march_g2_steer march_g2_steer2 march_g4_steer march_g4 taps; do
  if [[ " $UNITS " == *" $o "* ]]; then
    nvcc $FL "$@" -c $o.cu -o build/var_$name/$o.o &
    objs="$objs build/var_$name/$o.o"
  else
    objs="$objs build/$o.o"
  fi
done
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a -cudart static -o ../variants/libcvsteer_b200_$name.so $objs -ldl
echo built $name
