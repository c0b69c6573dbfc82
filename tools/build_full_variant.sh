#!/bin/bash
# usage: tools/build_full_variant.sh NAME "-DFLAG1 -DFLAG2"  -> adaflo_b200/lib/variants/libNAME.so
# (the whole library rebuilt with the flags: layout macros such as AB_R1_ROT touch every translation unit)
set -e
cd "$(dirname "$0")/../adaflo_b200/csrc"
mkdir -p build/full_$1 ../lib/variants
ARCH="-gencode arch=compute_100a,code=sm_100a"
pids=()
for f in api ls_api ls_ops ls_force comm vector_ops pressure_ops ns_residual ns_residual_r1 ns_vmult_cartesian ns_vmult_affine ns_vmult_general; do
  nvcc $ARCH -lineinfo -O3 -std=c++17 -Xcompiler -fPIC $2 -c $f.cu -o build/full_$1/$f.o &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
nvcc $ARCH -shared -o ../lib/variants/lib$1.so build/full_$1/*.o -ldl
echo built lib$1.so
