#!/bin/bash
# usage: tools/build_variant.sh NAME "-DFLAG1 -DFLAG2" [BASE]  -> adaflo_b200/lib/variants/libNAME.so
# Rebuilds only the Cartesian cell kernels with the extra macros; all other objects come from build/ (the main library)
# or from build/full_BASE (a tools/build_full_variant.sh build, for macros that change the data layouts).
set -e
cd "$(dirname "$0")/../adaflo_b200/csrc"
mkdir -p build/var_$1 ../lib/variants
BASE=build
[ -n "$3" ] && BASE=build/full_$3
ARCH="-gencode arch=compute_100a,code=sm_100a"
nvcc $ARCH -lineinfo -O3 -std=c++17 -Xcompiler -fPIC $2 -c ns_vmult_cartesian.cu -o build/var_$1/ns_vmult_cartesian.o
OBJS=$(ls $BASE/*.o | grep -v ns_vmult_cartesian.o)
nvcc $ARCH -shared -o ../lib/variants/lib$1.so $OBJS build/var_$1/ns_vmult_cartesian.o -ldl
echo built lib$1.so
