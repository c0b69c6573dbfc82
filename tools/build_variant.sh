#!/bin/bash
# usage: tools/build_variant.sh NAME "-DFLAG1 -DFLAG2"  -> adaflo_b200/lib/variants/libNAME.so (cartesian kernels rebuilt with the flags)
set -e
cd "$(dirname "$0")/../adaflo_b200/csrc"
mkdir -p build/var_$1 ../lib/variants
ARCH="-gencode arch=compute_100a,code=sm_100a"
nvcc $ARCH -lineinfo -O3 -std=c++17 -Xcompiler -fPIC $2 -c ns_vmult_cartesian.cu -o build/var_$1/ns_vmult_cartesian.o
OBJS=$(ls build/*.o | grep -v ns_vmult_cartesian.o)
nvcc $ARCH -shared -o ../lib/variants/lib$1.so $OBJS build/var_$1/ns_vmult_cartesian.o -ldl
echo built lib$1.so
