#!/bin/bash
# usage: tools/ab.sh OUT.jsonl lib1 lib2 ...   (names under adaflo_b200/lib/variants, or "main")
# Kernel-level A/B on the bench workload: bench.py --kernel-only with each library, one JSON line per run.
out=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = main ]; then unset ADAFLO_B200_LIB; else export ADAFLO_B200_LIB=$PWD/adaflo_b200/lib/variants/lib$v.so; fi
  timeout 600 python bench.py --kernel-only --steps 30 --warmup 10 2>>gpurun_out/ab.err | tail -n 1 | sed "s/^{/{\"variant\": \"$v\", /" >> gpurun_out/$out
done
unset ADAFLO_B200_LIB
cat gpurun_out/$out
