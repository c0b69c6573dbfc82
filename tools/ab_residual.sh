#!/bin/bash
# usage: tools/ab_residual.sh lib1 lib2 ...  (names under adaflo_b200/lib/variants, or "main"): residual (K2) time of the bench workload
for v in "$@"; do
  if [ "$v" = main ]; then unset ADAFLO_B200_LIB; else export ADAFLO_B200_LIB=$PWD/adaflo_b200/lib/variants/lib$v.so; fi
  timeout 200 python bench.py --no-cpu-baseline --no-level-set --no-nonlinear-step --steps 5 --warmup 3 --e2e-steps 1 --preroll 0.2 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', 'residual ms', round(d['residual']['ms'],4), 'vmult kernel', round(d['roofline']['kernel_ms'],4))"
done
