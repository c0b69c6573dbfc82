#!/bin/bash
# usage: tools/ab_env.sh VAR v1 v2 ...  -- bench.py --kernel-only with VAR set to each value in turn (developer A/B of a run-time switch)
var=$1; shift
for v in "$@"; do
  env $var=$v timeout 200 python bench.py --kernel-only --steps 30 --warmup 10 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$var=$v', 'burst kernel/step', round(d['burst_kernel_ms'],4), round(d['burst_ms_per_step'],4), 'sustained kernel/step', round(d['kernel_ms'],4), round(d['ms_per_step'],4), d['clocks']['sm_mhz'], d['checksum'])"
done
