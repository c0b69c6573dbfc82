#!/bin/bash
# usage (on the GPU box, from the repo root): tools/final_gpu_run.sh TAG
# Full GPU validation of a round: the -m gpu suite, the default bench line, the reference arm, the ncu launch list of
# `bench.py --steps 3 --warmup 3` and one `ncu --set full` capture of the dominant kernel.  Everything goes to gpurun_out/.
tag=${1:-final}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/pytest_${tag}.log
tail -3 gpurun_out/pytest_${tag}.log
timeout 600 python bench.py 2>gpurun_out/bench_${tag}.err | tail -1 > gpurun_out/bench_${tag}.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>>gpurun_out/bench_${tag}.err | tail -1 > gpurun_out/bench_${tag}_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${tag}.csv \
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_bench_${tag}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ns_vmult_r1_kernel -s 8 -c 1 -f -o gpurun_out/prof_vmult_${tag} \
  python bench.py --kernel-only --steps 5 --warmup 5 > gpurun_out/ncu_full_${tag}.log 2>&1
ls -la gpurun_out/prof_vmult_${tag}.ncu-rep
python - <<P
import json
d = json.loads(open("gpurun_out/bench_${tag}.json").read())
print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches", "clocks")}, d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["e2e"]["ms_per_step"], d["burst"]["roofline_frac"], d["residual"]["ms"], d["nonlinear_step"]["ms"])
print({k: round(v["ms"], 3) for k, v in d["level_set"]["operators"].items()})
print(open("gpurun_out/bench_${tag}_ref.json").read()[:600])
P
