"""Newton vmult throughput over the mesh sizes of the survey's synthetic sweep (one GPU): runs bench.py --kernel-only
per size in a fresh process and prints one line per size."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sizes = [int(a) for a in sys.argv[1:]] or [32, 64, 96, 128, 160, 192]
for n in sizes:
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--n", str(n), "--steps", "30", "--warmup", "5", "--kernel-only"],
                       capture_output=True, text=True, timeout=900)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    if not lines:
        print(n, "failed", r.stderr[-300:])
        continue
    d = json.loads(lines[-1])
    dofs = 3 * (2 * n + 1) ** 3 + (n + 1) ** 3
    print(f"n={n:4d} {dofs/1e6:8.2f} MDoF | burst {d['burst_ms_per_step']:.4f} ms/step (kernel {d['burst_kernel_ms']:.4f}) {dofs/d['burst_ms_per_step']/1e6:6.2f} GDoF/s"
          f" | sustained {d['ms_per_step']:.4f} ms/step (kernel {d['kernel_ms']:.4f}) {dofs/d['ms_per_step']/1e6:6.2f} GDoF/s", flush=True)
