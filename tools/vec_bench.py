"""Bandwidth of the Krylov-step vector kernels at the bench size (one GPU): python tools/vec_bench.py
ADAFLO_B200_DOT_V2=0 selects the 8-byte multi-dot."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adaflo_b200.operators import Context  # noqa: E402

ctx = Context(0)
n = 3 * 257 ** 3
g = torch.Generator(device=ctx.device).manual_seed(1)
x = torch.rand(n, dtype=torch.float64, device=ctx.device, generator=g)
ys = [torch.rand(n, dtype=torch.float64, device=ctx.device, generator=g) for _ in range(8)]
out = {"n": n, "dot_v2": os.environ.get("ADAFLO_B200_DOT_V2", "1")}
ref = [float(torch.dot(x, y)) for y in ys]
for k in (1, 4, 8):
    got = ctx.multi_dot(x, ys[:k])
    assert all(abs(a - b) <= 1e-12 * abs(b) for a, b in zip(got, ref)), (got, ref[:k])
    for _ in range(3):
        ctx.multi_dot(x, ys[:k])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        ctx.multi_dot(x, ys[:k])
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 20
    out[f"multi_dot_{k}"] = {"ms": dt * 1e3, "gbs": (k + 1) * n * 8 / dt / 1e9}
print(json.dumps(out))
