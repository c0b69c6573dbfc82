"""where does a GMRES iteration spend its time (developer tool); run with torchrun for N > 1"""
import os, sys, time, math
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adaflo_b200 import capi
from adaflo_b200.mesh import BoxMesh, beltrami_velocity, seeded_uniform
from adaflo_b200.operators import BlockVector, Context, MatrixFree, NavierStokesMatrix

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = Context(local)
if world > 1:
    def bcast(b):
        obj = [b]; dist.broadcast_object_list(obj, src=0); return obj[0]
    ctx.init_comm(rank, world, bcast)
n = 128
mesh = BoxMesh((n, n, n * world), upper=(1.0, 1.0, -1.0 + 2.0 * world), rank=rank, nranks=world)
mf = MatrixFree(ctx, mesh, "cartesian")
prm = capi.FlowParams(capi.INCOMPRESSIBLE, capi.NEWTON, 0, 0, 1.0, 1.0, 0.0, 0.0, 0.0, 0.0, 0.5)
tw = capi.TimeWeights(30.0, -40.0, 10.0, 1.0, 2.0, -1.0)
sol, old, old_old = BlockVector.like(mf), BlockVector.like(mf), BlockVector.like(mf)
A = NavierStokesMatrix(prm, 0, 1, 0, 1, old, old_old)
A.initialize(mf, tw, True)
sol.block(0).copy_(torch.from_numpy(beltrami_velocity(mesh.u.coords, 0.0).reshape(-1)).to(ctx.device))
rhs, zero = BlockVector.like(mf), BlockVector.like(mf)
A.residual(rhs, sol, zero)
nu, npp = mf.n_owned(0), mf.n_owned(1)
V = [[torch.rand_like(sol.block(0)), torch.rand_like(sol.block(1))] for _ in range(16)]
w = [torch.rand_like(sol.block(0)), torch.rand_like(sol.block(1))]
def sync():
    torch.cuda.synchronize()
    if world > 1: dist.barrier(); torch.cuda.synchronize()
def timeit(name, f, reps=20):
    for _ in range(3): f()
    sync(); t0 = time.perf_counter()
    for _ in range(reps): f()
    sync(); t = (time.perf_counter() - t0) / reps
    if rank == 0: print(f"N={world} {name}: {t*1e3:.3f} ms", flush=True)
timeit("vmult", lambda: A.vmult(BlockVector(*w), BlockVector(*V[0])))
timeit("residual", lambda: A.residual(rhs, sol, zero), 5)
timeit("multi_dot u k=17", lambda: ctx.multi_dot(w[0], [v[0] for v in V] + [w[0]], nu))
timeit("multi_dot p k=17", lambda: ctx.multi_dot(w[1], [v[1] for v in V] + [w[1]], npp))
timeit("multi_axpy u k=16", lambda: ctx.multi_axpy(w[0], [1e-12] * 16, [v[0] for v in V]))
timeit("multi_axpy p k=16", lambda: ctx.multi_axpy(w[1], [1e-12] * 16, [v[1] for v in V]))
timeit("dot u", lambda: ctx.multi_dot(w[0], [w[0]], nu))
timeit("scale-copy u", lambda: w[0] * 0.5)
