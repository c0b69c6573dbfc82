"""micro-timing of the multi-GPU reductions / halo exchange (developer tool)"""
import os, sys, time
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adaflo_b200 import capi
from adaflo_b200.mesh import BoxMesh
from adaflo_b200.operators import BlockVector, Context, MatrixFree, NavierStokesMatrix

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = Context(local)
def bcast(b):
    obj = [b]; dist.broadcast_object_list(obj, src=0); return obj[0]
ctx.init_comm(rank, world, bcast)
n = 64
mesh = BoxMesh((n, n, n * world), upper=(1.0, 1.0, -1.0 + 2.0 * world), rank=rank, nranks=world)
mf = MatrixFree(ctx, mesh, "cartesian")
x = mf.initialize_dof_vector(0); x.fill_(1.0)
ys = [mf.initialize_dof_vector(0) for _ in range(9)]
for y in ys: y.fill_(0.5)
nown = mf.n_owned(0)
def timeit(name, f, reps=100):
    for _ in range(5): f()
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); t = (time.perf_counter() - t0) / reps
    if rank == 0: print(f"{name}: {t*1e6:.1f} us", flush=True)
timeit("multi_dot k=1", lambda: ctx.multi_dot(x, ys[:1], nown))
timeit("multi_dot k=9", lambda: ctx.multi_dot(x, ys, nown))
timeit("multi_axpy k=8", lambda: ctx.multi_axpy(x, [1e-9] * 8, ys[:8]))
t = torch.ones(4, dtype=torch.float64, device=ctx.device)
timeit("torch allreduce", lambda: dist.all_reduce(t))
timeit("update_ghosts", lambda: mf.update_ghost_values(0, x) if hasattr(mf, "update_ghost_values") else None)
