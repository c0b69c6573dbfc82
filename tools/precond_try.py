import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adaflo_b200 import capi
from adaflo_b200.mesh import BoxMesh, beltrami_velocity, beltrami_pressure
from adaflo_b200.operators import BlockVector, Context, MatrixFree, NavierStokesMatrix
from adaflo_b200.preconditioner import NavierStokesPreconditioner
from adaflo_b200.solvers import SolverControl, SolverGMRES, NoConvergence
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
ctx = Context(0)
mesh = BoxMesh((n, n, n))
mf = MatrixFree(ctx, mesh, "cartesian")
prm = capi.FlowParams(capi.INCOMPRESSIBLE, capi.NEWTON, 0, 0, 1.0, 1.0, 0.0, 0.0, 0.0, 0.0, 0.5)
dt = 0.05
sol, old, old_old = BlockVector.like(mf), BlockVector.like(mf), BlockVector.like(mf)
A = NavierStokesMatrix(prm, 0, 1, 0, 1, old, old_old)
A.initialize(mf, capi.TimeWeights(1 / dt, -1 / dt, 0.0, 1.0, 1.0, 0.0), True)
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(ctx.device)
sol.block(0).copy_(dev(beltrami_velocity(mesh.u.coords, 0.0).reshape(-1)))
sol.block(1).copy_(dev(beltrami_pressure(mesh.p.coords, 0.0).reshape(-1)))
old.block(0).copy_(sol.block(0))
d = torch.from_numpy(np.nonzero(mesh.u.dirichlet.reshape(-1))[0]).to(ctx.device)
sol.block(0)[d] = dev(beltrami_velocity(mesh.u.coords, dt).reshape(-1))[d]
rhs, zero, upd = BlockVector.like(mf), BlockVector.like(mf), BlockVector.like(mf)
A.residual(rhs, sol, zero)
A.apply_pressure_average_projection(rhs.block(1))
print("residual", float(rhs.block(0).norm()), float(rhs.block(1).norm()))
P = NavierStokesPreconditioner(ctx, A, mf)
nu, npp = mf.n_owned(0), mf.n_owned(1)
ctl = SolverControl(200, 1e-14, 1e-8)
torch.cuda.synchronize(); t0 = time.perf_counter()
try:
    SolverGMRES(ctx, ctl, 100).solve(lambda dd, ss: A.vmult(BlockVector(*dd), BlockVector(*ss)), [upd.block(0), upd.block(1)],
                                     [rhs.block(0), rhs.block(1)], P.vmult, [nu, npp])
except NoConvergence as e:
    print("no conv", e)
torch.cuda.synchronize()
print("outer its", ctl.last_step, "res", ctl.last_value, "time", time.perf_counter() - t0, "inner", P.inner_iterations)
print([f"{v:.2e}" for v in ctl.history[:40]])
