"""The CUDA path against the reference's own golden log tests/beltrami_3d.output (3-D Beltrami flow, 16^3 cells,
BDF-2, dt = 0.05): four time steps are advanced ON THE DEVICE (Newton with NavierStokesMatrix::residual / vmult,
flexible GMRES with the device block preconditioner of adaflo_b200.preconditioner, converged to the log's nonlinear
tolerance 1e-9) and the quantities
of the log that do not depend on the reference's ILU preconditioner are compared digit by digit:
  * the first nonlinear residual (u, p) of time steps #1 - #4          (beltrami_3d.output:13,33,44,56)
  * the L2 errors of velocity and pressure at t = 0.2                  (beltrami_3d.output:63)
The golden numbers are the fixture tests/golden/beltrami_3d_output.json (copied from that file); the oracle is used only to integrate the error norms."""
import numpy as np
import pytest
import torch

from adaflo_b200 import capi
from adaflo_b200.mesh import BoxMesh, beltrami_pressure, beltrami_velocity
from adaflo_b200.operators import BlockVector, Context, MatrixFree, NavierStokesMatrix
from adaflo_b200.preconditioner import NavierStokesPreconditioner
from adaflo_b200.solvers import NoConvergence, SolverControl, SolverGMRES
from oracle import oracle as orc

from helpers import flow_params

pytestmark = pytest.mark.gpu

import json
import os

with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "beltrami_3d_output.json")) as _f:
    _GOLDEN = json.load(_f)
GOLDEN_FIRST_RESIDUALS = [(_GOLDEN["first_nonlinear_residual"][str(k)]["u"], _GOLDEN["first_nonlinear_residual"][str(k)]["p"]) for k in range(1, 5)]
GOLDEN_ERRORS_T02 = (_GOLDEN["l2_errors_t0.2"]["e_p"], _GOLDEN["l2_errors_t0.2"]["e_u"])   # ||e_p||_L2, ||e_u||_L2


def test_device_time_stepping_reproduces_the_golden_log():
    ctx = Context(0)
    mesh = BoxMesh((16, 16, 16))
    mf = MatrixFree(ctx, mesh, "cartesian")
    prm = capi.FlowParams(capi.INCOMPRESSIBLE, capi.NEWTON, 0, 0, 1.0, 1.0, 0.0, 0.0, 0.0, 0.0, 0.5)   # tests/beltrami_3d.prm
    dt = 0.05
    euler = capi.TimeWeights(1 / dt, -1 / dt, 0.0, 1.0, 1.0, 0.0)          # step #1 (time_stepping.cc:158-198)
    bdf2 = capi.TimeWeights(1.5 / dt, -2 / dt, 0.5 / dt, 1.0, 2.0, -1.0)
    sol, old, old_old = BlockVector.like(mf), BlockVector.like(mf), BlockVector.like(mf)
    A = NavierStokesMatrix(prm, 0, 1, 0, 1, old, old_old)
    A.initialize(mf, euler, True)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(ctx.device)
    sol.block(0).copy_(dev(beltrami_velocity(mesh.u.coords, 0.0).reshape(-1)))
    sol.block(1).copy_(dev(beltrami_pressure(mesh.p.coords, 0.0).reshape(-1)))
    dirichlet = torch.from_numpy(np.nonzero(mesh.u.dirichlet.reshape(-1))[0]).to(ctx.device)
    nu, npp = mf.n_owned(0), mf.n_owned(1)
    rhs, zero, upd = BlockVector.like(mf), BlockVector.like(mf), BlockVector.like(mf)
    lin_iterations = []
    for step in range(1, 5):
        # init_time_advance (navier_stokes.cc:659-685): shift the history, extrapolate, new Dirichlet values
        for b in range(2):
            old_old.block(b).copy_(old.block(b))
            old.block(b).copy_(sol.block(b))
            if step >= 3:
                sol.block(b).mul_(2.0).sub_(old_old.block(b))
        A.set_time_stepping(euler if step == 1 else bdf2)
        sol.block(0)[dirichlet] = dev(beltrami_velocity(mesh.u.coords, step * dt).reshape(-1))[dirichlet]
        for it in range(10):
            rhs.block(0).zero_()
            rhs.block(1).zero_()
            A.residual(rhs, sol, zero)
            A.apply_pressure_average_projection(rhs.block(1))              # compute_residual, navier_stokes.cc:781-790
            ru, rp = float(rhs.block(0).norm()), float(rhs.block(1).norm())
            if it == 0:
                assert (f"{ru:.3e}", f"{rp:.3e}") == GOLDEN_FIRST_RESIDUALS[step - 1], (step, ru, rp)
            if (ru * ru + rp * rp) ** 0.5 < 1e-9:                          # NL tolerance of the parameter file
                break
            upd.block(0).zero_()
            upd.block(1).zero_()
            ctl = SolverControl(100, 1e-14, 1e-7)
            P = NavierStokesPreconditioner(ctx, A, mf)                      # (diagonal of the new linearisation)
            try:
                SolverGMRES(ctx, ctl, 100).solve(lambda d, s: A.vmult(BlockVector(*d), BlockVector(*s)), [upd.block(0), upd.block(1)],
                                                 [rhs.block(0), rhs.block(1)], P.vmult, [nu, npp])
            except NoConvergence:
                pass                                                        # Newton continues with the inexact update
            lin_iterations.append(ctl.last_step)
            sol.block(0).add_(upd.block(0))
            sol.block(1).add_(upd.block(1))
        else:
            pytest.fail(f"Newton did not converge in time step {step}")
        # fix_pressure_constant (navier_stokes.cc:982-1043): shift the pressure so that it matches the analytic value at
        # the midpoint of the first boundary face of the first cell (face 0 = x = -1 of the cell at the corner)
        face_nodes = torch.from_numpy(mesh.p.cell_nodes[0][[0, 2, 4, 6]].astype(np.int64)).to(ctx.device)
        mid = np.array([[-1.0, -1.0 + 0.5 * mesh.h[1], -1.0 + 0.5 * mesh.h[2]]])
        shift = float(beltrami_pressure(mid, step * dt)[0]) - float(sol.block(1)[face_nodes].mean())
        A.apply_pressure_shift(shift, sol.block(1))
    # L2 errors at t = 0.2 (tests/beltrami.cc:254: QGauss(degree + 2))
    _, oprm = flow_params()
    O = orc.OracleNavierStokesMatrix(mesh, oprm, orc.TimeWeights(1.5 / dt, -2 / dt, 0.5 / dt, 1.0, 2.0, -1.0))
    xq = mesh.quadrature_points(4)
    e_u = O.l2_error(3, sol.block(0).cpu().numpy(), beltrami_velocity(xq, 0.2), n_q_1d=4)
    e_p = O.l2_error(1, sol.block(1).cpu().numpy(), beltrami_pressure(xq, 0.2)[..., None], n_q_1d=4)
    assert f"{e_u:.4g}" == GOLDEN_ERRORS_T02[1], e_u
    assert f"{e_p:.4g}" == GOLDEN_ERRORS_T02[0], e_p
    # the block preconditioner keeps the outer iteration counts in the range of the reference's (30 for 1e-5 with ILU)
    assert max(lin_iterations) <= 40, lin_iterations
