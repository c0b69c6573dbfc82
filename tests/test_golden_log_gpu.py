"""The CUDA path against the reference's own golden logs (3-D Beltrami flow, BDF-2, dt = 0.05):
  * tests/beltrami_3d.output       16^3 cells, Q2/Q1
  * tests/beltrami_3d_augp.output  624 cells with 2:1 hanging nodes, augmented Taylor-Hood Q2/(Q1+DG0)
Four time steps are advanced ON THE DEVICE (Newton with NavierStokesMatrix::residual / vmult, GMRES from
adaflo_b200.solvers -- flexible with the device block preconditioner on the uniform mesh, unpreconditioned on the
small refined one -- converged to the logs' nonlinear tolerance 1e-9, fix_pressure_constant restated) and the
quantities of the logs that do not depend on the reference's ILU/AMG preconditioners are compared digit by digit:
the first nonlinear residual (u, p) of time steps #1 - #4 and the L2 errors at t = 0.2.  The golden numbers are the
fixtures tests/golden/beltrami_3d*_output.json (copied from the logs); the oracle only integrates the error norms."""
import json
import os

import numpy as np
import pytest
import torch

from adaflo_b200 import capi
from adaflo_b200.mesh import BoxMesh, RefinedBoxMesh, beltrami_pressure, beltrami_velocity
from adaflo_b200.operators import BlockVector, Context, MatrixFree, NavierStokesMatrix
from adaflo_b200.preconditioner import NavierStokesPreconditioner
from adaflo_b200.solvers import NoConvergence, SolverControl, SolverGMRES
from oracle import oracle as orc

from helpers import flow_params

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def golden(name):
    with open(os.path.join(HERE, "golden", name)) as f:
        return json.load(f)


def advance_four_steps(mesh, augmented, gold, precondition):
    ctx = Context(0)
    mf = MatrixFree(ctx, mesh, "cartesian")
    prm = capi.FlowParams(capi.INCOMPRESSIBLE, capi.NEWTON, int(augmented), 0, 1.0, 1.0, 0.0, 0.0, 0.0, 0.0, 0.5)   # the .prm files
    dt = gold["parameters"]["dt"]
    euler = capi.TimeWeights(1 / dt, -1 / dt, 0.0, 1.0, 1.0, 0.0)          # step #1 (time_stepping.cc:158-198)
    bdf2 = capi.TimeWeights(1.5 / dt, -2 / dt, 0.5 / dt, 1.0, 2.0, -1.0)
    sol, old, old_old = BlockVector.like(mf), BlockVector.like(mf), BlockVector.like(mf)
    A = NavierStokesMatrix(prm, 0, 1, 0, 1, old, old_old)
    A.initialize(mf, euler, True)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(ctx.device)
    sol.block(0).copy_(dev(beltrami_velocity(mesh.u.coords, 0.0).reshape(-1)))
    p0 = beltrami_pressure(mesh.p.coords, 0.0).reshape(-1)
    if augmented:
        p0[mesh.p.n_q1_owned:] = 0.0                                      # interpolate_pressure_field, navier_stokes.cc:1414-1430
    sol.block(1).copy_(dev(p0))
    dirichlet = torch.from_numpy(np.nonzero(mesh.u.dirichlet.reshape(-1))[0]).to(ctx.device)
    nu, npp = mf.n_owned(0), mf.n_owned(1)
    rhs, zero, upd = BlockVector.like(mf), BlockVector.like(mf), BlockVector.like(mf)
    lin_iterations = []
    # first boundary face of the first cell (face 0 = x = -1 of the cell at the corner): fix_pressure_constant
    face_nodes = torch.from_numpy(mesh.p.cell_nodes[0][[0, 2, 4, 6]].astype(np.int64)).to(ctx.device)
    h0 = (mesh.h if hasattr(mesh, "h") else mesh.h_cell[0])
    mid = np.array([[-1.0, -1.0 + 0.5 * h0[1], -1.0 + 0.5 * h0[2]]])
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
                g = gold["first_nonlinear_residual"][str(step)]
                assert (f"{ru:.3e}", f"{rp:.3e}") == (g["u"], g["p"]), (step, ru, rp)
            if (ru * ru + rp * rp) ** 0.5 < 1e-9:                          # NL tolerance of the parameter files
                break
            upd.block(0).zero_()
            upd.block(1).zero_()
            if precondition:
                ctl = SolverControl(100, 1e-14, 1e-7)
                P = NavierStokesPreconditioner(ctx, A, mf).vmult            # (diagonal of the new linearisation)
            else:
                ctl = SolverControl(6000, 1e-14, 1e-7)
                P = None
            try:
                SolverGMRES(ctx, ctl, 100 if precondition else 200).solve(lambda d, s: A.vmult(BlockVector(*d), BlockVector(*s)),
                                                                         [upd.block(0), upd.block(1)], [rhs.block(0), rhs.block(1)], P, [nu, npp])
            except NoConvergence:
                pass                                                        # Newton continues with the inexact update
            lin_iterations.append(ctl.last_step)
            sol.block(0).add_(upd.block(0))
            sol.block(1).add_(upd.block(1))
        else:
            pytest.fail(f"Newton did not converge in time step {step}")
        # fix_pressure_constant (navier_stokes.cc:982-1043): shift the pressure (Q1 constants only) so that it matches
        # the analytic value at the midpoint of that face
        p_mid = float(sol.block(1)[face_nodes].mean())
        if augmented:
            p_mid += float(sol.block(1)[int(mesh.p.cell_nodes[0][8])])    # get_function_values: Q1 part + the cell's constant
        shift = float(beltrami_pressure(mid, step * dt)[0]) - p_mid
        A.apply_pressure_shift(shift, sol.block(1))
    u_h, p_h = sol.block(0).cpu().numpy(), sol.block(1).cpu().numpy()
    A.clear()
    mf.close()
    return u_h, p_h, lin_iterations


def test_device_time_stepping_reproduces_the_golden_log():
    gold = golden("beltrami_3d_output.json")
    mesh = BoxMesh((16, 16, 16))
    u_h, p_h, lin_iterations = advance_four_steps(mesh, False, gold, precondition=True)
    # L2 errors at t = 0.2 (tests/beltrami.cc:254: QGauss(degree + 2))
    _, oprm = flow_params()
    O = orc.OracleNavierStokesMatrix(mesh, oprm, orc.TimeWeights(30.0, -40.0, 10.0, 1.0, 2.0, -1.0))
    xq = mesh.quadrature_points(4)
    e_u = O.l2_error(3, u_h, beltrami_velocity(xq, 0.2), n_q_1d=4)
    e_p = O.l2_error(1, p_h, beltrami_pressure(xq, 0.2)[..., None], n_q_1d=4)
    assert f"{e_u:.4g}" == gold["l2_errors_t0.2"]["e_u"], e_u
    assert f"{e_p:.4g}" == gold["l2_errors_t0.2"]["e_p"], e_p
    # the block preconditioner keeps the outer iteration counts in the range of the reference's (30 for 1e-5 with ILU)
    assert max(lin_iterations) <= 40, lin_iterations


def test_device_time_stepping_on_the_refined_augmented_mesh_reproduces_the_golden_log():
    """624 cells with 2:1 hanging nodes, Q2/(Q1+DG0): in-kernel hanging-node interpolation, element-constant pressure
    and both null-space modes on the device, over four time steps"""
    gold = golden("beltrami_3d_augp_output.json")
    mesh = RefinedBoxMesh((8, 8, 8), ((4, 8), (0, 2), (0, 2)), augmented=True)     # tests/beltrami.cc:385-412
    assert mesh.n_cells == gold["n_active_cells"]["value"]
    u_h, p_h, _ = advance_four_steps(mesh, True, gold, precondition=False)
    _, oprm = flow_params(augmented=True)
    O = orc.OracleNavierStokesMatrix(mesh, oprm, orc.TimeWeights(30.0, -40.0, 10.0, 1.0, 2.0, -1.0), None, True, mesh.u_lists, mesh.p_lists)
    xq = mesh.quadrature_points(4)
    e_u = O.l2_error(3, u_h, beltrami_velocity(xq, 0.2), n_q_1d=4)
    e_p = O.l2_error(1, p_h, beltrami_pressure(xq, 0.2)[..., None], n_q_1d=4)
    assert f"{e_u:.4g}" == gold["l2_errors_t0.2"]["e_u"], e_u
    assert f"{e_p:.4g}" == gold["l2_errors_t0.2"]["e_p"], e_p


def test_device_time_stepping_of_beltrami_2d_q4q3_reproduces_the_golden_log():
    """tests/beltrami_2d.output: 2-D Taylor vortex, Q4/Q3 on 1048 cells with 2:1 hanging nodes -- the degree / dimension dispatch
    (csrc/ns_generic.cu) with general constraint rows.  Four BDF-2 steps on the device (Newton on residual / vmult, flexible
    device GMRES with the library's block preconditioner -- velocity_vmult / divergence_vmult_add / pressure_mass_vmult /
    pressure_poisson_vmult / compute_diagonal of the generic path --, AffineConstraints::distribute by
    adaflo_b200_mf_distribute_constraints, fix_pressure_constant restated); the first nonlinear residuals of time steps #1 - #4
    are compared digit by digit.  The outer iteration counts stay in the range of the reference's (20 - 25 with its AMG)."""
    from adaflo_b200.mesh import RefinedTensorMesh2D, gauss_lobatto_01, lagrange_values
    from adaflo_b200.operators import MatrixFreeGeneric
    gold = golden("beltrami_2d_output.json")
    par = gold["parameters"]
    n, k = par["coarse_cells_per_direction"], par["velocity_degree"]
    mesh = RefinedTensorMesh2D((n, n), par["refined_block"], k)                       # tests/beltrami.cc:385-412
    assert mesh.n_cells == gold["n_active_cells"]["value"]
    assert (2 * mesh.n_u_nodes, mesh.n_p_nodes) == (gold["n_dofs"]["velocity"], gold["n_dofs"]["pressure"])
    ctx = Context(0)
    mf = MatrixFreeGeneric(ctx, mesh)
    dt = par["dt"]
    prm = capi.FlowParams(capi.INCOMPRESSIBLE, capi.NEWTON, 0, 0, 1.0, par["viscosity"], 0.0, 0.0, 0.0, 0.0, 0.5)
    euler = capi.TimeWeights(1 / dt, -1 / dt, 0.0, 1.0, 1.0, 0.0)
    bdf2 = capi.TimeWeights(1.5 / dt, -2 / dt, 0.5 / dt, 1.0, 2.0, -1.0)
    sol, old, old_old = BlockVector.like(mf), BlockVector.like(mf), BlockVector.like(mf)
    A = NavierStokesMatrix(prm, 0, 1, 0, 1, old, old_old)
    A.initialize(mf, euler, True)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(ctx.device)
    sol.block(0).copy_(dev(beltrami_velocity(mesh.u_coords, 0.0).reshape(-1)))
    sol.block(1).copy_(dev(beltrami_pressure(mesh.p_coords, 0.0)))                   # nodal values, also at the hanging nodes
    dirichlet = torch.from_numpy(np.nonzero(mesh.u_dirichlet.reshape(-1))[0]).to(ctx.device)
    nu, npp = mf.n_owned(0), mf.n_owned(1)
    rhs, zero, upd = BlockVector.like(mf), BlockVector.like(mf), BlockVector.like(mf)
    lin_iterations = []
    # fix_pressure_constant (navier_stokes.cc:982-1043): first boundary face of the first cell = the face x = -1 of the coarse cell
    # at the corner; the Q3 pressure at its midpoint
    w_mid = torch.from_numpy(lagrange_values(gauss_lobatto_01(k), np.array([0.5]))[0][0]).to(ctx.device)
    face_nodes = torch.from_numpy(mesh.cell_p[0].reshape(k, k)[:, 0].astype(np.int64)).to(ctx.device)
    mid = np.array([[-1.0, -1.0 + 0.5 * mesh.h[1]]])
    for step in range(1, 5):
        for b in range(2):                                                           # init_time_advance, navier_stokes.cc:659-685
            old_old.block(b).copy_(old.block(b))
            old.block(b).copy_(sol.block(b))
            if step >= 3:
                sol.block(b).mul_(2.0).sub_(old_old.block(b))
        A.set_time_stepping(euler if step == 1 else bdf2)
        sol.block(0)[dirichlet] = dev(beltrami_velocity(mesh.u_coords, step * dt).reshape(-1))[dirichlet]
        mf.distribute_constraints(0, sol.block(0))                                   # :1252
        for it in range(10):
            rhs.block(0).zero_()
            rhs.block(1).zero_()
            A.residual(rhs, sol, zero)
            A.apply_pressure_average_projection(rhs.block(1))
            ru, rp = float(rhs.block(0).norm()), float(rhs.block(1).norm())
            if it == 0:
                g = gold["first_nonlinear_residual"][str(step)]
                assert (f"{ru:.3e}", f"{rp:.3e}") == (g["u"], g["p"]), (step, ru, rp)
            if (ru * ru + rp * rp) ** 0.5 < 1e-9:
                break
            upd.block(0).zero_()
            upd.block(1).zero_()
            ctl = SolverControl(200, 2e-10, 1e-7)
            P = NavierStokesPreconditioner(ctx, A, mf)                               # diagonal of the new linearisation
            try:
                SolverGMRES(ctx, ctl, 100).solve(lambda d, s: A.vmult(BlockVector(*d), BlockVector(*s)), [upd.block(0), upd.block(1)],
                                                 [rhs.block(0), rhs.block(1)], P.vmult, [nu, npp])
            except NoConvergence:
                pass
            P.close()
            lin_iterations.append(ctl.last_step)
            mf.distribute_constraints(0, upd.block(0))                               # :643-644
            mf.distribute_constraints(1, upd.block(1))
            sol.block(0).add_(upd.block(0))
            sol.block(1).add_(upd.block(1))
        else:
            pytest.fail(f"Newton did not converge in time step {step}")
        p_mid = float((w_mid * sol.block(1)[face_nodes]).sum())
        A.apply_pressure_shift(float(beltrami_pressure(mid, step * dt)[0]) - p_mid, sol.block(1))
        mf.distribute_constraints(1, sol.block(1))                                   # :1042
    assert max(lin_iterations) <= 40, lin_iterations
    A.clear()
    mf.close()
