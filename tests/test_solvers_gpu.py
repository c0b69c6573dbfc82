"""Krylov loops on the device (callers of the operators): iteration counts and solutions
match the same algorithms run in numpy on the CPU oracle's operators."""
import math

import numpy as np
import pytest
import torch

from adaflo_b200 import capi
from adaflo_b200.mesh import BoxMesh, seeded_uniform
from adaflo_b200.operators import BlockVector, DiagonalPreconditioner, LevelSetOperators, MatrixFree, NavierStokesMatrix
from adaflo_b200.solvers import SolverBicgstab, SolverCG, SolverControl, SolverGMRES
from oracle import oracle as orc

from helpers import flow_params, rel_err, time_weights

pytestmark = pytest.mark.gpu


def dev(ctx, a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(ctx.device)


def numpy_pcg(A, b, inv_diag, rel_tol, max_iter=2000):
    x = np.zeros_like(b)
    r = b - A(x)
    res0 = np.linalg.norm(r)
    tol = rel_tol * res0
    z = inv_diag * r
    p = z.copy()
    rz = r @ z
    it = 0
    while np.linalg.norm(r) > tol and it < max_iter:
        it += 1
        Ap = A(p)
        alpha = rz / (p @ Ap)
        x += alpha * p
        r -= alpha * Ap
        if np.linalg.norm(r) <= tol:
            break
        z = inv_diag * r
        rz_new = r @ z
        p = z + (rz_new / rz) * p
        rz = rz_new
    return x, it


def test_reinitialization_cg_with_jacobi(ctx):
    """the reinitialisation solve of level_set_okz_reinitialization.cc:334-336: CG, rel. tol 1e-6, Jacobi"""
    n_sub = 3
    mesh = BoxMesh((4, 4, 4))
    mesh.build_ls_space(n_sub)
    mf = MatrixFree(ctx, mesh, "cartesian")
    h_min = float(mesh.h.min())
    eps, eps_used = 1.5, 1.5 * h_min / n_sub
    A = LevelSetOperators(mf, capi.LSParams(3, n_sub, 0, 0, eps, eps_used, 4.0, h_min), mesh.cell_diameters())
    A.set_time(75.0, -100.0, 25.0, 0.02)
    O = orc.OracleLevelSet(mesh, orc.LSParams(eps, eps_used, 4.0, h_min, 0.02, 3, 0))
    x = mesh.ls.coords
    r = np.linalg.norm(x, axis=1)
    c = -np.tanh((r - 0.5) / (2 * eps_used)) * (1 + 0.05 * seeded_uniform(mesh.ls.global_ids, 1, 2))
    normal = [-x[:, d] / np.maximum(r, 1e-3) for d in range(3)]
    n = mesh.ls.n_owned
    diag = O.mass_diagonal()
    inv_diag = np.where(np.abs(diag) > 1e-10 * np.abs(diag).max(), 1.0 / diag, 1.0)   # diagonal_preconditioner.cc:46-51
    rhs_o = np.zeros(n)
    O.reinitialization_rhs(rhs_o, c, normal, False, True)
    x_ref, it_ref = numpy_pcg(lambda v: O.reinitialization_vmult(v, False), rhs_o, inv_diag, 1e-6)
    # device
    d_diag = A.initialize_mass_matrix_diagonal()
    P = DiagonalPreconditioner(ctx, d_diag)
    rhs_g = mf.initialize_dof_vector(2)
    A.reinitialization_rhs(rhs_g, dev(ctx, c), [dev(ctx, v) for v in normal], False, True)
    ctl = SolverControl(2000, 1e-30, 1e-6)
    xg = mf.initialize_dof_vector(2)
    its = SolverCG(ctx, ctl).solve(lambda d, s: A.reinitialization_vmult(d, s, False), xg, rhs_g, P.vmult)
    assert abs(its - it_ref) <= 1 and its > 3
    assert rel_err(xg.cpu().numpy(), x_ref) < 1e-6
    A.close()
    mf.close()


def test_advection_bicgstab(ctx):
    """advance_concentration solve (advance_concentration.cc:626-636): BiCGStab + Jacobi on the nonsymmetric operator"""
    n_sub = 3
    mesh = BoxMesh((4, 4, 3))
    mesh.build_ls_space(n_sub)
    mf = MatrixFree(ctx, mesh, "cartesian")
    h_min = float(mesh.h.min())
    A = LevelSetOperators(mf, capi.LSParams(3, n_sub, 0, 0, 1.5, 1.5 * h_min / n_sub, 4.0, h_min), mesh.cell_diameters())
    gamma = 75.0
    A.set_time(gamma, -100.0, 25.0, 0.02)
    from adaflo_b200.mesh import beltrami_velocity
    vel = beltrami_velocity(mesh.u.coords, 0.0).reshape(-1)
    c = np.tanh(mesh.ls.coords[:, 0] * 3)
    rhs = mf.initialize_dof_vector(2)
    A.advance_concentration_rhs(rhs, dev(ctx, vel), dev(ctx, c), dev(ctx, c), dev(ctx, c))
    P = DiagonalPreconditioner(ctx, A.initialize_mass_matrix_diagonal())
    xg = mf.initialize_dof_vector(2)
    ctl = SolverControl(200, 1e-30, 1e-8)
    its = SolverBicgstab(ctx, ctl).solve(A.advance_concentration_vmult, xg, rhs, P.vmult)
    # true residual of the computed solution
    Ax = mf.initialize_dof_vector(2)
    A.advance_concentration_vmult(Ax, xg)
    true_res = float((rhs - Ax).norm() / rhs.norm())
    assert 1 <= its < 60 and true_res < 1e-7
    A.close()
    mf.close()


def test_navier_stokes_gmres_matches_oracle_gmres(ctx):
    """unpreconditioned GMRES(30) on NavierStokesMatrix::vmult vs scipy GMRES on the oracle's vmult: same
    residual reduction after the same number of matrix-vector products"""
    import scipy.sparse.linalg as spla
    mesh = BoxMesh((3, 3, 2))
    dprm, oprm = flow_params()
    dtw, otw = time_weights()
    mf = MatrixFree(ctx, mesh, "cartesian")
    old, old_old = BlockVector.like(mf), BlockVector.like(mf)
    A = NavierStokesMatrix(dprm, 0, 1, 0, 1, old, old_old)
    A.initialize(mf, dtw, True)
    O = orc.OracleNavierStokesMatrix(mesh, oprm, otw)
    nu, np_ = O.n_u, O.n_p
    bu, bp = seeded_uniform(mesh.u.global_ids, 3, 1), seeded_uniform(mesh.p.global_ids, 1, 2)
    O.apply_pressure_average_projection(bp)       # consistent right-hand side
    n_iter = 25
    x = [mf.initialize_dof_vector(0), mf.initialize_dof_vector(1)]
    ctl = SolverControl(n_iter, 1e-30, 1e-14)
    try:
        SolverGMRES(ctx, ctl, 30).solve(lambda d, s: A.vmult(BlockVector(*d), BlockVector(*s)), x, [dev(ctx, bu), dev(ctx, bp)])
    except Exception:
        pass
    hist = ctl.history

    def mv(v):
        ou, op = O.vmult(v[:nu].copy(), v[nu:].copy())
        return np.concatenate([ou, op])
    b = np.concatenate([bu, bp])
    res = []
    spla.gmres(spla.LinearOperator((nu + np_, nu + np_), matvec=mv), b, rtol=1e-14, restart=30, maxiter=1,
               callback=lambda r: res.append(r), callback_type="pr_norm")
    k = min(len(res), len(hist) - 1, n_iter)
    assert k >= 20
    # scipy reports the preconditioned relative residual per inner iteration
    for j in range(k):
        assert abs(hist[j + 1] / hist[0] - res[j]) < 1e-8 * max(1.0, res[j]) + 1e-10
    A.clear()
    mf.close()


def test_multi_axpy_matches_torch(ctx):
    """y += sum_i a_i x_i in one pass (k = 1, 8, 11: one and two kernel passes)"""
    g = torch.Generator(device="cpu").manual_seed(5)
    n = 10007
    for k in (1, 8, 11):
        y = torch.rand(n, generator=g, dtype=torch.float64).to(ctx.device)
        xs = [torch.rand(n, generator=g, dtype=torch.float64).to(ctx.device) for _ in range(k)]
        a = [0.25 * (i + 1) * (-1) ** i for i in range(k)]
        ref = y.clone()
        for ai, xi in zip(a, xs):
            ref += ai * xi
        ctx.multi_axpy(y, a, xs)
        assert float((y - ref).abs().max()) < 1e-12


def test_block_preconditioner_on_the_first_newton_system_of_beltrami_3d(ctx):
    """flexible GMRES + device block preconditioner (Cahouet-Chabard Schur complement, matrix-free
    inner solves) on the first Newton system of tests/beltrami_3d at 8^3 cells: rel. 1e-8 in a few dozen iterations,
    where unpreconditioned GMRES needs thousands"""
    from adaflo_b200.mesh import BoxMesh, beltrami_pressure, beltrami_velocity
    from adaflo_b200.operators import MatrixFree, NavierStokesMatrix
    from adaflo_b200.preconditioner import NavierStokesPreconditioner
    mesh = BoxMesh((8, 8, 8))
    mf = MatrixFree(ctx, mesh, "cartesian")
    prm = capi.FlowParams(capi.INCOMPRESSIBLE, capi.NEWTON, 0, 0, 1.0, 1.0, 0.0, 0.0, 0.0, 0.0, 0.5)
    dt = 0.05
    sol, old, old_old = BlockVector.like(mf), BlockVector.like(mf), BlockVector.like(mf)
    A = NavierStokesMatrix(prm, 0, 1, 0, 1, old, old_old)
    A.initialize(mf, capi.TimeWeights(1 / dt, -1 / dt, 0.0, 1.0, 1.0, 0.0), True)
    sol.block(0).copy_(dev(ctx, beltrami_velocity(mesh.u.coords, 0.0).reshape(-1)))
    sol.block(1).copy_(dev(ctx, beltrami_pressure(mesh.p.coords, 0.0).reshape(-1)))
    old.block(0).copy_(sol.block(0))
    d = torch.from_numpy(np.nonzero(mesh.u.dirichlet.reshape(-1))[0]).to(ctx.device)
    sol.block(0)[d] = dev(ctx, beltrami_velocity(mesh.u.coords, dt).reshape(-1))[d]
    rhs, zero, upd = BlockVector.like(mf), BlockVector.like(mf), BlockVector.like(mf)
    A.residual(rhs, sol, zero)
    A.apply_pressure_average_projection(rhs.block(1))
    P = NavierStokesPreconditioner(ctx, A, mf)
    ctl = SolverControl(60, 1e-14, 1e-8)
    SolverGMRES(ctx, ctl, 60).solve(lambda dd, ss: A.vmult(BlockVector(*dd), BlockVector(*ss)), [upd.block(0), upd.block(1)],
                                    [rhs.block(0), rhs.block(1)], P.vmult, [mf.n_owned(0), mf.n_owned(1)])
    assert ctl.last_step <= 40 and ctl.last_value <= 1e-8 * ctl.history[0]
    # the true residual agrees with the recurrence
    Ax = BlockVector.like(mf)
    A.vmult(Ax, upd)
    res = float(((Ax.block(0) - rhs.block(0)).norm() ** 2 + (Ax.block(1) - rhs.block(1)).norm() ** 2).sqrt())
    assert res <= 2e-8 * ctl.history[0]
    # the same solve entirely inside the library (operator and preconditioner as library callbacks): same iteration
    upd2 = BlockVector.like(mf)
    ctl2 = SolverControl(60, 1e-14, 1e-8)
    from adaflo_b200.solvers import library_operator
    SolverGMRES(ctx, ctl2, 60).solve(library_operator(capi.OP_NS_VMULT, A.h), [upd2.block(0), upd2.block(1)], [rhs.block(0), rhs.block(1)],
                                     P.as_library_callback(), [mf.n_owned(0), mf.n_owned(1)])
    assert abs(ctl2.last_step - ctl.last_step) <= 2 and sum(P.inner_iterations) > 0
    P.close()


def test_stall_watchdog_reports_a_stream_that_does_not_drain():
    """csrc/comm.cu stream_wait: with ADAFLO_B200_COMM_TIMEOUT armed, a host-side wait that outlasts the time-out returns
    ADAFLO_B200_ERR_COMM instead of blocking (the multi-rank case -- a collective whose peer never arrives -- is
    tests/test_multi_gpu.py::test_rank_mismatch_fails_loudly_instead_of_hanging; here the stream is kept busy by a spin kernel)"""
    import os
    import time
    from adaflo_b200.operators import Context
    os.environ["ADAFLO_B200_COMM_TIMEOUT"] = "0.2"
    try:
        wctx = Context(0)
    finally:
        os.environ.pop("ADAFLO_B200_COMM_TIMEOUT")
    wctx.synchronize()                                   # an idle stream: returns at once
    torch.cuda._sleep(int(4e9))                          # about 2 s of spinning on the (default) stream of the context
    t0 = time.time()
    with pytest.raises(capi.AdafloB200Error) as info:
        wctx.synchronize()
    waited = time.time() - t0
    assert info.value.code == capi.ERR_COMM and 0.15 < waited < 1.5, (info.value, waited)
    torch.cuda.synchronize()
    wctx.synchronize()                                   # drained: fine again
    # without the variable a single-rank context waits plainly
    plain = Context(0)
    torch.cuda._sleep(int(1e9))
    plain.synchronize()
