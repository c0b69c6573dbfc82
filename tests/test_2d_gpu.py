"""2-D Navier-Stokes operators on the GPU (NavierStokesMatrix2D: 3-D kernels on a one-cell-thick extrusion) against
the independent dense 2-D oracle."""
import numpy as np
import pytest
import torch

from adaflo_b200 import capi
from adaflo_b200.mesh import BoxMesh2D
from adaflo_b200.operators import Context, NavierStokesMatrix2D
from oracle.ns_oracle_2d import OracleNavierStokes2D

from helpers import flow_params, rel_err, time_weights
from test_2d_cpu import fields

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module")
def ctx():
    return Context(0)


@pytest.mark.parametrize("lin", [capi.NEWTON, capi.PICARD])
def test_2d_operators_match_dense_2d_oracle(ctx, lin):
    m = BoxMesh2D((6, 5))
    dprm, _ = flow_params(linearization=lin, damping=-0.2, tau_grad_div=0.3, density=1.3, viscosity=0.7)
    dtw, otw = time_weights()
    A = NavierStokesMatrix2D(dprm, ctx, m, dtw, True)
    O = OracleNavierStokes2D(m, capi.INCOMPRESSIBLE, lin, 1.3, 0.7, -0.2, 0.3, 0.5, (otw.weight, otw.weight_old, otw.weight_old_old, otw.tau1))
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(ctx.device)
    u, p = fields(m, 3)
    uo, _ = fields(m, 5)
    uoo, _ = fields(m, 7)
    O.old_u, O.old_old_u = uo.copy(), uoo.copy()
    A.set_old_solutions(dev(uo), dev(uoo))
    pre_u, pre_p = fields(m, 21)
    usr_u, usr_p = fields(m, 23)
    ru, rp = pre_u.copy(), pre_p.copy()
    O.residual(ru, rp, u, p, usr_u, usr_p)
    g = (dev(pre_u), dev(pre_p))
    A.residual(g, (dev(u), dev(p)), (dev(usr_u), dev(usr_p)))
    assert rel_err(g[0].cpu().numpy(), ru) < TOL and rel_err(g[1].cpu().numpy(), rp) < TOL
    su, sp = fields(m, 11)
    ou, op = O.vmult(su, sp)
    d = A.initialize_dof_vectors()
    A.vmult(d, (dev(su), dev(sp)))
    assert rel_err(d[0].cpu().numpy(), ou) < TOL and rel_err(d[1].cpu().numpy(), op) < TOL
    dv = A.initialize_dof_vectors()[0]
    A.velocity_vmult(dv, dev(su))
    assert rel_err(dv.cpu().numpy(), O.velocity_vmult(su)) < TOL
    # transpose identity in the 2-D inner product
    yu, yp = fields(m, 13)
    c = m.u_dirichlet.reshape(-1)      # constrained rows are identity rows: compare on vectors that vanish there
    su0, yu0 = su.copy(), yu.copy()
    su0[c] = 0.0
    yu0[c] = 0.0
    t0, a0 = A.initialize_dof_vectors(), A.initialize_dof_vectors()
    A.Tvmult(t0, (dev(su0), dev(sp)))
    A.vmult(a0, (dev(yu0), dev(yp)))
    lhs = float(t0[0].cpu().numpy() @ yu0 + t0[1].cpu().numpy() @ yp)
    rhs = float(su0 @ a0[0].cpu().numpy() + sp @ a0[1].cpu().numpy())
    assert abs(lhs - rhs) <= 1e-11 * max(abs(lhs), abs(rhs), 1.0)
    A.clear()


@pytest.mark.parametrize("dirichlet_edges", [(), (0,)])
def test_2d_level_set_operators_match_dense_2d_oracle(ctx, dirichlet_edges):
    from adaflo_b200.operators import LevelSetOperators2D
    from oracle import oracle as orc
    from test_2d_cpu import ls_fields, ls_setup, taylor_velocity_2d
    from adaflo_b200.mesh import seeded_uniform
    n_sub = 3
    m, O, _, (eps, eps_used, h_min, step) = ls_setup(n=(5, 4), n_sub=n_sub, dirichlet_edges=dirichlet_edges)
    L = LevelSetOperators2D(ctx, m, capi.LSParams(2, n_sub, 0, 0, eps, eps_used, 4.0, h_min))
    L.set_time(75.0, -100.0, 25.0, step)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(ctx.device)
    c, vec = ls_fields(m)
    assert rel_err(L.initialize_mass_matrix_diagonal().cpu().numpy(), O.diag) < TOL
    vel = taylor_velocity_2d(m.u_coords, 0.0).reshape(-1) + 0.1 * seeded_uniform(m.u_ids, 2, 6)
    c_old, c_oo = c + 0.05 * vec(7), c + 0.05 * vec(8)
    pre = vec(9)
    o = pre.copy()
    O.advance_concentration_rhs(o, vel, c, c_old, c_oo, 75.0, -100.0, 25.0)
    g = dev(pre)
    L.advance_concentration_rhs(g, dev(vel), dev(c), dev(c_old), dev(c_oo))
    assert rel_err(g.cpu().numpy(), o) < TOL
    src = vec(10)
    g = L.new_vector()
    L.advance_concentration_vmult(g, dev(src))
    assert rel_err(g.cpu().numpy(), O.advance_concentration_vmult(src, 75.0)) < TOL
    ro = [np.zeros(O.n) for _ in range(2)]
    O.compute_normal_rhs(ro, c)
    rg = [L.new_vector() for _ in range(2)]
    L.compute_normal_rhs(rg, dev(c))
    normal = [m.ls_coords[:, d] - [0.1, -0.05][d] + 0.02 * vec(20 + d) for d in range(2)]
    vg = [L.new_vector() for _ in range(2)]
    L.compute_normal_vmult(vg, [dev(v) for v in normal])
    vo = O.compute_normal_vmult(normal, 4.0)
    for d in range(2):
        assert rel_err(rg[d].cpu().numpy(), ro[d]) < TOL and rel_err(vg[d].cpu().numpy(), vo[d]) < TOL
    o = np.zeros(O.n)
    O.compute_curvature_rhs(o, normal)
    g = L.new_vector()
    L.compute_curvature_rhs(g, [dev(v) for v in normal])
    assert rel_err(g.cpu().numpy(), o) < TOL
    g = L.new_vector()
    L.compute_curvature_vmult(g, dev(src), True)
    assert rel_err(g.cpu().numpy(), O.compute_curvature_vmult(src, True)) < TOL
    o = np.zeros(O.n)
    O.reinitialization_rhs(o, c, normal, False, True)
    g = L.new_vector()
    L.reinitialization_rhs(g, dev(c), [dev(v) for v in normal], False, True)
    assert rel_err(g.cpu().numpy(), o) < TOL
    for diffuse_only in (False, True):
        g = L.new_vector()
        L.reinitialization_vmult(g, dev(src), diffuse_only)
        assert rel_err(g.cpu().numpy(), O.reinitialization_vmult(src, diffuse_only)) < TOL
    # point-wise operations
    H = L.new_vector()
    L.compute_heaviside(H, dev(c))
    H_ref = orc.compute_heaviside(m.cell_ls, c, eps, n_sub)
    assert np.abs(H.cpu().numpy() - H_ref).max() < 1e-14
    kap = 2.0 / np.maximum(np.linalg.norm(m.ls_coords, axis=1), 0.2)
    kg = dev(kap)
    L.curvature_correction(kg, dev(c))
    ko = kap.copy()
    orc.ls_lib().ls_oracle_curvature_correction(__import__("ctypes").c_int64(O.n), __import__("ctypes").c_double(eps_used), 2, orc._p(c), orc._p(ko))
    assert rel_err(kg.cpu().numpy(), ko) < TOL
    L.close()


@pytest.mark.parametrize("interpolate", [False, True])
def test_2d_force_matches_dense_2d_oracle(ctx, interpolate):
    """surface tension + gravity (on y = component dim - 1) and the rho / mu update of the 2-D operator; with
    'interpolate' the Heaviside gradient comes from the pressure-space interpolant as in tests/rising_bubble_ls.prm"""
    from adaflo_b200.operators import LevelSetOperators2D
    from oracle import oracle as orc
    from oracle.ls_oracle_2d import compute_force_2d
    from test_2d_cpu import ls_fields, ls_setup
    n_sub = 3
    m, O, _, (eps, eps_used, h_min, step) = ls_setup(n=(5, 4), n_sub=n_sub)
    dprm, _ = flow_params(density=1.0, viscosity=0.1, density_diff=9.0, viscosity_diff=0.4)
    dtw, _ = time_weights()
    A = NavierStokesMatrix2D(dprm, ctx, m, dtw, True)          # built after build_ls_space: its MatrixFree has the LS space
    L = LevelSetOperators2D(ctx, m, capi.LSParams(2, n_sub, 0, 0, eps, eps_used, 4.0, h_min), matrix_free=A.mf)
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(ctx.device)
    c, vec = ls_fields(m)
    H = orc.compute_heaviside(m.cell_ls, c, eps, n_sub)
    kap = 2.0 / np.maximum(np.linalg.norm(m.ls_coords, axis=1), 0.2) + 0.1 * vec(3)
    from adaflo_b200.mesh import seeded_uniform
    pre = seeded_uniform(m.u_ids, 2, 4)
    o = pre.copy()
    rho_q, mu_q = compute_force_2d(m, n_sub, H, kap, 0.07, 0.98, 1.0, 9.0, 0.1, 0.4, True, o, interpolate=interpolate)
    g = dev(pre)
    L.compute_force(A, g, dev(H), dev(kap), 0.07, 0.98, interpolate)
    assert rel_err(g.cpu().numpy(), o) < TOL
    rho3 = A.A.get_coefficient(capi.COEF_DENSITY).cpu().numpy().reshape(-1, 3, 9)
    mu3 = A.A.get_coefficient(capi.COEF_VISCOSITY).cpu().numpy().reshape(-1, 3, 9)
    assert rel_err(rho3, np.repeat(rho_q[:, None, :], 3, axis=1)) < TOL and rel_err(mu3, np.repeat(mu_q[:, None, :], 3, axis=1)) < TOL
    L.close()
    A.clear()


def test_device_reproduces_the_reinitialisation_counts_of_the_rising_bubble_log(ctx):
    """tests/rising_bubble_ls.output:6,13 on the device: 2-D level-set operators (extrusion path), device CG with the
    inverse mass diagonal -- 'reinitialize (8 + 8)' at set-up and '(7 + 7)' in time step #1"""
    from adaflo_b200.operators import LevelSetOperators2D
    from adaflo_b200.solvers import SolverCG, SolverControl
    from test_2d_cpu import rising_bubble_setup
    gold, m, eps_used, c0 = rising_bubble_setup()
    prm = gold["parameters"]
    L = LevelSetOperators2D(ctx, m, capi.LSParams(2, prm["concentration_subdivisions"], 0, 0, prm["epsilon"], eps_used, 4.0, prm["h"]),
                            cell_diameters=np.full(m.n_cells, prm["h"]))
    L.set_time(1.0 / prm["step_size"], -1.0 / prm["step_size"], 0.0, prm["step_size"])
    c = torch.from_numpy(c0).to(ctx.device)
    inv_diag = 1.0 / L.diag
    zero = L.new_vector()

    def solve(A, b, P, rel_tol, max_iter):
        x = L.new_vector()
        ctl = SolverControl(max_iter, 1e-50, rel_tol)
        SolverCG(ctx, ctl).solve(A, x, b, P)
        return x, ctl.last_step

    def normal_component(d, rhs_d):
        def A(dst, src):
            out = [L.new_vector(), L.new_vector()]
            L.compute_normal_vmult(out, [src, zero] if d == 0 else [zero, src])
            dst.copy_(out[d])
        return solve(A, rhs_d, None, 1e-5, 4000)[0]

    counts = []
    for call in range(2):
        for tau in range(2):
            first = tau == 0
            if first:
                rhs_n = [L.new_vector(), L.new_vector()]
                L.compute_normal_rhs(rhs_n, c)
                normal = [normal_component(d, rhs_n[d]) for d in range(2)]
            rhs = L.new_vector()
            L.reinitialization_rhs(rhs, c, normal, False, first)
            inc, it = solve(lambda dst, src: L.reinitialization_vmult(dst, src, False), rhs,
                            lambda dst, src: torch.mul(src, inv_diag, out=dst), 1e-6, 2000)
            counts.append(it)
            c = c + inc
    assert counts[:2] == gold["initial_reinitialization_cg_iterations"]["value"], counts
    assert counts[2:] == gold["time_step_1_reinitialization_cg_iterations"]["value"], counts
    L.close()


@pytest.mark.parametrize("fixture", ["rising_bubble_ls_picard_output.json", "rising_bubble_ls_output.json"])
def test_device_reproduces_the_advection_of_time_step_2_of_the_rising_bubble_log(ctx, fixture):
    """tests/rising_bubble_ls_picard.output:21 and tests/rising_bubble_ls.output:21 on the device: 'advect [0.000546/10] and
    reinitialize (10 + 10)' (FE_Q_iso_Q1(3)) and 'advect [0.000471/9] and reinitialize (11 + 10)' (FE_Q_iso_Q1(4)).  The fields that enter time step 2 (level
    sets after set-up and after step 1, converged velocity of step 1) come from the oracle chain on the CPU
    (tests/test_2d_cpu.py); local_advance_concentration_rhs, advance_concentration_vmult and the BiCGStab solve with the
    inverse mass diagonal (level_set_okz_advance_concentration.cc:217-258, 300-420, 626-640) run in the library."""
    from adaflo_b200.operators import LevelSetOperators2D
    from adaflo_b200.solvers import SolverBicgstab, SolverControl
    from test_2d_cpu import rising_bubble_first_step
    gold, m, eps_used, c0, c1, u1, _ = rising_bubble_first_step(fixture)
    prm = gold["parameters"]
    dt = prm["step_size"]
    L = LevelSetOperators2D(ctx, m, capi.LSParams(2, prm["concentration_subdivisions"], 0, 0, prm["epsilon"], eps_used, 4.0, prm["h"]),
                            cell_diameters=np.full(m.n_cells, prm["h"]))
    L.set_time(1.5 / dt, -2.0 / dt, 0.5 / dt, dt)                    # BDF-2 weights of the second step
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(ctx.device)
    rhs = L.new_vector()
    # level set extrapolated (two_phase_base.cc:443-456), velocity not (TimeStepping::next, time_stepping.cc:182-193)
    L.advance_concentration_rhs(rhs, dev(u1), dev(2.0 * c1 - c0), dev(c1), dev(c0))
    g2 = gold["time_step_2_concentration_advance"]
    assert f"{float(rhs.norm()):.3g}" == g2["advection_initial_residual"]
    inv_diag = 1.0 / L.diag
    x = L.new_vector()
    ctl = SolverControl(30, 0.05 * 1e-9, 1e-8)
    SolverBicgstab(ctx, ctl).solve(lambda dst, src: L.advance_concentration_vmult(dst, src), x, rhs,
                                   lambda dst, src: torch.mul(src, inv_diag, out=dst))
    assert ctl.last_step == g2["advection_bicgstab_iterations"]
    assert f"{ctl.initial_value:.3g}" == g2["advection_initial_residual"]
    # ... 'and reinitialize (11 + 10)' / '(10 + 10)': the two reinitialisation steps of the advected level set on the device
    from adaflo_b200.solvers import SolverCG
    L.set_time(1.0 / dt, -1.0 / dt, 0.0, dt)    # the reinitialisation uses the step size only
    c = dev(2.0 * c1 - c0) + x
    zero = L.new_vector()

    def solve(A, b, P, rel_tol, max_iter):
        y = L.new_vector()
        control = SolverControl(max_iter, 1e-50, rel_tol)
        SolverCG(ctx, control).solve(A, y, b, P)
        return y, control.last_step

    def normal_component(d, rhs_d):
        def A(dst, src):
            out = [L.new_vector(), L.new_vector()]
            L.compute_normal_vmult(out, [src, zero] if d == 0 else [zero, src])
            dst.copy_(out[d])
        return solve(A, rhs_d, None, 1e-5, 4000)[0]

    counts = []
    for tau in range(2):
        if tau == 0:
            rhs_n = [L.new_vector(), L.new_vector()]
            L.compute_normal_rhs(rhs_n, c)
            normal = [normal_component(d, rhs_n[d]) for d in range(2)]
        r = L.new_vector()
        L.reinitialization_rhs(r, c, normal, False, tau == 0)
        inc, it = solve(lambda dst, src: L.reinitialization_vmult(dst, src, False), r, lambda dst, src: torch.mul(src, inv_diag, out=dst), 1e-6, 2000)
        counts.append(it)
        c = c + inc
    assert counts == g2["reinitialization_cg_iterations"], counts
    L.close()
