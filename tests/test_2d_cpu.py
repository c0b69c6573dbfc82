"""2-D Navier-Stokes operators (CPU): the dense 2-D oracle equals the pinned 3-D oracle on the one-cell-thick
extrusion, A_2d = R A_3d E -- the identity adaflo_b200.operators.NavierStokesMatrix2D is built on."""
import numpy as np
import pytest

from adaflo_b200 import capi
from adaflo_b200.mesh import BoxMesh2D, seeded_uniform
from oracle import oracle as orc
from oracle.ns_oracle_2d import OracleNavierStokes2D

from helpers import flow_params, rel_err, time_weights


def taylor_velocity_2d(x, t=0.0):
    """2-D Taylor flow of tests/beltrami.cc (dim = 2 branch)"""
    u = -np.cos(np.pi * x[:, 0]) * np.sin(np.pi * x[:, 1]) * np.exp(-2 * np.pi ** 2 * t)
    v = np.sin(np.pi * x[:, 0]) * np.cos(np.pi * x[:, 1]) * np.exp(-2 * np.pi ** 2 * t)
    return np.stack([u, v], axis=1)


def fields(m, seed):
    u = taylor_velocity_2d(m.u_coords, 0.01 * seed).reshape(-1) + 0.1 * seeded_uniform(m.u_ids, 2, seed)
    p = seeded_uniform(m.p_ids, 1, seed + 1)
    return u, p


@pytest.mark.parametrize("lin", [capi.NEWTON, capi.PICARD])
def test_2d_oracle_equals_extruded_3d_oracle(lin):
    m = BoxMesh2D((5, 4))
    _, oprm = flow_params(linearization=lin, damping=-0.2, tau_grad_div=0.3, density=1.3, viscosity=0.7)
    _, otw = time_weights()
    O3 = orc.OracleNavierStokesMatrix(m.mesh3d, oprm, otw)
    O2 = OracleNavierStokes2D(m, capi.INCOMPRESSIBLE, lin, 1.3, 0.7, -0.2, 0.3, 0.5, (otw.weight, otw.weight_old, otw.weight_old_old, otw.tau1))
    u, p = fields(m, 3)
    uo, _ = fields(m, 5)
    uoo, _ = fields(m, 7)
    O2.old_u, O2.old_old_u = uo.copy(), uoo.copy()
    O3.old_u, _ = m.extend(uo, np.zeros(m.n_p_nodes))
    O3.old_old_u, _ = m.extend(uoo, np.zeros(m.n_p_nodes))
    # residual (also writes the linearisation point on both sides)
    r2u, r2p = np.zeros(O2.n_u), np.zeros(O2.n_p)
    O2.residual(r2u, r2p, u, p, np.zeros(O2.n_u), np.zeros(O2.n_p))
    u3, p3 = m.extend(u, p)
    r3u, r3p = np.zeros(O3.n_u), np.zeros(O3.n_p)
    O3.residual(r3u, r3p, u3, p3, np.zeros(O3.n_u), np.zeros(O3.n_p))
    ru, rp = m.restrict(r3u, r3p)
    assert rel_err(ru, r2u) < 1e-13 and rel_err(rp, r2p) < 1e-13
    # vmult
    su, sp = fields(m, 11)
    d2u, d2p = O2.vmult(su, sp)
    d3u, d3p = O3.vmult(*m.extend(su, sp))
    du, dp = m.restrict(d3u, d3p)
    c = m.u_dirichlet.reshape(-1)
    du[c] = su[c]                      # constrained rows: one copy, not one per node plane
    assert rel_err(du, d2u) < 1e-13 and rel_err(dp, d2p) < 1e-13
    # velocity block
    v2 = O2.velocity_vmult(su)
    v3 = O3.velocity_vmult(m.extend(su, sp)[0])
    vu, _ = m.restrict(v3, np.zeros(O3.n_p))
    vu[c] = su[c]
    assert rel_err(vu, v2) < 1e-13


def test_2d_stokes_is_symmetric():
    m = BoxMesh2D((4, 4))
    O2 = OracleNavierStokes2D(m, capi.STOKES, capi.NEWTON, 1.0, 0.7, 0.0, 0.0, 0.5, (0.0, 0.0, 0.0, 1.0), pressure_average_fix=False)
    (xu, xp), (yu, yp) = fields(m, 1), fields(m, 2)
    c = m.u_dirichlet.reshape(-1)
    xu[c] = 0.0
    yu[c] = 0.0
    ax, ay = O2.vmult(xu, xp), O2.vmult(yu, yp)
    lhs = ax[0] @ yu + ax[1] @ yp
    rhs = xu @ ay[0] + xp @ ay[1]
    assert abs(lhs - rhs) < 1e-11 * max(abs(lhs), 1.0)


def test_2d_oracle_reproduces_the_first_residual_of_the_couette_log():
    """tests/couette.output:10 -- first nonlinear residual 1.601e+01 of time step #1 (tests/couette.cc:113-145: 64 x 16
    cells on [-2,2] x [-1,0], no-slip at the bottom, lid velocity (2, 0) at the top, open boundaries with zero external
    pressure left and right: there the tangential velocity is constrained and no face term arises).  The solution is zero
    (ZeroFunction start) except for the lid values of apply_boundary_conditions; step #1 of BDF-2 uses implicit Euler
    weights (time_stepping.cc:158-198).  A direct pin of the dense 2-D restatement against a 2-D log of the reference."""
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "couette_output.json")) as f:
        gold = json.load(f)
    prm = gold["parameters"]
    m = BoxMesh2D((64, 16), lower=(-2.0, -1.0), upper=(2.0, 0.0))
    assert m.n_cells == gold["n_active_cells"]["value"] and float(m.h[0]) == gold["mesh_size"]["value"]
    assert (2 * m.n_u_nodes, m.n_p_nodes) == (gold["n_dofs"]["velocity"], gold["n_dofs"]["pressure"])
    x = m.u_coords
    bottom, top = np.abs(x[:, 1] + 1) < 1e-12, np.abs(x[:, 1]) < 1e-12
    sides = (np.abs(x[:, 0] - 2) < 1e-12) | (np.abs(x[:, 0] + 2) < 1e-12)
    dirichlet = np.zeros_like(m.u_dirichlet)
    dirichlet[bottom | top, :] = True
    dirichlet[sides, 1] = True                    # set_open_boundary_with_normal_flux: tangential component constrained
    m.u_dirichlet = dirichlet
    dt = prm["dt"]
    O = OracleNavierStokes2D(m, capi.INCOMPRESSIBLE, capi.NEWTON, prm["density"], prm["viscosity"], 0.0, 0.0, 0.5,
                             (1 / dt, -1 / dt, 0.0, 1.0), pressure_average_fix=False)
    u = np.zeros(O.n_u)
    u.reshape(-1, 2)[top] = prm["lid_velocity"]
    ru, rp = np.zeros(O.n_u), np.zeros(O.n_p)
    O.residual(ru, rp, u, np.zeros(O.n_p), np.zeros(O.n_u), np.zeros(O.n_p))
    assert f"{np.sqrt(ru @ ru + rp @ rp):.3e}" == gold["first_nonlinear_residual"]["value"]
    assert np.linalg.norm(rp) < 1e-14               # the lid field is divergence free


def test_2d_oracle_reproduces_the_second_time_step_of_the_couette_log():
    """tests/couette.output:10,31 -- the Newton increment 5.83e+01 of time step #1, and the first nonlinear residual 1.930e-01 and
    increment 1.60e+01 of time step #2, which need the converged solution of step #1: Newton on the oracle's residual / vmult with
    an exactly recovered Jacobian (open boundaries fix the pressure level: nothing is pinned).  Step #2 uses BDF-2 weights,
    does not extrapolate (time_stepping.cc:182-193), and its solution_old_old is the ZERO start vector -- the lid values enter
    'solution' only (with the lid field there the residual would be 1.919e-01)."""
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "couette_output.json")) as f:
        gold = json.load(f)
    prm = gold["parameters"]
    m = BoxMesh2D((64, 16), lower=(-2.0, -1.0), upper=(2.0, 0.0))
    x = m.u_coords
    bottom, top = np.abs(x[:, 1] + 1) < 1e-12, np.abs(x[:, 1]) < 1e-12
    sides = (np.abs(x[:, 0] - 2) < 1e-12) | (np.abs(x[:, 0] + 2) < 1e-12)
    dirichlet = np.zeros_like(m.u_dirichlet)
    dirichlet[bottom | top, :] = True
    dirichlet[sides, 1] = True
    m.u_dirichlet = dirichlet
    dt = prm["dt"]

    def operator(tw):
        return OracleNavierStokes2D(m, capi.INCOMPRESSIBLE, capi.NEWTON, prm["density"], prm["viscosity"], 0.0, 0.0, 0.5, tw, pressure_average_fix=False)
    O = operator((1 / dt, -1 / dt, 0.0, 1.0))
    u0 = np.zeros(O.n_u)
    u0.reshape(-1, 2)[top] = prm["lid_velocity"]
    zero_u, zero_p = np.zeros(O.n_u), np.zeros(O.n_p)
    increments = []
    u1, p1, history = newton_with_probed_jacobian(O, m, zero_u, tol=1e-11, start=(u0, zero_p), pin_pressure=False, increments=increments)
    assert f"{history[0]:.3e}" == gold["first_nonlinear_residual"]["value"] and history[-1] < 1e-11
    assert f"{increments[0]:.2e}" == gold["time_step_1_first_newton_increment"]["value"]
    O2 = operator((1.5 / dt, -2.0 / dt, 0.5 / dt, 1.0))
    O2.old_u, O2.old_old_u = u1.copy(), zero_u.copy()
    increments = []
    _, _, history = newton_with_probed_jacobian(O2, m, zero_u, tol=1e-11, start=(u1, p1), pin_pressure=False, increments=increments)
    assert f"{history[0]:.3e}" == gold["time_step_2_first_nonlinear_residual"]["value"] and history[-1] < 1e-11
    assert f"{increments[0]:.2e}" == gold["time_step_2_first_nonlinear_residual"]["increment"]


# ---------------------------------------------------------------------------------------------- level set
def ls_setup(n=(4, 3), n_sub=3, dirichlet_edges=()):
    from oracle.ls_oracle_2d import OracleLevelSet2D
    m = BoxMesh2D(n)
    m.build_ls_space(n_sub, dirichlet_edges)
    h_min = float(m.h.min())
    eps, eps_used, step = 1.5, 1.5 * h_min / n_sub, 0.02
    cd = m.cell_diameters()
    O2 = OracleLevelSet2D(m, n_sub, eps, eps_used, h_min, step, cd)
    O3 = orc.OracleLevelSet(m.mesh3d, orc.LSParams(eps, eps_used, 4.0, h_min, step, 2, 0))
    O3.cd[:] = cd                                    # the 2-D cell diameters, not those of the extruded hexahedra
    d3 = O3.mass_diagonal()
    O3.diag = d3
    O2.diag = O2.mass_diagonal()
    return m, O2, O3, (eps, eps_used, h_min, step)


def ls_fields(m):
    x = m.ls_coords
    r = np.linalg.norm(x - np.array([0.1, -0.05]), axis=1)
    c = -np.tanh((r - 0.45) / 0.3)
    vec = lambda s: seeded_uniform(m.ls_ids, 1, s)
    return c, vec


@pytest.mark.parametrize("dirichlet_edges", [(), (0,)])
def test_2d_level_set_oracle_equals_extruded_3d_oracle(dirichlet_edges):
    m, O2, O3, _ = ls_setup(dirichlet_edges=dirichlet_edges)
    c, vec = ls_fields(m)
    E, R = m.extend_ls, m.restrict_ls
    con = m.ls_dirichlet
    # mass diagonal: sum_a int N_a^2 dz = 2/3 for the hats with the composite Gauss rule
    assert rel_err(1.5 * R(O3.diag), O2.diag) < 1e-13
    # advection rhs (evaluates the Q2 velocity) and vmult
    vel = taylor_velocity_2d(m.u_coords, 0.0).reshape(-1) + 0.1 * seeded_uniform(m.u_ids, 2, 6)
    c_old, c_oo = c + 0.05 * vec(7), c + 0.05 * vec(8)
    r2 = vec(9)
    r3 = E(np.zeros_like(r2))
    pre = r2.copy()
    O2.advance_concentration_rhs(r2, vel, c, c_old, c_oo, 75.0, -100.0, 25.0)
    O3.advance_concentration_rhs(r3, m.extend(vel, np.zeros(m.n_p_nodes))[0], E(c), E(c_old), E(c_oo), 75.0, -100.0, 25.0)
    assert rel_err(R(r3) + pre, r2) < 1e-13
    src = vec(10)

    def fix(d, s):                                   # constrained rows: diag * src once, not per node plane
        d = d.copy()
        d[con] = O2.diag[con] * s[con]
        return d
    assert rel_err(fix(R(O3.advance_concentration_vmult(E(src), 75.0)), src), O2.advance_concentration_vmult(src, 75.0)) < 1e-13
    # normal projection
    n2 = [np.zeros(O2.n) for _ in range(2)]
    n3 = [np.zeros(O3.n) for _ in range(3)]
    O2.compute_normal_rhs(n2, c)
    O3.compute_normal_rhs(n3, E(c))
    assert rel_err(R(n3[0]), n2[0]) < 1e-13 and rel_err(R(n3[1]), n2[1]) < 1e-13 and np.abs(n3[2]).max() < 1e-13
    normal = [m.ls_coords[:, d] - [0.1, -0.05][d] + 0.02 * vec(20 + d) for d in range(2)]
    v3 = O3.compute_normal_vmult([E(normal[0]), E(normal[1]), np.zeros(O3.n)])
    v2 = O2.compute_normal_vmult(normal, 4.0)
    for d in range(2):
        assert rel_err(fix(R(v3[d]), normal[d]), v2[d]) < 1e-13
    # curvature
    k2, k3 = np.zeros(O2.n), np.zeros(O3.n)
    O2.compute_curvature_rhs(k2, normal)
    O3.compute_curvature_rhs(k3, [E(normal[0]), E(normal[1]), np.zeros(O3.n)])
    assert rel_err(R(k3), k2) < 1e-13
    assert rel_err(fix(R(O3.compute_curvature_vmult(E(src), True)), src), O2.compute_curvature_vmult(src, True)) < 1e-13
    # reinitialisation
    q2, q3 = np.zeros(O2.n), np.zeros(O3.n)
    O2.reinitialization_rhs(q2, c, normal, False, True)
    O3.reinitialization_rhs(q3, E(c), [E(normal[0]), E(normal[1]), np.zeros(O3.n)], False, True)
    assert rel_err(R(q3), q2) < 1e-13
    for diffuse_only in (False, True):
        assert rel_err(fix(R(O3.reinitialization_vmult(E(src), diffuse_only)), src), O2.reinitialization_vmult(src, diffuse_only)) < 1e-13


def test_2d_force_oracle_equals_extruded_3d_oracle_without_gravity():
    """surface-tension part of local_compute_force (the 3-D oracle applies gravity in z, so gravity = 0 here; the
    gravity term of the 2-D path is checked on the device against the 2-D oracle)"""
    from oracle.ls_oracle_2d import compute_force_2d
    m, O2, O3, (eps, _, _, _) = ls_setup()
    c, vec = ls_fields(m)
    H = orc.compute_heaviside(m.cell_ls, c, eps, m.n_sub)
    kap = 2.0 / np.maximum(np.linalg.norm(m.ls_coords, axis=1), 0.2) + 0.1 * vec(3)
    f2 = np.zeros(m.n_u_nodes * 2)
    rho2, mu2 = compute_force_2d(m, m.n_sub, H, kap, 0.07, 0.0, 1.0, 9.0, 0.1, 0.4, True, f2)
    f3 = np.zeros(m.mesh3d.u.n_owned * 3)
    rho3, mu3 = orc.compute_force(m.mesh3d, m.extend_ls(H), m.extend_ls(kap), 0.07, 0.0, 1.0, 9.0, 0.1, 0.4, True, None, f3)
    fu, _ = m.restrict(f3, np.zeros(m.mesh3d.p.n_owned))
    assert rel_err(fu, f2) < 1e-13
    assert rel_err(rho3.reshape(-1, 3, 9), np.repeat(rho2[:, None, :], 3, axis=1)) < 1e-13
    assert rel_err(mu3.reshape(-1, 3, 9), np.repeat(mu2[:, None, :], 3, axis=1)) < 1e-13


# --------------------------------------------------------------------- golden log of the 2-D rising bubble (level set)
def rising_bubble_setup(fixture="rising_bubble_ls_output.json"):
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", fixture)) as f:
        gold = json.load(f)
    prm = gold["parameters"]
    m = BoxMesh2D((40, 80), lower=(0.0, 0.0), upper=(1.0, 2.0))          # tests/rising_bubble.cc:122-131 + 3 refinements
    m.build_ls_space(prm["concentration_subdivisions"])
    eps_used = prm["epsilon"] / prm["concentration_subdivisions"] * prm["h"]   # two_phase_base.cc:290-291
    dist = np.linalg.norm(m.ls_coords - np.array(prm["bubble_centre"]), axis=1) - prm["bubble_radius"]
    c0 = -np.tanh(dist / (2.0 * eps_used))                                 # transform_distance_function, level_set_okz.cc:201-210
    return gold, m, eps_used, c0


def cg_iterations(A, b, P, rel_tol, max_iter=4000):
    """deal.II SolverCG with ReductionControl(max_iter, 1e-50, rel_tol), zero start vector: solution, last_step()"""
    x, r = np.zeros_like(b), b.copy()
    r0 = np.linalg.norm(r)
    z = P(r) if P else r.copy()
    p, rz, it = z.copy(), r @ z, 0
    while np.linalg.norm(r) > rel_tol * r0 and it < max_iter:
        it += 1
        Ap = A(p)
        alpha = rz / (p @ Ap)
        x += alpha * p
        r -= alpha * Ap
        z = P(r) if P else r.copy()
        rz_new = r @ z
        p = z + (rz_new / rz) * p
        rz = rz_new
    return x, it


def assembled_projection(O):
    """The projection operator mass + 4 delta^2 Laplace (O._scalar(0, 4.0, .)) as a sparse matrix, recovered exactly from 9
    coloured applications (a node of the FE_Q_iso_Q1 grid couples only with its 8 neighbours): the Krylov solves of the normal
    and curvature projections apply it a few hundred times per call.  Cached on the oracle object."""
    if getattr(O, "_projection_matrix", None) is None:
        import scipy.sparse as sp
        m = O.m
        ij = np.rint((m.ls_coords - m.ls_coords.min(axis=0)) / (np.asarray(m.h) / O.n_sub)).astype(int)
        nx, ny = ij[:, 0].max() + 1, ij[:, 1].max() + 1
        node_of = -np.ones((nx, ny), dtype=int)
        node_of[ij[:, 0], ij[:, 1]] = np.arange(O.n)
        rows, cols, vals = [], [], []
        for gi in range(3):
            for gj in range(3):
                e = ((ij[:, 0] % 3 == gi) & (ij[:, 1] % 3 == gj)).astype(float)
                y = O._scalar(0, 4.0, e)
                nz = np.nonzero(y)[0]
                ci = ij[nz, 0] + ((gi - ij[nz, 0] + 1) % 3) - 1
                cj = ij[nz, 1] + ((gj - ij[nz, 1] + 1) % 3) - 1
                rows.append(nz)
                cols.append(node_of[ci, cj])
                vals.append(y[nz])
        O._projection_matrix = sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(O.n, O.n))
        probe = np.cos(np.arange(O.n) * 0.37)
        assert np.linalg.norm(O._projection_matrix @ probe - O._scalar(0, 4.0, probe)) < 1e-12 * np.linalg.norm(probe)
    return lambda v: O._projection_matrix @ v


def corrected_curvature(O, c, eps_used):
    """LevelSetOKZSolverComputeCurvature::compute_curvature (level_set_okz_compute_curvature.cc:326-378) for dim = 2:
    normal with tolerance 1e-7 (compute_normal(false)), right-hand side, projection with tolerance 1e-8, correction by the
    distance.  Both projections are solved with the ASSEMBLED projection matrix of
    level_set_okz_preconditioner.cc:61-180, mass + 4 delta^2 Laplace -- i.e. compute_normal_vmult's operator with
    damping_scale_factor 4 -- not with compute_curvature_vmult (1 delta^2); the surface-tension dominated log of
    spurious_currents_ls tells the two apart (0.365 vs 0.363)."""
    rhs_n = [np.zeros(O.n), np.zeros(O.n)]
    O.compute_normal_rhs(rhs_n, c)
    project = assembled_projection(O)
    normal = [cg_iterations(project, rhs_n[d], None, 1e-7)[0] for d in range(2)]
    rk = np.zeros(O.n)
    O.compute_curvature_rhs(rk, normal)
    kappa, _ = cg_iterations(project, rk, None, 1e-8)
    big = kappa > 1e-4
    with np.errstate(divide="ignore", invalid="ignore"):
        distance = np.where((1 - c * c) > 1e-2, eps_used * np.log((1 + c) / (1 - c)), 0.0)
    kappa[big] = 1.0 / (1.0 / kappa[big] + distance[big])                     # distance / (dim - 1), dim = 2
    return kappa


def reinitialize(O, c, n_steps):
    """LevelSetOKZSolverReinitialization::reinitialize (reinitialization.cc:256-369) with n_steps stabilised steps:
    returns the new level set and the CG iteration count of every step"""
    counts = []
    for tau in range(n_steps):
        first = tau == 0
        if first:                         # compute_normal(true), compute_normal.cc:208-268
            rhs_n = [np.zeros(O.n), np.zeros(O.n)]
            O.compute_normal_rhs(rhs_n, c)
            normal = [cg_iterations(assembled_projection(O), rhs_n[d], None, 1e-5)[0] for d in range(2)]
        rhs = np.zeros(O.n)
        O.reinitialization_rhs(rhs, c, normal, False, first)
        inc, it = cg_iterations(lambda v: O.reinitialization_vmult(v, False), rhs, lambda r: r / O.diag, 1e-6, 2000)
        counts.append(it)
        c = c + inc
    return c, counts


@pytest.mark.parametrize("fixture", ["rising_bubble_ls_output.json", "rising_bubble_ls_picard_output.json"])
def test_level_set_oracle_reproduces_the_reinitialisation_counts_of_the_rising_bubble_log(fixture):
    """tests/rising_bubble_ls.output:6,13 -- 'reinitialize (8 + 8)' and 'reinitialize (7 + 7)' with FE_Q_iso_Q1(4) -- and
    tests/rising_bubble_ls_picard.output:6,13 -- '(8 + 7)' and '(7 + 7)' with FE_Q_iso_Q1(3): the CG iteration counts
    of LevelSetOKZSolverReinitialization::reinitialize (reinitialization.cc:256-369: normal projection, right-hand side
    with the frozen normals, reinitialization_vmult, inverse mass diagonal, ReductionControl 1e-6).  This pins the
    level-set oracle (normal / reinitialisation operators and right-hand sides, mass diagonal) against the reference's
    own output; the 3-D oracle oracle/ls_oracle.c is tied to this 2-D restatement by the extrusion identity above."""
    from oracle.ls_oracle_2d import OracleLevelSet2D
    gold, m, eps_used, c = rising_bubble_setup(fixture)
    prm = gold["parameters"]
    assert m.n_cells == gold["n_active_cells"]["value"]
    assert (2 * m.n_u_nodes, m.n_p_nodes, m.n_ls_nodes) == (gold["n_dofs"]["velocity"], gold["n_dofs"]["pressure"], gold["n_dofs"]["level_set"])
    O = OracleLevelSet2D(m, prm["concentration_subdivisions"], prm["epsilon"], eps_used, prm["h"], prm["step_size"], np.full(m.n_cells, prm["h"]))
    O.diag = O.mass_diagonal()
    counts = []
    for call in range(2):                     # setup_problem (level_set_base.cc:152), then time step #1 (zero velocity)
        c, its = reinitialize(O, c, 2)
        counts += its
    assert counts[:2] == gold["initial_reinitialization_cg_iterations"]["value"]
    assert counts[2:] == gold["time_step_1_reinitialization_cg_iterations"]["value"]
    assert abs(c.min() + 1) < 1e-6 and abs(c.max() - 1) < 1e-6              # "Range of level set values: -1 / 1"

    # ---- first Navier-Stokes residual of time step #1 (line 14: 0.0198).  The velocity and pressure are still zero, so
    # the residual is the force vector: curvature, Heaviside function, compute_force with the level-set -> pressure
    # interpolation ("grad pressure compatible = 1"), variable density, gravity on y; no-slip on bottom / top, symmetry
    # on the sides
    from oracle.ls_oracle_2d import compute_force_2d
    n_sub = prm["concentration_subdivisions"]
    kappa = corrected_curvature(O, c, eps_used)
    H = orc.compute_heaviside(m.cell_ls, c, prm["epsilon"], n_sub)
    x = m.u_coords
    dirichlet = np.zeros_like(m.u_dirichlet)
    dirichlet[(np.abs(x[:, 1]) < 1e-12) | (np.abs(x[:, 1] - 2) < 1e-12), :] = True      # set_no_slip_boundary(0)
    dirichlet[(np.abs(x[:, 0]) < 1e-12) | (np.abs(x[:, 0] - 1) < 1e-12), 0] = True      # set_symmetry_boundary(2)
    f = np.zeros(m.n_u_nodes * 2)
    compute_force_2d(m, n_sub, H, kappa, prm["surface_tension"], prm["gravity"], prm["density"], prm["density_difference"],
                     prm["viscosity"], prm["viscosity_difference"], True, f, interpolate=True, u_dirichlet=dirichlet)
    assert f"{np.linalg.norm(f):.3g}" == gold["time_step_1_first_navier_stokes_residual"]["value"]


def test_level_set_oracle_reproduces_the_first_step_of_the_spurious_currents_log():
    """tests/spurious_currents_ls.output:9,10 -- 'advect [0/0] and reinitialize (8 + 7)' and the first Navier-Stokes
    residual 0.365 of time step #1 (tests/spurious_currents.cc:263-270: 80 x 80 cells on [-2.5, 2.5]^2, no-slip everywhere,
    FE_Q_iso_Q1(3), surface tension 1, no gravity, constant density).  Velocity and pressure are zero, so the residual is
    the norm of sigma kappa grad H: unlike the gravity-dominated rising bubble this pins the curvature itself (normal
    projection, normalised divergence, projection with the 4 delta^2 matrix, distance correction) to three digits."""
    import json
    import os
    from oracle.ls_oracle_2d import OracleLevelSet2D, compute_force_2d
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "spurious_currents_ls_output.json")) as f:
        gold = json.load(f)
    prm = gold["parameters"]
    n_sub, h = prm["concentration_subdivisions"], prm["h"]
    m = BoxMesh2D((80, 80), lower=(-2.5, -2.5), upper=(2.5, 2.5))
    m.build_ls_space(n_sub)
    assert m.n_cells == gold["n_active_cells"]["value"] and float(m.h[0]) == gold["mesh_size"]["value"]
    assert (2 * m.n_u_nodes, m.n_p_nodes, m.n_ls_nodes) == (gold["n_dofs"]["velocity"], gold["n_dofs"]["pressure"], gold["n_dofs"]["level_set"])
    eps_used = prm["epsilon"] / n_sub * h
    dist = np.linalg.norm(m.ls_coords - np.array(prm["bubble_centre"]), axis=1) - prm["bubble_radius"]
    c = -np.tanh(dist / (2.0 * eps_used))
    O = OracleLevelSet2D(m, n_sub, prm["epsilon"], eps_used, h, prm["step_size"], np.full(m.n_cells, h))
    O.diag = O.mass_diagonal()
    c, counts = reinitialize(O, c, 2)          # 'number initial reinit steps = 0': the first reinitialisation is in step #1
    assert counts == gold["time_step_1_reinitialization_cg_iterations"]["value"]
    kappa = corrected_curvature(O, c, eps_used)
    inside = dist < -0.2
    assert abs(np.median(kappa[np.abs(dist) < 0.02]) - 2.0) < 0.05          # circle of radius 0.5: kappa = (dim - 1) / r
    H = orc.compute_heaviside(m.cell_ls, c, prm["epsilon"], n_sub)
    assert H[inside].min() == 1.0 and H[dist > 0.2].max() == 0.0
    f = np.zeros(m.n_u_nodes * 2)
    compute_force_2d(m, n_sub, H, kappa, prm["surface_tension"], prm["gravity"], prm["density"], prm["density_difference"],
                     prm["viscosity"], prm["viscosity_difference"], True, f, interpolate=True)
    assert f"{np.linalg.norm(f):.3g}" == gold["time_step_1_first_navier_stokes_residual"]["value"]


# --------------------------------------------------------------------- second time step of the rising-bubble log
def bicgstab_iterations(A, b, P, max_iter, abs_tol, rel_tol):
    """deal.II SolverBicgstab (exact_residual = false) with ReductionControl(max_iter, abs_tol, rel_tol), zero start vector,
    right-preconditioned directions as in deal.II: solution, last_step(), initial_value()"""
    x, r = np.zeros_like(b), b.copy()
    rbar = r.copy()
    res0 = res = np.linalg.norm(r)
    tol = max(abs_tol, rel_tol * res)
    rho = alpha = omega = 1.0
    p = v = None
    it = 0
    while res > tol and it < max_iter:
        it += 1
        rhobar = r @ rbar
        beta = rhobar * alpha / (rho * omega)
        rho = rhobar
        p = r.copy() if p is None else r + beta * p - beta * omega * v
        y = P(p)
        v = A(y)
        alpha = rho / (rbar @ v)
        r = r - alpha * v
        res = np.linalg.norm(r)
        if res <= tol:
            x += alpha * y
            break
        z = P(r)
        t = A(z)
        omega = (t @ r) / (t @ t)
        x += alpha * y + omega * z
        r = r - omega * t
        res = np.linalg.norm(r)
    return x, it, res0


def newton_with_probed_jacobian(A, m, f, tol=1e-9, start=None, pin_pressure=True, increments=None):
    """Solve the nonlinear Navier-Stokes system of one time step with the oracle operator A (residual / vmult): the Jacobian
    is recovered exactly from 75 coloured applications of vmult (two nodes of the Q2 grid couple only within +-2 grid lines, so
    the columns with equal coordinates mod 5 are probed together) and factorised; one pressure dof is pinned (closed
    container).  Returns (u, p, residual norms)."""
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla
    nu, npr = A.n_u, A.n_p
    h = np.asarray(m.h)
    iu = np.rint((m.u_coords - m.u_coords.min(axis=0)) / (0.5 * h)).astype(int)
    ip = np.rint((m.p_coords - m.u_coords.min(axis=0)) / (0.5 * h)).astype(int)
    nx, ny = iu[:, 0].max() + 1, iu[:, 1].max() + 1
    u_of = -np.ones((nx, ny), dtype=int)
    u_of[iu[:, 0], iu[:, 1]] = np.arange(len(iu))
    p_of = -np.ones((nx, ny), dtype=int)
    p_of[ip[:, 0], ip[:, 1]] = np.arange(len(ip))
    row_ij = np.concatenate([np.repeat(iu, 2, axis=0), ip])

    def jacobian():
        rows, cols, vals = [], [], []
        for kind in range(3):                                       # columns: u_x, u_y, p
            ij = iu if kind < 2 else ip
            for gi in range(5):
                for gj in range(5):
                    sel = np.nonzero((ij[:, 0] % 5 == gi) & (ij[:, 1] % 5 == gj))[0]
                    if len(sel) == 0:
                        continue
                    eu, ep = np.zeros(nu), np.zeros(npr)
                    if kind < 2:
                        eu[2 * sel + kind] = 1.0
                    else:
                        ep[sel] = 1.0
                    y = np.concatenate(A.vmult(eu, ep))
                    nz = np.nonzero(y)[0]
                    ci = row_ij[nz, 0] + ((gi - row_ij[nz, 0] + 2) % 5) - 2
                    cj = row_ij[nz, 1] + ((gj - row_ij[nz, 1] + 2) % 5) - 2
                    node = (u_of if kind < 2 else p_of)[ci, cj]
                    assert (node >= 0).all()
                    rows.append(nz)
                    cols.append(2 * node + kind if kind < 2 else nu + node)
                    vals.append(y[nz])
        return sp.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(nu + npr, nu + npr))

    u, p, history = (np.zeros(nu), np.zeros(npr), []) if start is None else (start[0].copy(), start[1].copy(), [])
    lu = None
    for _ in range(12):
        ru, rp = np.zeros(nu), np.zeros(npr)
        A.residual(ru, rp, u, p, f, np.zeros(npr))          # also moves the linearisation point of vmult to (u, p)
        history.append(float(np.sqrt(ru @ ru + rp @ rp)))
        if history[-1] < tol:
            break
        if lu is None:    # the first update is an exact Newton step; later ones reuse its factorisation (the velocities are
            J = jacobian().tocoo()    # small against the viscous and mass terms: every further step gains two to three digits)
            if pin_pressure:                                 # closed container: one pressure dof pinned (identity row)
                keep = J.row != nu
                J = sp.csc_matrix((np.append(J.data[keep], 1.0), (np.append(J.row[keep], nu), np.append(J.col[keep], nu))), shape=J.shape)
            lu = spla.splu(sp.csc_matrix(J))
        b = np.concatenate([ru, rp])
        if pin_pressure:
            b[nu] = 0.0
        d = lu.solve(b)
        if increments is not None:
            increments.append(float(np.linalg.norm(d)))
        u += d[:nu]
        p += d[nu:]
    return u, p, history


@pytest.mark.parametrize("fixture", ["rising_bubble_ls_output.json", "rising_bubble_ls_picard_output.json"])
def test_oracle_chain_reproduces_the_second_and_third_time_step_of_the_rising_bubble_log(fixture):
    """tests/rising_bubble_ls.output:21-22 -- 'Concentration advance: advect [0.000471/9] and reinitialize (11 + 10)' and
    'Residual/iterations: [0.00581/...' -- and :29-30 -- 'advect [0.00108/10] and reinitialize (11 + 11)', '[0.000246/...': the
    first quantities of the log that depend on a converged two-phase Navier-Stokes solution.  Chain: time step 1 (reinitialisation, curvature, Heaviside, force, variable density / viscosity, Newton on the
    oracle's residual / vmult solved to 1e-9 with an exactly recovered Jacobian); time step 2: level set extrapolated
    (two_phase_base.cc:443-456), velocity not (TimeStepping::next, time_stepping.cc:182-193), BDF-2 weights,
    local_advance_concentration_rhs, BiCGStab on advance_concentration_vmult with the inverse mass diagonal
    (level_set_okz_advance_concentration.cc:217-258, 300-420, 626-640), reinitialisation, force, residual.  This pins the
    advection operator and right-hand side, which no other log reaches, and the variable-coefficient
    Navier-Stokes operator incl. its Jacobian.  Second fixture: tests/rising_bubble_ls_picard.output (FE_Q_iso_Q1(3); its Picard
    iteration converges to the same nonlinear solution, so the same chain applies): advect [0.000546/10], (10 + 10), 0.0058;
    advect [0.00125/11], (11 + 11), 0.000244."""
    from oracle.ls_oracle_2d import OracleLevelSet2D, compute_force_2d
    from oracle.ns_oracle_2d import OracleNavierStokes2DVariable
    gold, m, eps_used, c = rising_bubble_setup(fixture)
    prm = gold["parameters"]
    n_sub, dt = prm["concentration_subdivisions"], prm["step_size"]
    O = OracleLevelSet2D(m, n_sub, prm["epsilon"], eps_used, prm["h"], dt, np.full(m.n_cells, prm["h"]))
    O.diag = O.mass_diagonal()
    x = m.u_coords
    dirichlet = np.zeros_like(m.u_dirichlet)
    dirichlet[(np.abs(x[:, 1]) < 1e-12) | (np.abs(x[:, 1] - 2) < 1e-12), :] = True      # set_no_slip_boundary(0)
    dirichlet[(np.abs(x[:, 0]) < 1e-12) | (np.abs(x[:, 0] - 1) < 1e-12), 0] = True      # set_symmetry_boundary(2)
    m.u_dirichlet = dirichlet

    def force(level_set):
        kappa = corrected_curvature(O, level_set, eps_used)
        H = orc.compute_heaviside(m.cell_ls, level_set, prm["epsilon"], n_sub)
        f = np.zeros(m.n_u_nodes * 2)
        rho_q, mu_q = compute_force_2d(m, n_sub, H, kappa, prm["surface_tension"], prm["gravity"], prm["density"], prm["density_difference"],
                                       prm["viscosity"], prm["viscosity_difference"], True, f, interpolate=True, u_dirichlet=dirichlet)
        return f, rho_q, mu_q

    # ---- set-up and time step 1 (zero velocity: 'advect [0/0]')
    c, _ = reinitialize(O, c, 2)
    c_old = c.copy()
    c1, counts = reinitialize(O, c, 2)
    assert counts == gold["time_step_1_reinitialization_cg_iterations"]["value"]
    f, rho_q, mu_q = force(c1)
    A = OracleNavierStokes2DVariable(m, rho_q, mu_q, 0.0, 0.0, 0.5, (1.0 / dt, -1.0 / dt, 0.0, 1.0))   # BDF-2 start-up step
    u1, p1, history = newton_with_probed_jacobian(A, m, f)
    assert f"{history[0]:.3g}" == gold["time_step_1_first_navier_stokes_residual"]["value"]
    assert history[1] < 1e-5 and history[-1] < 1e-9       # the log: [0.0198] [4.61e-06] [3.05e-10/conv.] (inexact solves)

    # ---- time step 2
    g2 = gold["time_step_2_concentration_advance"]
    c_star = 2.0 * c1 - c_old
    rhs = np.zeros(O.n)
    O.advance_concentration_rhs(rhs, u1, c_star, c1, c_old, 1.5 / dt, -2.0 / dt, 0.5 / dt)
    inc, n_it, res0 = bicgstab_iterations(lambda v: O.advance_concentration_vmult(v, 1.5 / dt), rhs, lambda r: r / O.diag, 30, 0.05 * 1e-9, 1e-8)
    assert f"{res0:.3g}" == g2["advection_initial_residual"]
    assert n_it == g2["advection_bicgstab_iterations"]
    c2, counts = reinitialize(O, c_star + inc, 2)
    assert counts == g2["reinitialization_cg_iterations"]
    f2, rho_q, mu_q = force(c2)
    A2 = OracleNavierStokes2DVariable(m, rho_q, mu_q, 0.0, 0.0, 0.5, (1.5 / dt, -2.0 / dt, 0.5 / dt, 1.0))
    A2.old_u, A2.old_old_u = u1.copy(), np.zeros(A2.n_u)
    ru, rp = np.zeros(A2.n_u), np.zeros(A2.n_p)
    A2.residual(ru, rp, u1, p1, f2, np.zeros(A2.n_p))
    assert f"{np.sqrt(ru @ ru + rp @ rp):.3g}" == gold["time_step_2_first_navier_stokes_residual"]["value"]

    # ---- time step 3 (lines 29-30): now the velocity and the pressure are extrapolated too (step_no > 1)
    g3 = gold["time_step_3_concentration_advance"]
    u2, p2, history = newton_with_probed_jacobian(A2, m, f2, start=(u1, p1))
    assert history[-1] < 1e-9
    if "time_step_2_second_newton_residual" in gold:   # [4.08e-06/14] after the log's first (inexact, 1e-4) linear solve: same to 3 digits
        assert f"{history[1]:.3g}" == gold["time_step_2_second_newton_residual"]["value"]
    u_star, p_star, c_star = 2.0 * u2 - u1, 2.0 * p2 - p1, 2.0 * c2 - c1
    rhs = np.zeros(O.n)
    O.advance_concentration_rhs(rhs, u_star, c_star, c2, c1, 1.5 / dt, -2.0 / dt, 0.5 / dt)
    inc, n_it, res0 = bicgstab_iterations(lambda v: O.advance_concentration_vmult(v, 1.5 / dt), rhs, lambda r: r / O.diag, 30, 0.05 * 1e-9, 1e-8)
    assert f"{res0:.3g}" == g3["advection_initial_residual"]
    assert n_it == g3["advection_bicgstab_iterations"]
    c3, counts = reinitialize(O, c_star + inc, 2)
    assert counts == g3["reinitialization_cg_iterations"]
    f3, rho_q, mu_q = force(c3)
    A3 = OracleNavierStokes2DVariable(m, rho_q, mu_q, 0.0, 0.0, 0.5, (1.5 / dt, -2.0 / dt, 0.5 / dt, 1.0))
    A3.old_u, A3.old_old_u = u2.copy(), u1.copy()
    ru, rp = np.zeros(A3.n_u), np.zeros(A3.n_p)
    A3.residual(ru, rp, u_star, p_star, f3, np.zeros(A3.n_p))
    assert f"{np.sqrt(ru @ ru + rp @ rp):.3g}" == gold["time_step_3_first_navier_stokes_residual"]["value"]


def test_oracle_chain_reproduces_the_second_time_step_of_the_spurious_currents_log():
    """tests/spurious_currents_ls.output:15-16 -- 'advect [0.000803/8] and reinitialize (7 + 7)', '[0.00024/...': the same chain
    as for the rising bubble on the surface-tension dominated problem (80 x 80 cells, no-slip everywhere, constant density,
    no initial reinitialisation): the spurious velocity of step 1 advects the level set in step 2.  The size of that velocity,
    printed with 7 digits (3.776641e-05, line 12), is reproduced to 2.5e-7: a pin of the converged velocity field itself."""
    import json
    import os
    from oracle.ls_oracle_2d import OracleLevelSet2D, compute_force_2d
    from oracle.ns_oracle_2d import OracleNavierStokes2DVariable
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "spurious_currents_ls_output.json")) as f:
        gold = json.load(f)
    prm = gold["parameters"]
    n_sub, h, dt = prm["concentration_subdivisions"], prm["h"], prm["step_size"]
    m = BoxMesh2D((80, 80), lower=(-2.5, -2.5), upper=(2.5, 2.5))
    m.build_ls_space(n_sub)
    eps_used = prm["epsilon"] / n_sub * h
    dist = np.linalg.norm(m.ls_coords - np.array(prm["bubble_centre"]), axis=1) - prm["bubble_radius"]
    c_old = -np.tanh(dist / (2.0 * eps_used))
    O = OracleLevelSet2D(m, n_sub, prm["epsilon"], eps_used, h, dt, np.full(m.n_cells, h))
    O.diag = O.mass_diagonal()

    def force(level_set):
        kappa = corrected_curvature(O, level_set, eps_used)
        H = orc.compute_heaviside(m.cell_ls, level_set, prm["epsilon"], n_sub)
        f = np.zeros(m.n_u_nodes * 2)
        rho_q, mu_q = compute_force_2d(m, n_sub, H, kappa, prm["surface_tension"], prm["gravity"], prm["density"], prm["density_difference"],
                                       prm["viscosity"], prm["viscosity_difference"], True, f, interpolate=True)
        return f, rho_q, mu_q

    c1, counts = reinitialize(O, c_old, 2)
    assert counts == gold["time_step_1_reinitialization_cg_iterations"]["value"]
    f, rho_q, mu_q = force(c1)
    A = OracleNavierStokes2DVariable(m, rho_q, mu_q, 0.0, 0.0, 0.5, (1.0 / dt, -1.0 / dt, 0.0, 1.0))
    u1, p1, history = newton_with_probed_jacobian(A, m, f)
    assert f"{history[0]:.3g}" == gold["time_step_1_first_navier_stokes_residual"]["value"] and history[-1] < 1e-9
    # 'Size spurious currents, absolute: 3.776641e-05' (line 12): the velocity field itself, to 6 digits
    from oracle.ns_oracle_2d import _q2
    v1 = _q2(np.linspace(0.0, 1.0, 5))[0]
    phi = np.einsum("bj,ai->baji", v1, v1).reshape(25, 9)
    size = np.linalg.norm(np.einsum("qn,cnd->cqd", phi, u1.reshape(-1, 2)[m.cell_u]), axis=2).max()
    assert abs(size / float(gold["time_step_1_size_spurious_currents"]["value"]) - 1.0) < 1e-6
    # 'Error in pressure jump: 0.57390569 %' (line 11, spurious_currents.cc:152-215): mean pressure of the 12 cells around the
    # centre against the mean over the boundary faces (Q1: vertex means), relative to 2 sigma -- the pressure field, to 3e-7
    X, pc = m.p_coords, p1[m.cell_p]
    inner = np.linalg.norm(X[m.cell_p].mean(axis=1), axis=1) < 0.1
    face_means = []
    for a, b in ((0, 1), (2, 3), (0, 2), (1, 3)):            # the four faces of a cell as vertex pairs (x fastest)
        xa, xb = X[m.cell_p[:, a]], X[m.cell_p[:, b]]
        for d in range(2):
            on = (np.abs(np.abs(xa[:, d]) - 2.5) < 1e-12) & (xa[:, d] == xb[:, d])
            face_means.append(0.5 * (pc[on, a] + pc[on, b]))
    face_means = np.concatenate(face_means)
    assert inner.sum() == 12 and len(face_means) == 320
    jump_error = ((pc[inner].mean() - face_means.mean()) - 2.0 * prm["surface_tension"]) / (2.0 * prm["surface_tension"]) * 100.0
    assert abs(jump_error - float(gold["time_step_1_pressure_jump_error_percent"]["value"])) < 1e-6
    g2 = gold["time_step_2_concentration_advance"]
    c_star = 2.0 * c1 - c_old
    rhs = np.zeros(O.n)
    O.advance_concentration_rhs(rhs, u1, c_star, c1, c_old, 1.5 / dt, -2.0 / dt, 0.5 / dt)
    inc, n_it, res0 = bicgstab_iterations(lambda v: O.advance_concentration_vmult(v, 1.5 / dt), rhs, lambda r: r / O.diag, 30, 0.05 * 1e-9, 1e-8)
    assert f"{res0:.3g}" == g2["advection_initial_residual"] and n_it == g2["advection_bicgstab_iterations"]
    c2, counts = reinitialize(O, c_star + inc, 2)
    assert counts == g2["reinitialization_cg_iterations"]
    f2, rho_q, mu_q = force(c2)
    A2 = OracleNavierStokes2DVariable(m, rho_q, mu_q, 0.0, 0.0, 0.5, (1.5 / dt, -2.0 / dt, 0.5 / dt, 1.0))
    A2.old_u, A2.old_old_u = u1.copy(), np.zeros(A2.n_u)
    ru, rp = np.zeros(A2.n_u), np.zeros(A2.n_p)
    A2.residual(ru, rp, u1, p1, f2, np.zeros(A2.n_p))
    assert f"{np.sqrt(ru @ ru + rp @ rp):.3g}" == gold["time_step_2_first_navier_stokes_residual"]["value"]


def test_2d_variable_coefficient_oracle_equals_extruded_3d_oracle():
    """OracleNavierStokes2DVariable (the operator of the golden-log chains above) against the pinned 3-D oracle with
    per-point density / viscosity on the one-cell extrusion, and against the constant-coefficient 2-D oracle."""
    from oracle.ns_oracle_2d import OracleNavierStokes2DVariable
    m = BoxMesh2D((5, 4))
    _, oprm = flow_params(linearization=capi.NEWTON, damping=-0.2, tau_grad_div=0.3, density=1.3, viscosity=0.7, density_diff=9.0, viscosity_diff=0.5)
    _, otw = time_weights()
    tw = (otw.weight, otw.weight_old, otw.weight_old_old, otw.tau1)
    # a smooth two-phase-like field at the 9 points of every cell
    H = 0.5 * (1.0 + np.tanh((0.6 - np.linalg.norm(m.u_coords[m.cell_u].mean(axis=1), axis=1)) / 0.2))[:, None] * (1.0 + 0.05 * np.arange(9)[None, :])
    rho_q, mu_q = 1.3 + 9.0 * H, 0.7 + 0.5 * H
    O3 = orc.OracleNavierStokesMatrix(m.mesh3d, oprm, otw, None, False)
    O3.rho[:], O3.muq[:], O3.damp[:] = np.tile(rho_q, (1, 3)), np.tile(mu_q, (1, 3)), -0.2   # points q = qx + 3 qy + 9 qz
    O2 = OracleNavierStokes2DVariable(m, rho_q, mu_q, -0.2, 0.3, 0.5, tw)
    u, p = fields(m, 3)
    uo, _ = fields(m, 5)
    uoo, _ = fields(m, 7)
    O2.old_u, O2.old_old_u = uo.copy(), uoo.copy()
    O3.old_u, _ = m.extend(uo, np.zeros(m.n_p_nodes))
    O3.old_old_u, _ = m.extend(uoo, np.zeros(m.n_p_nodes))
    r2u, r2p = np.zeros(O2.n_u), np.zeros(O2.n_p)
    O2.residual(r2u, r2p, u, p, np.zeros(O2.n_u), np.zeros(O2.n_p))
    r3u, r3p = np.zeros(O3.n_u), np.zeros(O3.n_p)
    O3.residual(r3u, r3p, *m.extend(u, p), np.zeros(O3.n_u), np.zeros(O3.n_p))
    ru, rp = m.restrict(r3u, r3p)
    assert rel_err(ru, r2u) < 1e-13 and rel_err(rp, r2p) < 1e-13
    su, sp = fields(m, 11)
    d2u, d2p = O2.vmult(su, sp)
    d3u, d3p = O3.vmult(*m.extend(su, sp))
    du, dp = m.restrict(d3u, d3p)
    c = m.u_dirichlet.reshape(-1)
    du[c] = su[c]
    assert rel_err(du, d2u) < 1e-13 and rel_err(dp, d2p) < 1e-13
    # constant coefficients: identical to the base class
    Oc = OracleNavierStokes2DVariable(m, np.full((m.n_cells, 9), 1.3), np.full((m.n_cells, 9), 0.7), -0.2, 0.3, 0.5, tw)
    Ob = OracleNavierStokes2D(m, capi.INCOMPRESSIBLE, capi.NEWTON, 1.3, 0.7, -0.2, 0.3, 0.5, tw, pressure_average_fix=False)
    for O in (Oc, Ob):
        O.old_u, O.old_old_u = uo.copy(), uoo.copy()
    ra, rb = [np.zeros(Oc.n_u), np.zeros(Oc.n_p)], [np.zeros(Oc.n_u), np.zeros(Oc.n_p)]
    Oc.residual(ra[0], ra[1], u, p, np.zeros(Oc.n_u), np.zeros(Oc.n_p))
    Ob.residual(rb[0], rb[1], u, p, np.zeros(Oc.n_u), np.zeros(Oc.n_p))
    assert rel_err(ra[0], rb[0]) < 1e-14 and rel_err(ra[1], rb[1]) < 1e-14
    va, vb = Oc.vmult(su, sp), Ob.vmult(su, sp)
    assert rel_err(va[0], vb[0]) < 1e-14 and rel_err(va[1], vb[1]) < 1e-14


def rising_bubble_first_step(fixture):
    """Oracle chain of set-up and time step 1 of a rising-bubble log (shared with tests/test_2d_gpu.py): returns the golden
    values, the mesh, eps_used and the fields that enter time step 2 -- level sets c_0 (after set-up), c_1 (after step 1)
    and the converged velocity / pressure (u_1, p_1)."""
    from oracle.ls_oracle_2d import OracleLevelSet2D, compute_force_2d
    from oracle.ns_oracle_2d import OracleNavierStokes2DVariable
    gold, m, eps_used, c = rising_bubble_setup(fixture)
    prm = gold["parameters"]
    n_sub, dt = prm["concentration_subdivisions"], prm["step_size"]
    O = OracleLevelSet2D(m, n_sub, prm["epsilon"], eps_used, prm["h"], dt, np.full(m.n_cells, prm["h"]))
    O.diag = O.mass_diagonal()
    x = m.u_coords
    dirichlet = np.zeros_like(m.u_dirichlet)
    dirichlet[(np.abs(x[:, 1]) < 1e-12) | (np.abs(x[:, 1] - 2) < 1e-12), :] = True
    dirichlet[(np.abs(x[:, 0]) < 1e-12) | (np.abs(x[:, 0] - 1) < 1e-12), 0] = True
    m.u_dirichlet = dirichlet
    c0, _ = reinitialize(O, c, 2)
    c1, _ = reinitialize(O, c0, 2)
    kappa = corrected_curvature(O, c1, eps_used)
    H = orc.compute_heaviside(m.cell_ls, c1, prm["epsilon"], n_sub)
    f = np.zeros(m.n_u_nodes * 2)
    rho_q, mu_q = compute_force_2d(m, n_sub, H, kappa, prm["surface_tension"], prm["gravity"], prm["density"], prm["density_difference"],
                                   prm["viscosity"], prm["viscosity_difference"], True, f, interpolate=True, u_dirichlet=dirichlet)
    A = OracleNavierStokes2DVariable(m, rho_q, mu_q, 0.0, 0.0, 0.5, (1.0 / dt, -1.0 / dt, 0.0, 1.0))
    u1, p1, history = newton_with_probed_jacobian(A, m, f)
    assert history[-1] < 1e-9
    return gold, m, eps_used, c0, c1, u1, p1
