"""The generic-degree / generic-dimension oracle (oracle/ns_oracle_generic.py) against the pinned oracles for Q2/Q1 and its
degree-specific parts (CPU only)."""
import numpy as np
import pytest

from adaflo_b200.mesh import BoxMesh, BoxMesh2D, TensorMesh, seeded_uniform
from oracle import ns_oracle_2d as o2
from oracle import ns_oracle_generic as og
from oracle import oracle as orc

from helpers import flow_params, rel_err, time_weights


def fields(m, seed=0):
    u = seeded_uniform(m.u_ids, m.dim, 1 + seed)
    p = seeded_uniform(m.p_ids, 1, 2 + seed)
    return u, p


def test_generic_oracle_equals_the_pinned_2d_oracle_for_q2q1():
    m2 = BoxMesh2D((5, 4), lower=(0.0, -0.5), upper=(1.25, 0.5))
    mt = TensorMesh(2, (5, 4), 2, lower=(0.0, -0.5), upper=(1.25, 0.5))
    assert np.array_equal(mt.cell_u, m2.cell_u) and np.array_equal(mt.cell_p, m2.cell_p) and np.array_equal(mt.u_dirichlet, m2.u_dirichlet)
    assert np.allclose(mt.u_coords, m2.u_coords) and np.allclose(mt.p_coords, m2.p_coords)
    tw = (1.5 / 0.1, -2.0 / 0.1, 0.5 / 0.1, 1.0)
    for lin in (o2.NEWTON, o2.PICARD):
        A = o2.OracleNavierStokes2D(m2, o2.INCOMPRESSIBLE, lin, 1.3, 0.7, -0.2, 0.3, 0.5, tw)
        B = og.OracleNavierStokesGeneric(mt, og.INCOMPRESSIBLE, lin, 1.3, 0.7, -0.2, 0.3, 0.5, tw)
        u, p = fields(mt)
        for O in (A, B):
            O.old_u, O.old_old_u = seeded_uniform(mt.u_ids, 2, 7), seeded_uniform(mt.u_ids, 2, 8)
        ra, rb = [np.zeros(A.n_u), np.zeros(A.n_p)], [np.zeros(B.n_u), np.zeros(B.n_p)]
        A.residual(ra[0], ra[1], u, p, np.zeros(A.n_u), np.zeros(A.n_p))
        B.residual(rb[0], rb[1], u, p, np.zeros(B.n_u), np.zeros(B.n_p))
        assert rel_err(rb[0], ra[0]) < 1e-13 and rel_err(rb[1], ra[1]) < 1e-13
        su, sp = fields(mt, 10)
        (au, ap), (bu, bp) = A.vmult(su, sp), B.vmult(su, sp)
        assert rel_err(bu, au) < 1e-13 and rel_err(bp, ap) < 1e-13


def test_generic_oracle_equals_the_pinned_3d_oracle_for_q2q1():
    m3 = BoxMesh((3, 2, 2))
    mt = TensorMesh(3, (3, 2, 2), 2)
    assert np.array_equal(mt.cell_u, m3.u.cell_nodes.astype(np.int64)) and np.array_equal(mt.cell_p, m3.p.cell_nodes.astype(np.int64))
    _, oprm = flow_params(damping=-0.2, tau_grad_div=0.3, viscosity=0.7, density=1.3)
    _, otw = time_weights()
    A = orc.OracleNavierStokesMatrix(m3, oprm, otw)
    B = og.OracleNavierStokesGeneric(mt, og.INCOMPRESSIBLE, og.NEWTON, 1.3, 0.7, -0.2, 0.3, oprm.beta,
                                     (otw.weight, otw.weight_old, otw.weight_old_old, otw.tau1))
    u, p = fields(mt)
    uo, uoo = seeded_uniform(mt.u_ids, 3, 7), seeded_uniform(mt.u_ids, 3, 8)
    A.old_u, A.old_old_u, B.old_u, B.old_old_u = uo, uoo, uo, uoo
    ra, rb = [np.zeros(A.n_u), np.zeros(A.n_p)], [np.zeros(B.n_u), np.zeros(B.n_p)]
    A.residual(ra[0], ra[1], u, p, np.zeros(A.n_u), np.zeros(A.n_p))
    B.residual(rb[0], rb[1], u, p, np.zeros(B.n_u), np.zeros(B.n_p))
    assert rel_err(rb[0], ra[0]) < 1e-13 and rel_err(rb[1], ra[1]) < 1e-13
    su, sp = fields(mt, 10)
    (au, ap), (bu, bp) = A.vmult(su, sp), B.vmult(su, sp)
    assert rel_err(bu, au) < 1e-13 and rel_err(bp, ap) < 1e-13
    # sub-block operators of the block preconditioner (navier_stokes_matrix.cc:300-483)
    for weight in (False, True):
        da, db = np.zeros(A.n_p), np.zeros(B.n_p)
        A.divergence_vmult_add(da, su, weight)
        B.divergence_vmult_add(db, su, weight)
        assert rel_err(db, da) < 1e-13
    assert rel_err(B.pressure_mass_vmult(sp), A.pressure_mass_vmult(sp)) < 1e-13
    assert rel_err(B.pressure_poisson_vmult(sp), A.pressure_poisson_vmult(sp)) < 1e-13


@pytest.mark.parametrize("dim,k", [(2, 3), (2, 4), (3, 3)])
def test_higher_degrees_integrate_exactly_and_vmult_is_the_jacobian(dim, k):
    n = (3, 2) if dim == 2 else (2, 2, 1)
    m = TensorMesh(dim, n, k)
    O = og.OracleNavierStokesGeneric(m, og.INCOMPRESSIBLE, og.NEWTON, 1.3, 0.7, 0.0, 0.0, 0.5, (10.0, -10.0, 0.0, 1.0))
    # partition of unity and exact quadrature: sum_n phi_n = 1, the weights integrate 1 over the cell
    assert np.allclose(O.phi_u.sum(axis=1), 1.0) and np.allclose(O.phi_p.sum(axis=1), 1.0) and abs(O.jxw[0].sum() - np.prod(m.h)) < 1e-14
    # a polynomial of degree k in every variable is reproduced at the quadrature points
    x = m.u_coords
    f = np.prod(1.0 + 0.3 * x ** k - 0.2 * x, axis=1)
    cell0 = m.cell_u[0]
    xq = og.gauss(k + 1)[0]
    pts = [m.lower[d] + m.h[d] * xq for d in range(dim)]
    grid = np.stack(np.meshgrid(*pts[::-1], indexing="ij")[::-1], axis=-1).reshape(-1, dim)
    assert np.allclose(O.phi_u @ f[cell0], np.prod(1.0 + 0.3 * grid ** k - 0.2 * grid, axis=1), atol=1e-12)
    # vmult = Jacobian of the residual (central differences) at an interior direction
    u, p = seeded_uniform(m.u_ids, dim, 1), seeded_uniform(m.p_ids, 1, 2)
    u[m.u_dirichlet.reshape(-1)] = 0.0
    du, dp = seeded_uniform(m.u_ids, dim, 3), seeded_uniform(m.p_ids, 1, 4)
    du[m.u_dirichlet.reshape(-1)] = 0.0

    def res(uu, pp):
        ru, rp = np.zeros(O.n_u), np.zeros(O.n_p)
        O.residual(ru, rp, uu, pp, np.zeros(O.n_u), np.zeros(O.n_p))
        return ru, rp
    eps = 1e-6
    (ru1, rp1), (ru0, rp0) = res(u + eps * du, p + eps * dp), res(u - eps * du, p - eps * dp)
    res(u, p)                                  # linearisation point = (u, p)
    O.mode_w = None                            # compare the raw operator
    ju, jp = O.vmult(du, dp)
    free = ~m.u_dirichlet.reshape(-1)
    assert rel_err((-(ru1 - ru0) / (2 * eps))[free], ju[free]) < 1e-7 and rel_err(-(rp1 - rp0) / (2 * eps), jp) < 1e-7


# ---------------------------------------------------------------- hanging nodes of any degree: tests/beltrami_2d.output (Q4/Q3)
import json
import os

from adaflo_b200.mesh import RefinedTensorMesh2D, beltrami_pressure, beltrami_velocity

with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "beltrami_2d_output.json")) as _f:
    GOLDEN_2D = json.load(_f)   # numbers of the reference's golden log tests/beltrami_2d.output


@pytest.fixture(scope="module")
def beltrami_2d():
    par = GOLDEN_2D["parameters"]
    n = par["coarse_cells_per_direction"]
    mesh = RefinedTensorMesh2D((n, n), par["refined_block"], par["velocity_degree"])     # tests/beltrami.cc:385-412
    dt = par["dt"]
    # time step #1 of BDF-2: implicit Euler weights (time_stepping.cc:158-198); Newton, beta = 0.5, nu = rho = 1
    O = og.OracleNavierStokesGeneric(mesh, og.INCOMPRESSIBLE, og.NEWTON, 1.0, par["viscosity"], 0.0, 0.0, 0.5, (1 / dt, -1 / dt, 0.0, 1.0))
    return mesh, O, dt


def test_beltrami_2d_dof_counts_match_the_golden_log(beltrami_2d):
    mesh, _, _ = beltrami_2d
    # beltrami_2d.output:2-3; deal.II counts the constrained (hanging) dofs
    assert mesh.n_cells == GOLDEN_2D["n_active_cells"]["value"] == 1048
    assert 2 * mesh.n_u_nodes == GOLDEN_2D["n_dofs"]["velocity"] == 34158
    assert mesh.n_p_nodes == GOLDEN_2D["n_dofs"]["pressure"] == 9663
    # 10 coarse faces carry hanging nodes (4 above, 4 below, 2 right of the block; the left side is the boundary): 2 k - 1 each
    assert mesh.n_hanging_u == 10 * 7 and mesh.n_hanging_p == 10 * 5


def test_beltrami_2d_interpolation_errors_match_the_golden_log(beltrami_2d):
    mesh, O, _ = beltrami_2d
    u0 = beltrami_velocity(mesh.u_coords, 0.0).reshape(-1)
    p0 = beltrami_pressure(mesh.p_coords, 0.0)
    e_u = O.l2_error(0, u0, lambda x: beltrami_velocity(x, 0.0), 6)          # QGauss(degree + 2), tests/beltrami.cc:254
    e_p = O.l2_error(1, p0, lambda x: beltrami_pressure(x, 0.0)[..., None], 6)
    # beltrami_2d.output:5 -- this pins the Gauss-Lobatto support points and the Q4 / Q3 shape functions
    assert f"{e_p:.3e}" == GOLDEN_2D["l2_errors_t0"]["e_p"] == "9.507e-09"
    assert f"{e_u:.3e}" == GOLDEN_2D["l2_errors_t0"]["e_u"] == "8.461e-12"


def test_beltrami_2d_first_nonlinear_residual_matches_the_golden_log(beltrami_2d):
    mesh, O, dt = beltrami_2d
    u0 = beltrami_velocity(mesh.u_coords, 0.0).reshape(-1)
    p0 = beltrami_pressure(mesh.p_coords, 0.0)                       # interpolate_pressure_field: nodal values, also at hanging nodes
    u = u0.copy()
    d = mesh.u_dirichlet.reshape(-1)
    u[d] = beltrami_velocity(mesh.u_coords, dt).reshape(-1)[d]       # apply_boundary_conditions, navier_stokes.cc:1216-1251
    mesh.distribute(u, 0)                                            # hanging_node_constraints_u.distribute, :1252
    O.old_u, O.old_old_u = u0.copy(), np.zeros_like(u0)
    ru, rp = np.zeros(O.n_u), np.zeros(O.n_p)
    O.residual(ru, rp, u, p0, np.zeros(O.n_u), np.zeros(O.n_p))
    O.apply_pressure_average_projection(rp)
    # beltrami_2d.output:13: "4.149e-01   1.021e-03" -- Q4/Q3, hanging nodes through the constraint rows, two cell sizes
    assert f"{np.linalg.norm(ru):.3e}" == GOLDEN_2D["first_nonlinear_residual"]["1"]["u"] == "4.149e-01"
    assert f"{np.linalg.norm(rp):.3e}" == GOLDEN_2D["first_nonlinear_residual"]["1"]["p"] == "1.021e-03"


@pytest.mark.parametrize("k", [2, 3, 4, 5])
def test_constraint_rows_reproduce_polynomials_and_keep_the_operator_consistent(k):
    """the hanging-node rows interpolate degree-k polynomials exactly; on distributed vectors the constraint-resolving read of
    vmult equals the plain read, so vmult stays the Jacobian of the residual on the refined mesh"""
    mesh = RefinedTensorMesh2D((4, 3), ((1, 3), (0, 2)), k, lower=(-1.0, -0.5), upper=(1.0, 0.75))
    f = lambda x: (0.3 + x[:, 0]) ** k * (1.1 - 0.7 * x[:, 1]) ** k
    for space, coords in ((0, mesh.u_coords), (1, mesh.p_coords)):
        kk = k - space
        v = ((0.3 + coords[:, 0]) ** kk * (1.1 - 0.7 * coords[:, 1]) ** kk).copy()
        assert np.abs(mesh.distribute(v.copy(), space) - v).max() < 1e-13
    O = og.OracleNavierStokesGeneric(mesh, og.INCOMPRESSIBLE, og.NEWTON, 1.3, 0.7, 0.0, 0.0, 0.5, (10.0, -10.0, 0.0, 1.0))
    rng = np.random.default_rng(k)
    free_u = ~O.u_constrained.reshape(-1)

    def consistent(seed):
        r = np.random.default_rng(seed)
        u = r.uniform(-1, 1, O.n_u)
        p = r.uniform(-1, 1, O.n_p)
        return mesh.distribute(u, 0), mesh.distribute(p, 1)

    u, p = consistent(1)
    du, dp = consistent(2)
    du[~free_u] = 0.0
    mesh.distribute(du, 0)
    O.old_u, O.old_old_u = consistent(3)[0], consistent(4)[0]

    def R(uu, pp):
        ru, rp = np.zeros(O.n_u), np.zeros(O.n_p)
        O.residual(ru, rp, uu, pp, np.zeros(O.n_u), np.zeros(O.n_p))
        return -ru, -rp

    eps = 1e-6
    (rpu, rpp), (rmu, rmp) = R(u + eps * du, p + eps * dp), R(u - eps * du, p - eps * dp)
    R(u, p)                                                        # linearisation point = (u, p)
    O.mode_w = None                                                # compare the cell integrals
    ju, jp = O.vmult(du, dp)
    fd_u, fd_p = (rpu - rmu) / (2 * eps), (rpp - rmp) / (2 * eps)
    free_p = ~O.p_constrained
    assert np.abs(ju[free_u] - fd_u[free_u]).max() < 1e-6 * max(1.0, np.abs(fd_u).max())
    assert np.abs(jp[free_p] - fd_p[free_p]).max() < 1e-6 * max(1.0, np.abs(fd_p).max())
    # constrained rows: identity / minus identity (navier_stokes_matrix.cc:247-256)
    assert np.array_equal(ju[~free_u], du[~free_u]) and np.array_equal(jp[~free_p], -dp[~free_p])
