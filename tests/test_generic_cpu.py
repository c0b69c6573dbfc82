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


@pytest.mark.parametrize("dim,k", [(2, 3), (2, 4), (3, 3)])
def test_higher_degrees_integrate_exactly_and_vmult_is_the_jacobian(dim, k):
    n = (3, 2) if dim == 2 else (2, 2, 1)
    m = TensorMesh(dim, n, k)
    O = og.OracleNavierStokesGeneric(m, og.INCOMPRESSIBLE, og.NEWTON, 1.3, 0.7, 0.0, 0.0, 0.5, (10.0, -10.0, 0.0, 1.0))
    # partition of unity and exact quadrature: sum_n phi_n = 1, the weights integrate 1 over the cell
    assert np.allclose(O.phi_u.sum(axis=1), 1.0) and np.allclose(O.phi_p.sum(axis=1), 1.0) and abs(O.jxw.sum() - np.prod(m.h)) < 1e-14
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
