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
