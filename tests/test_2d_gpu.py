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
