"""The optimised CPU port (bench baseline) agrees with the dense oracle (CPU only)."""
import numpy as np
import pytest

from adaflo_b200 import capi
from adaflo_b200.mesh import BoxMesh, beltrami_velocity, seeded_uniform
from oracle import cpu_port, oracle as orc

from helpers import flow_params, rel_err, time_weights


@pytest.mark.parametrize("lin,mode", [(capi.NEWTON, 1), (capi.PICARD, 2)])
def test_port_matches_dense_oracle(lin, mode):
    mesh = BoxMesh((5, 4, 6))
    _, oprm = flow_params(linearization=lin, damping=-0.2, tau_grad_div=0.3, density=1.3, viscosity=0.7)
    _, tw = time_weights()
    ns = orc.OracleNavierStokesMatrix(mesh, oprm, tw)
    u = beltrami_velocity(mesh.u.coords, 0.1).reshape(-1)
    p = seeded_uniform(mesh.p.global_ids, 1, 3)
    ns.old_u = u.copy()
    ns.old_old_u = u.copy()
    z = np.zeros
    ns.residual(z(ns.n_u), z(ns.n_p), u, p, z(ns.n_u), z(ns.n_p))      # fills the linearisation point
    su, sp = seeded_uniform(mesh.u.global_ids, 3, 4), seeded_uniform(mesh.p.global_ids, 1, 5)
    ou, op = z(ns.n_u), z(ns.n_p)
    ns._loop(orc.OP_VMULT, su, sp, ou, op)
    port = cpu_port.CpuPortVmult(mesh, mode, tw.weight, tw.tau1, oprm.beta, oprm.density, oprm.viscosity, oprm.damping,
                                 oprm.tau_grad_div, ns.lin)
    gu, gp = z(ns.n_u), z(ns.n_p)
    port.cells(su, sp, gu, gp)
    assert rel_err(gu, ou) < 1e-13 and rel_err(gp, op) < 1e-13


@pytest.mark.parametrize("n", [(3, 3, 3), (9, 3, 2)])   # scalar cells only / two SIMD batches of 4 cells and a scalar remainder per row
def test_port_stokes(n):
    mesh = BoxMesh(n)
    _, oprm = flow_params(physical_type=capi.STOKES, viscosity=0.7)
    _, tw = time_weights()
    ns = orc.OracleNavierStokesMatrix(mesh, oprm, tw)
    su, sp = seeded_uniform(mesh.u.global_ids, 3, 4), seeded_uniform(mesh.p.global_ids, 1, 5)
    ou, op = np.zeros(ns.n_u), np.zeros(ns.n_p)
    ns._loop(orc.OP_VMULT, su, sp, ou, op)
    port = cpu_port.CpuPortVmult(mesh, 0, 0.0, 1.0, 0.5, 1.0, 0.7, 0.0, 0.0)
    gu, gp = np.zeros(ns.n_u), np.zeros(ns.n_p)
    port.cells(su, sp, gu, gp)
    assert rel_err(gu, ou) < 1e-13 and rel_err(gp, op) < 1e-13
