"""Parity at the sizes of the benchmark (BASELINE.json configs[1] scaled to 128^3 cells, 53 M DoFs).

The dense oracle cannot run there, so (1) a 128 x 128 x 8 slab of the same workload (3.5 M DoFs) is compared with
the optimised CPU port -- which tests/test_cpu_port.py ties to the dense oracle -- entry by entry, and (2) the full
128^3 operator is checked through size-independent properties: linearity, the transpose identity, the constant
pressure null space, host path == device path, run-to-run repeatability.
"""
import numpy as np
import pytest
import torch

from adaflo_b200 import capi
from adaflo_b200.mesh import BoxMesh, beltrami_velocity, seeded_uniform
from adaflo_b200.operators import BlockVector, Context, MatrixFree, NavierStokesMatrix

from helpers import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module")
def ctx():
    c = Context(0)
    yield c


def _setup(ctx, n, upper, pressure_mean=True):
    mesh = BoxMesh(n, upper=upper)
    mf = MatrixFree(ctx, mesh, "cartesian")
    prm = capi.FlowParams(capi.INCOMPRESSIBLE, capi.NEWTON, 0, 0, 1.0, 1.0, 0.0, 0.0, 0.0, 0.0, 0.5)
    dt = 0.05
    tw = capi.TimeWeights(1.5 / dt, -2.0 / dt, 0.5 / dt, 1.0, 2.0, -1.0)     # BDF-2, tests/beltrami_3d.prm
    sol, old, old_old = BlockVector.like(mf), BlockVector.like(mf), BlockVector.like(mf)
    A = NavierStokesMatrix(prm, 0, 1, 0, 1, old, old_old)
    A.initialize(mf, tw, pressure_mean)
    u0 = beltrami_velocity(mesh.u.coords, 0.0).reshape(-1)
    sol.block(0).copy_(torch.from_numpy(u0).to(ctx.device))
    old.block(0).copy_(sol.block(0))
    old_old.block(0).copy_(sol.block(0))
    mesh.u.coords = None
    rhs, zero = BlockVector.like(mf), BlockVector.like(mf)
    A.residual(rhs, sol, zero)           # produces the linearisation point
    return mesh, mf, A, tw, (sol, old, old_old)


def _seeded(ctx, mesh, seed):
    return BlockVector(torch.from_numpy(seeded_uniform(mesh.u.global_ids, 3, seed)).to(ctx.device),
                       torch.from_numpy(seeded_uniform(mesh.p.global_ids, 1, seed + 1)).to(ctx.device))


def test_slab_of_the_workload_matches_cpu_port(ctx):
    """128 x 128 x 8 cells of the 128^3 workload: cell loop of vmult vs oracle/ns_cpu_port.c on all 3.5 M entries"""
    from oracle import cpu_port
    mesh, mf, A, tw, keep = _setup(ctx, (128, 128, 8), (1.0, 1.0, -1.0 + 2.0 * 8 / 128), pressure_mean=False)
    lin = A.get_coefficient(capi.COEF_LINEARIZED_VELOCITY).cpu().numpy()
    port = cpu_port.CpuPortVmult(mesh, 1, tw.weight, tw.tau1, 0.5, 1.0, 1.0, 0.0, 0.0, np.ascontiguousarray(lin))
    src = _seeded(ctx, mesh, 12345)
    dst = BlockVector.like(mf)
    A.vmult(dst, src)
    su, sp = src.block(0).cpu().numpy(), src.block(1).cpu().numpy()
    ou, op = np.zeros_like(su), np.zeros_like(sp)
    port.cells(su, sp, ou, op)
    cu = mesh.u.constrained_dofs()                       # constrained rows, navier_stokes_matrix.cc:247-256
    ou[cu] = su[cu]
    cp = mesh.p.constrained_dofs()
    op[cp] = -sp[cp]
    gu, gp = dst.block(0).cpu().numpy(), dst.block(1).cpu().numpy()
    assert rel_err(gu, ou) < TOL and rel_err(gp, op) < TOL
    assert np.max(np.abs(gu - ou)) < TOL * np.max(np.abs(ou))      # entry by entry, not only in the norm


def test_full_size_properties(ctx):
    mesh, mf, A, tw, keep = _setup(ctx, (128, 128, 128), (1.0, 1.0, 1.0))
    n_u, n_p = mf.n_owned(0), mf.n_owned(1)
    assert n_u + n_p == 53070468
    x, y = _seeded(ctx, mesh, 1), _seeded(ctx, mesh, 3)
    Ax, Ay, Az = BlockVector.like(mf), BlockVector.like(mf), BlockVector.like(mf)
    A.vmult(Ax, x)
    A.vmult(Ay, y)
    # linearity: A (2 x - 3 y) = 2 A x - 3 A y
    z = BlockVector(2.0 * x.block(0) - 3.0 * y.block(0), 2.0 * x.block(1) - 3.0 * y.block(1))
    A.vmult(Az, z)
    for b in range(2):
        ref = 2.0 * Ax.block(b) - 3.0 * Ay.block(b)
        assert float((Az.block(b) - ref).norm() / ref.norm()) < TOL
    # repeatability (atomics reorder the additions: round-off, not bitwise)
    Ax2 = BlockVector.like(mf)
    A.vmult(Ax2, x)
    assert float((Ax2.block(0) - Ax.block(0)).norm() / Ax.block(0).norm()) < 1e-14
    # transpose identity <A^T x, y> = <x, A y>
    Atx = BlockVector.like(mf)
    A.Tvmult(Atx, x)
    lhs = float(Atx.block(0) @ y.block(0) + Atx.block(1) @ y.block(1))
    rhs = float(x.block(0) @ Ay.block(0) + x.block(1) @ Ay.block(1))
    assert abs(lhs - rhs) <= 1e-11 * max(abs(lhs), abs(rhs))
    # constant pressure is in the null space: with u = 0 on the whole boundary, int p div v = 0 for constant p, and
    # the mean projection removes the constant from the pressure rows
    c = BlockVector.like(mf)
    c.block(1).fill_(1.0)
    Ac = BlockVector.like(mf)
    A.vmult(Ac, c)
    scale = float(Ax.block(0).abs().max())
    assert float(Ac.block(0).abs().max()) < 1e-11 * scale and float(Ac.block(1).abs().max()) < 1e-11
    # host entry point (three-stream pipeline, 16 cell chunks) == device path
    hu, hp = x.block(0).cpu().pin_memory(), x.block(1).cpu().pin_memory()
    ou, op = torch.empty_like(hu).pin_memory(), torch.empty_like(hp).pin_memory()
    A.vmult_host(ou, op, hu, hp)
    assert float((ou.to(ctx.device) - Ax.block(0)).norm() / Ax.block(0).norm()) < 1e-14
    assert float((op.to(ctx.device) - Ax.block(1)).norm() / Ax.block(1).norm()) < 1e-14
