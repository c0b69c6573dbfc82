"""Degree / dimension dispatch of the Navier-Stokes operator (csrc/ns_generic.cu; reference: EXPAND_OPERATIONS,
source/navier_stokes_matrix.cc:64-82, dim 1-3 :1210-1212) against the oracles, through the C ABI:
  * 3-D Q2/Q1 forced onto the generic kernel vs the pinned dense oracle (oracle/ns_oracle.c), all geometry types;
  * native 2-D Q2/Q1 vs the pinned 2-D oracle (couette.output) and Q3/Q2, Q4/Q3 (tests/beltrami_2d.prm's pair) in 2-D,
    Q3/Q2 and Q4/Q3 in 3-D vs oracle/ns_oracle_generic.py (proved equal to the pinned oracles for Q2/Q1 in
    tests/test_generic_cpu.py)."""
import os

import numpy as np
import pytest
import torch

from adaflo_b200 import capi
from adaflo_b200.mesh import BoxMesh, RefinedTensorMesh2D, TensorMesh, seeded_uniform
from adaflo_b200.operators import BlockVector, Context, MatrixFree, MatrixFreeGeneric, NavierStokesMatrix
from oracle import ns_oracle_generic as og
from oracle import oracle as orc

from helpers import deformation, flow_params, rel_err, time_weights

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module")
def ctx():
    return Context(0)


def dev(ctx, a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(ctx.device)


@pytest.mark.parametrize("geometry", ["cartesian", "affine", "general"])
@pytest.mark.parametrize("linearization", [capi.NEWTON, capi.PICARD])
def test_generic_kernel_on_3d_q2q1_matches_the_pinned_oracle(ctx, geometry, linearization):
    mesh = BoxMesh((4, 3, 2))
    deform = deformation(geometry)
    dprm, oprm = flow_params(linearization=linearization, damping=-0.2, tau_grad_div=0.3, viscosity=0.7, density=1.3)
    dtw, otw = time_weights()
    os.environ["ADAFLO_B200_NS_GENERIC"] = "1"
    try:
        mf = MatrixFree(ctx, mesh, geometry, deform)
        old, old_old = BlockVector.like(mf), BlockVector.like(mf)
        A = NavierStokesMatrix(dprm, 0, 1, 0, 1, old, old_old)
        A.initialize(mf, dtw, True)
    finally:
        os.environ.pop("ADAFLO_B200_NS_GENERIC")
    O = orc.OracleNavierStokesMatrix(mesh, oprm, otw, deform)
    u, p = seeded_uniform(mesh.u.global_ids, 3, 1), seeded_uniform(mesh.p.global_ids, 1, 2)
    uo, uoo = seeded_uniform(mesh.u.global_ids, 3, 3), seeded_uniform(mesh.u.global_ids, 3, 4)
    old.block(0).copy_(dev(ctx, uo))
    old_old.block(0).copy_(dev(ctx, uoo))
    O.old_u, O.old_old_u = uo, uoo
    ru, rp = np.zeros(O.n_u), np.zeros(O.n_p)
    O.residual(ru, rp, u, p, np.zeros(O.n_u), np.zeros(O.n_p))
    rhs, zero, sol = BlockVector.like(mf), BlockVector.like(mf), BlockVector(dev(ctx, u), dev(ctx, p))
    A.residual(rhs, sol, zero)
    assert rel_err(rhs.block(0).cpu().numpy(), ru) < TOL and rel_err(rhs.block(1).cpu().numpy(), rp) < TOL
    su, sp = seeded_uniform(mesh.u.global_ids, 3, 11), seeded_uniform(mesh.p.global_ids, 1, 12)
    ou, op = O.vmult(su, sp)
    dst = BlockVector.like(mf)
    A.vmult(dst, BlockVector(dev(ctx, su), dev(ctx, sp)))
    assert rel_err(dst.block(0).cpu().numpy(), ou) < TOL and rel_err(dst.block(1).cpu().numpy(), op) < TOL
    O.fix_linearization_point()
    A.fix_linearization_point()
    dv = mf.initialize_dof_vector(0)
    A.velocity_vmult(dv, dev(ctx, su))
    assert rel_err(dv.cpu().numpy(), O.velocity_vmult(su)) < TOL
    # sub-block operators of the block preconditioner on the generic path vs the pinned oracle
    for weight in (False, True):
        dp, od = dev(ctx, sp.copy()), sp.copy()
        A.divergence_vmult_add(dp, dev(ctx, su), weight)
        O.divergence_vmult_add(od, su, weight)
        assert rel_err(dp.cpu().numpy(), od) < TOL
    dp = mf.initialize_dof_vector(1)
    A.pressure_mass_vmult(dp, dev(ctx, sp))
    assert rel_err(dp.cpu().numpy(), O.pressure_mass_vmult(sp)) < TOL
    A.pressure_poisson_vmult(dp, dev(ctx, sp))
    assert rel_err(dp.cpu().numpy(), O.pressure_poisson_vmult(sp)) < TOL
    # entry points that exist for the tuned 3-D Q2/Q1 kernels only fail loudly
    with pytest.raises(capi.AdafloB200Error):
        A.pressure_convdiff_vmult(mf.initialize_dof_vector(1), dev(ctx, sp))
    A.clear()
    mf.close()


def check_sub_blocks_and_diagonal(ctx, mf, A, O, mesh, su, sp, exact_diagonal):
    """divergence_vmult_add (both weights, adds into dst), pressure_mass_vmult, pressure_poisson_vmult, compute_diagonal of
    the degree / dimension dispatch against the generic oracle"""
    dim = mesh.dim
    for weight in (False, True):
        dp, od = dev(ctx, sp.copy()), sp.copy()
        A.divergence_vmult_add(dp, dev(ctx, su), weight)
        O.divergence_vmult_add(od, su, weight)
        assert rel_err(dp.cpu().numpy(), od) < TOL
    dp = mf.initialize_dof_vector(1)
    A.pressure_mass_vmult(dp, dev(ctx, sp))
    assert rel_err(dp.cpu().numpy(), O.pressure_mass_vmult(sp)) < TOL
    A.pressure_poisson_vmult(dp, dev(ctx, sp))
    assert rel_err(dp.cpu().numpy(), O.pressure_poisson_vmult(sp)) < TOL
    diag = BlockVector.like(mf)
    A.compute_diagonal(diag)
    du = diag.block(0).cpu().numpy()
    con = O.u_constrained.reshape(-1)
    assert np.all(du[con] == 1.0) and np.all(du[~con] > 0.0)
    if exact_diagonal:                                    # e_i^T A e_i for a sample of free dofs
        rng = np.random.default_rng(5)
        for i in rng.choice(np.nonzero(~con)[0], 12, replace=False):
            e = np.zeros(O.n_u)
            e[i] = 1.0
            assert abs(O.velocity_vmult(e)[i] - du[i]) < 1e-12 * abs(du[i])


@pytest.mark.parametrize("dim,k,n", [(2, 2, (6, 5)), (2, 3, (4, 3)), (2, 4, (4, 4)), (2, 5, (2, 3)), (3, 3, (3, 2, 2)), (3, 4, (2, 2, 2))])
@pytest.mark.parametrize("linearization", [capi.NEWTON, capi.PICARD])
def test_taylor_hood_degrees_and_native_2d(ctx, dim, k, n, linearization):
    """Q_k / Q_(k-1) on uniform meshes in 2-D and 3-D (Q4/Q3 in 2-D is the pair of tests/beltrami_2d.prm)"""
    lower = (-1.0, -0.5, 0.0)[:dim]
    upper = (1.0, 0.75, 0.5)[:dim]
    mesh = TensorMesh(dim, n, k, lower=lower, upper=upper)
    mf = MatrixFreeGeneric(ctx, mesh)
    dt = 0.05
    tw = (1.5 / dt, -2.0 / dt, 0.5 / dt, 1.0)
    dprm = capi.FlowParams(capi.INCOMPRESSIBLE, linearization, 0, 0, 1.3, 0.7, 0.0, 0.0, -0.2, 0.3, 0.5)
    old, old_old = BlockVector.like(mf), BlockVector.like(mf)
    A = NavierStokesMatrix(dprm, 0, 1, 0, 1, old, old_old)
    A.initialize(mf, capi.TimeWeights(tw[0], tw[1], tw[2], tw[3], 2.0, -1.0), True)
    O = og.OracleNavierStokesGeneric(mesh, og.INCOMPRESSIBLE, og.NEWTON if linearization == capi.NEWTON else og.PICARD, 1.3, 0.7, -0.2, 0.3, 0.5, tw)
    u, p = seeded_uniform(mesh.u_ids, dim, 1), seeded_uniform(mesh.p_ids, 1, 2)
    uo, uoo = seeded_uniform(mesh.u_ids, dim, 3), seeded_uniform(mesh.u_ids, dim, 4)
    old.block(0).copy_(dev(ctx, uo))
    old_old.block(0).copy_(dev(ctx, uoo))
    O.old_u, O.old_old_u = uo, uoo
    ru, rp = np.zeros(O.n_u), np.zeros(O.n_p)
    O.residual(ru, rp, u, p, np.zeros(O.n_u), np.zeros(O.n_p))
    O.apply_pressure_average_projection(rp)
    rhs, zero = BlockVector.like(mf), BlockVector.like(mf)
    A.residual(rhs, BlockVector(dev(ctx, u), dev(ctx, p)), zero)
    A.apply_pressure_average_projection(rhs.block(1))
    assert rel_err(rhs.block(0).cpu().numpy(), ru) < TOL and rel_err(rhs.block(1).cpu().numpy(), rp) < TOL
    su, sp = seeded_uniform(mesh.u_ids, dim, 11), seeded_uniform(mesh.p_ids, 1, 12)
    ou, op = O.vmult(su, sp)
    dst = BlockVector.like(mf)
    A.vmult(dst, BlockVector(dev(ctx, su), dev(ctx, sp)))
    assert rel_err(dst.block(0).cpu().numpy(), ou) < TOL and rel_err(dst.block(1).cpu().numpy(), op) < TOL
    dv = mf.initialize_dof_vector(0)
    A.velocity_vmult(dv, dev(ctx, su))
    assert rel_err(dv.cpu().numpy(), O.velocity_vmult(su)) < TOL
    # host-pointer entry on the generic path
    hu, hp = np.zeros_like(su), np.zeros_like(sp)
    A.vmult_host(hu, hp, np.ascontiguousarray(su), np.ascontiguousarray(sp))
    assert rel_err(hu, ou) < TOL and rel_err(hp, op) < TOL
    check_sub_blocks_and_diagonal(ctx, mf, A, O, mesh, su, sp, exact_diagonal=True)
    A.clear()
    mf.close()


@pytest.mark.parametrize("k", [2, 3, 4, 5])
@pytest.mark.parametrize("linearization", [capi.NEWTON, capi.PICARD])
def test_hanging_nodes_of_any_degree(ctx, k, linearization):
    """2:1 hanging nodes on Q_k / Q_(k-1) in 2-D through general constraint rows (adaflo_b200_space_desc::hanging_*): the mesh
    type of tests/beltrami_2d (Q4/Q3).  vmult resolves the rows in the read and in the scatter, the residual reads the vector
    entries plainly (read_dof_values_plain) and scatters through the rows; the constrained rows are +-identity."""
    mesh = RefinedTensorMesh2D((4, 3), ((1, 3), (0, 2)), k, lower=(-1.0, -0.5), upper=(1.0, 0.75))
    assert mesh.n_hanging_u > 0 and mesh.n_hanging_p > 0
    mf = MatrixFreeGeneric(ctx, mesh)
    dt = 0.05
    tw = (1.5 / dt, -2.0 / dt, 0.5 / dt, 1.0)
    dprm = capi.FlowParams(capi.INCOMPRESSIBLE, linearization, 0, 0, 1.3, 0.7, 0.0, 0.0, -0.2, 0.3, 0.5)
    old, old_old = BlockVector.like(mf), BlockVector.like(mf)
    A = NavierStokesMatrix(dprm, 0, 1, 0, 1, old, old_old)
    A.initialize(mf, capi.TimeWeights(tw[0], tw[1], tw[2], tw[3], 2.0, -1.0), True)
    O = og.OracleNavierStokesGeneric(mesh, og.INCOMPRESSIBLE, og.NEWTON if linearization == capi.NEWTON else og.PICARD, 1.3, 0.7, -0.2, 0.3, 0.5, tw)
    u, p = seeded_uniform(mesh.u_ids, 2, 1), seeded_uniform(mesh.p_ids, 1, 2)          # NOT distributed: the residual must read plainly
    uo, uoo = seeded_uniform(mesh.u_ids, 2, 3), seeded_uniform(mesh.u_ids, 2, 4)
    old.block(0).copy_(dev(ctx, uo))
    old_old.block(0).copy_(dev(ctx, uoo))
    O.old_u, O.old_old_u = uo, uoo
    ru, rp = np.zeros(O.n_u), np.zeros(O.n_p)
    O.residual(ru, rp, u, p, np.zeros(O.n_u), np.zeros(O.n_p))
    O.apply_pressure_average_projection(rp)
    rhs, zero = BlockVector.like(mf), BlockVector.like(mf)
    A.residual(rhs, BlockVector(dev(ctx, u), dev(ctx, p)), zero)
    A.apply_pressure_average_projection(rhs.block(1))
    assert rel_err(rhs.block(0).cpu().numpy(), ru) < TOL and rel_err(rhs.block(1).cpu().numpy(), rp) < TOL
    su, sp = seeded_uniform(mesh.u_ids, 2, 11), seeded_uniform(mesh.p_ids, 1, 12)
    ou, op = O.vmult(su, sp)
    dst = BlockVector.like(mf)
    A.vmult(dst, BlockVector(dev(ctx, su), dev(ctx, sp)))
    du, dp = dst.block(0).cpu().numpy(), dst.block(1).cpu().numpy()
    assert rel_err(du, ou) < TOL and rel_err(dp, op) < TOL
    cu, cp = O.u_constrained.reshape(-1), O.p_constrained
    assert np.array_equal(du[cu], su[cu])                                               # navier_stokes_matrix.cc:247-256
    assert cp.sum() == mesh.n_hanging_p
    dv = mf.initialize_dof_vector(0)
    A.velocity_vmult(dv, dev(ctx, su))
    assert rel_err(dv.cpu().numpy(), O.velocity_vmult(su)) < TOL
    hu, hp = np.zeros_like(su), np.zeros_like(sp)
    A.vmult_host(hu, hp, np.ascontiguousarray(su), np.ascontiguousarray(sp))
    assert rel_err(hu, ou) < TOL and rel_err(hp, op) < TOL
    # AffineConstraints::distribute on the device
    for space, vec in ((0, u), (1, p)):
        d = dev(ctx, vec.copy())
        mf.distribute_constraints(space, d)
        assert rel_err(d.cpu().numpy(), mesh.distribute(vec.copy(), space)) < 1e-15
    check_sub_blocks_and_diagonal(ctx, mf, A, O, mesh, su, sp, exact_diagonal=False)
    A.clear()
    mf.close()


def test_unsupported_combinations_fail_loudly(ctx):
    mesh = TensorMesh(2, (3, 3), 3)
    mf = MatrixFreeGeneric(ctx, mesh)
    old, old_old = BlockVector.like(mf), BlockVector.like(mf)
    # variable coefficients outside 3-D Q2/Q1
    A = NavierStokesMatrix(capi.FlowParams(capi.INCOMPRESSIBLE, capi.NEWTON, 0, 0, 1.0, 1.0, 0.5, 0.0, 0.0, 0.0, 0.5), 0, 1, 0, 1, old, old_old)
    with pytest.raises(capi.AdafloB200Error):
        A.initialize(mf, capi.TimeWeights(1.0, -1.0, 0.0, 1.0, 1.0, 0.0), True)
    mf.close()
