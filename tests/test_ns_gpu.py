"""GPU parity: CUDA NavierStokesMatrix (through the C ABI) vs the CPU oracle on the
same seeded inputs.  Tolerance: 1e-12 relative in the l2 norm (BASELINE.json
north_star: "match ... within 1e-12 relative (fp64)")."""
import numpy as np
import pytest
import torch

from adaflo_b200 import capi
from adaflo_b200.mesh import BoxMesh, beltrami_velocity, beltrami_pressure, seeded_uniform
from adaflo_b200.operators import BlockVector, MatrixFree, NavierStokesMatrix
from oracle import oracle as orc

from helpers import deformation, flow_params, rel_err, time_weights, variable_coefficients

pytestmark = pytest.mark.gpu
TOL = 1e-12


def dev(ctx, a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(ctx.device)


class Pair:
    """The oracle and the CUDA operator set up on identical data."""

    def __init__(self, ctx, n=(5, 4, 3), geometry="cartesian", var=False, augmented=False, cell_perm_seed=None, **kw):
        self.ctx = ctx
        self.mesh = mesh = BoxMesh(n, augmented=augmented)
        if cell_perm_seed is not None:   # arbitrary cell numbering (uniform Cartesian cells only: same geometry row)
            assert geometry == "cartesian"
            perm = np.random.default_rng(cell_perm_seed).permutation(mesh.n_cells)
            mesh.u.cell_nodes = np.ascontiguousarray(mesh.u.cell_nodes[perm])
            mesh.p.cell_nodes = np.ascontiguousarray(mesh.p.cell_nodes[perm])
        deform = deformation(geometry)
        if var:
            kw.update(density_diff=9.0, viscosity_diff=0.5)
        dprm, oprm = flow_params(augmented=augmented, **kw)
        dtw, otw = time_weights()
        self.mf = MatrixFree(ctx, mesh, geometry, deform)
        self.old = BlockVector.like(self.mf)
        self.old_old = BlockVector.like(self.mf)
        self.A = NavierStokesMatrix(dprm, 0, 1, 0, 1, self.old, self.old_old)
        self.A.initialize(self.mf, dtw, True)
        self.O = orc.OracleNavierStokesMatrix(mesh, oprm, otw, deform, True)
        self.O.density_diff = kw.get("density_diff", 0.0)
        if var:
            rho, mu, damp = variable_coefficients(mesh, deform)
            self.O.rho[:], self.O.muq[:], self.O.damp[:] = rho, mu, damp
            self.A.set_coefficient(capi.COEF_DENSITY, dev(ctx, rho))
            self.A.set_coefficient(capi.COEF_VISCOSITY, dev(ctx, mu))
            self.A.set_coefficient(capi.COEF_DAMPING, dev(ctx, damp))
        # a smooth linearisation point + seeded noise, produced by the residual on both sides
        m = mesh
        xu = m.u.coords
        if deform is not None:
            xu = deform(xu)[0]
        self.u_lin = (beltrami_velocity(xu, 0.1).reshape(-1) + 0.1 * seeded_uniform(m.u.global_ids, 3, 7))
        self.p_lin = seeded_uniform(m.p.global_ids, 1, 8)
        self.u_old = beltrami_velocity(xu, 0.05).reshape(-1) + 0.05 * seeded_uniform(m.u.global_ids, 3, 9)
        self.u_oo = beltrami_velocity(xu, 0.0).reshape(-1) + 0.05 * seeded_uniform(m.u.global_ids, 3, 10)
        self.O.old_u, self.O.old_old_u = self.u_old.copy(), self.u_oo.copy()
        self.old.block(0).copy_(dev(ctx, self.u_old))
        self.old_old.block(0).copy_(dev(ctx, self.u_oo))

    def residual_both(self):
        m = self.mesh
        cu, cp = seeded_uniform(m.u.global_ids, 3, 21), seeded_uniform(m.p.global_ids, 1, 22)   # const_rhs preload
        uu, up = seeded_uniform(m.u.global_ids, 3, 23), seeded_uniform(m.p.global_ids, 1, 24)   # user_rhs
        ru, rp = cu.copy(), cp.copy()
        self.O.residual(ru, rp, self.u_lin, self.p_lin, uu, up)
        rhs = BlockVector(dev(self.ctx, cu), dev(self.ctx, cp))
        self.A.residual(rhs, BlockVector(dev(self.ctx, self.u_lin), dev(self.ctx, self.p_lin)),
                        BlockVector(dev(self.ctx, uu), dev(self.ctx, up)))
        return (ru, rp), (rhs.block(0).cpu().numpy(), rhs.block(1).cpu().numpy())

    def vmult_both(self, seed=31):
        m = self.mesh
        su, sp = seeded_uniform(m.u.global_ids, 3, seed), seeded_uniform(m.p.global_ids, 1, seed + 1)
        ou, op = self.O.vmult(su, sp)
        dst = BlockVector.like(self.mf)
        dst.block(0).fill_(7.0)  # vmult must overwrite
        self.A.vmult(dst, BlockVector(dev(self.ctx, su), dev(self.ctx, sp)))
        return (ou, op), (dst.block(0).cpu().numpy(), dst.block(1).cpu().numpy())

    def close(self):
        self.A.clear()
        self.mf.close()


@pytest.mark.parametrize("geometry", ["cartesian", "affine", "general"])
@pytest.mark.parametrize("linearization", [capi.NEWTON, capi.PICARD, capi.SEMI_IMPLICIT, capi.EXPLICIT])
@pytest.mark.parametrize("var", [False, True])
def test_residual_then_vmult(ctx, geometry, linearization, var):
    P = Pair(ctx, geometry=geometry, var=var, linearization=linearization, damping=-0.2, tau_grad_div=0.3)
    (ru, rp), (gu, gp) = P.residual_both()
    assert rel_err(gu, ru) < TOL and rel_err(gp, rp) < TOL
    if linearization != capi.EXPLICIT:
        lin_gpu = P.A.get_coefficient(capi.COEF_LINEARIZED_VELOCITY).cpu().numpy()
        rows = 12 if linearization == capi.NEWTON else 4
        assert rel_err(lin_gpu[:, :rows], P.O.lin[:, :rows]) < TOL
    (ou, op), (gu, gp) = P.vmult_both()
    assert rel_err(gu, ou) < TOL and rel_err(gp, op) < TOL
    P.close()


@pytest.mark.parametrize("physical_type", [capi.STOKES, capi.INCOMPRESSIBLE_STATIONARY])
@pytest.mark.parametrize("geometry", ["cartesian", "general"])
def test_stokes_and_stationary(ctx, physical_type, geometry):
    P = Pair(ctx, geometry=geometry, physical_type=physical_type, viscosity=0.7)
    (ru, rp), (gu, gp) = P.residual_both()
    assert rel_err(gu, ru) < TOL and rel_err(gp, rp) < TOL
    (ou, op), (gu, gp) = P.vmult_both()
    assert rel_err(gu, ou) < TOL and rel_err(gp, op) < TOL
    P.close()


@pytest.mark.parametrize("geometry", ["cartesian", "general"])
@pytest.mark.parametrize("var", [False, True])
def test_sub_operators(ctx, geometry, var):
    """velocity_vmult, divergence_vmult_add, pressure_{mass,poisson,convdiff}_vmult (rows a4-a8)."""
    P = Pair(ctx, geometry=geometry, var=var, tau_grad_div=0.1)
    P.residual_both()
    m = P.mesh
    su, sp = seeded_uniform(m.u.global_ids, 3, 41), seeded_uniform(m.p.global_ids, 1, 42)
    P.O.fix_linearization_point()
    P.A.fix_linearization_point()
    # velocity block
    du = P.mf.initialize_dof_vector(0)
    P.A.velocity_vmult(du, dev(ctx, su))
    assert rel_err(du.cpu().numpy(), P.O.velocity_vmult(su)) < TOL
    # divergence adds into dst
    for wbv in (False, True):
        pre = seeded_uniform(m.p.global_ids, 1, 43)
        o = pre.copy()
        P.O.divergence_vmult_add(o, su, wbv)
        g = dev(ctx, pre)
        P.A.divergence_vmult_add(g, dev(ctx, su), wbv)
        assert rel_err(g.cpu().numpy(), o) < TOL
    for name in ("pressure_mass_vmult", "pressure_poisson_vmult", "pressure_convdiff_vmult"):
        g = P.mf.initialize_dof_vector(1)
        g.fill_(3.0)
        getattr(P.A, name)(g, dev(ctx, sp))
        o = getattr(P.O, name)(sp)
        assert rel_err(g.cpu().numpy(), o) < TOL, name
        # host-pointer entry (what the Epetra-side wrappers of navier_stokes_preconditioner.cc:421-461 call)
        gh = np.full_like(sp, 3.0)
        getattr(P.A, name + "_host")(gh, np.ascontiguousarray(sp))
        assert rel_err(gh, o) < TOL, name + "_host"
    duh = np.zeros_like(su)
    P.A.velocity_vmult_host(duh, np.ascontiguousarray(su))     # VelocityMatrix::Multiply, :332-343
    assert rel_err(duh, P.O.velocity_vmult(su)) < TOL
    pre = seeded_uniform(m.p.global_ids, 1, 43)
    o = pre.copy()
    P.O.divergence_vmult_add(o, su, True)
    gh = pre.copy()
    P.A.divergence_vmult_add_host(gh, np.ascontiguousarray(su), True)
    assert rel_err(gh, o) < TOL
    P.close()


def test_augmented_taylor_hood(ctx):
    """Q2/(Q1+DG0) pressure and the second null-space mode."""
    P = Pair(ctx, n=(4, 4, 3), augmented=True)
    (ru, rp), (gu, gp) = P.residual_both()
    assert rel_err(gu, ru) < TOL and rel_err(gp, rp) < TOL
    (ou, op), (gu, gp) = P.vmult_both()
    assert rel_err(gu, ou) < TOL and rel_err(gp, op) < TOL
    g = P.mf.initialize_dof_vector(1)
    sp = seeded_uniform(P.mesh.p.global_ids, 1, 51)
    P.A.pressure_mass_vmult(g, dev(ctx, sp))
    assert rel_err(g.cpu().numpy(), P.O.pressure_mass_vmult(sp)) < TOL
    P.close()


def test_tvmult_is_the_transpose(ctx):
    """<Tvmult(x), y> == <x, vmult(y)> (no reference Tvmult exists)."""
    for lin in (capi.NEWTON, capi.PICARD):
        P = Pair(ctx, geometry="general", var=True, linearization=lin, damping=-0.1)
        P.residual_both()
        m = P.mesh
        x = BlockVector(dev(ctx, seeded_uniform(m.u.global_ids, 3, 61)), dev(ctx, seeded_uniform(m.p.global_ids, 1, 62)))
        y = BlockVector(dev(ctx, seeded_uniform(m.u.global_ids, 3, 63)), dev(ctx, seeded_uniform(m.p.global_ids, 1, 64)))
        Ay, Atx = BlockVector.like(P.mf), BlockVector.like(P.mf)
        P.A.vmult(Ay, y)
        P.A.Tvmult(Atx, x)
        lhs = float(Atx.block(0) @ y.block(0) + Atx.block(1) @ y.block(1))
        rhs = float(x.block(0) @ Ay.block(0) + x.block(1) @ Ay.block(1))
        assert abs(lhs - rhs) <= 1e-11 * max(abs(lhs), abs(rhs), 1.0)
        P.close()


def test_ragged_and_tiny_meshes(ctx):
    """cell counts that do not fill a warp triple / a block, and the 1-cell mesh."""
    for n in [(1, 1, 1), (2, 1, 1), (3, 2, 2), (7, 1, 1), (13, 1, 1)]:
        P = Pair(ctx, n=n)
        P.residual_both()
        (ou, op), (gu, gp) = P.vmult_both()
        assert rel_err(gu, ou) < TOL and rel_err(gp, op) < TOL, n
        P.close()


def test_vmult_is_linear_and_deterministic_enough(ctx):
    P = Pair(ctx, n=(6, 6, 6))
    P.residual_both()
    (_, _), (g1u, g1p) = P.vmult_both(seed=71)
    (_, _), (g2u, g2p) = P.vmult_both(seed=71)
    # atomics reorder the adds: repeatability to round-off, not bitwise
    assert rel_err(g1u, g2u) < 1e-14 and rel_err(g1p, g2p) < 1e-14
    P.close()


def test_uninitialized_operator_fails_loudly(ctx):
    A = NavierStokesMatrix(flow_params()[0], 0, 1, 0, 1, None, None)
    with pytest.raises(capi.AdafloB200Error):
        A.vmult(None, None)


def test_host_entry_point_matches_device(ctx):
    P = Pair(ctx, n=(4, 4, 4))
    P.residual_both()
    m = P.mesh
    su, sp = seeded_uniform(m.u.global_ids, 3, 81), seeded_uniform(m.p.global_ids, 1, 82)
    ou, op = P.O.vmult(su, sp)
    hu, hp = torch.from_numpy(su).pin_memory(), torch.from_numpy(sp).pin_memory()
    du, dp = torch.empty_like(hu).pin_memory(), torch.empty_like(hp).pin_memory()
    P.A.vmult_host(du, dp, hu, hp)
    assert rel_err(du.numpy(), ou) < TOL and rel_err(dp.numpy(), op) < TOL
    s, n = P.A.get_matvec_statistics()
    assert n == 1 and s >= 0.0
    P.close()


def test_host_entry_point_pipelined_chunks(ctx):
    """>= 16 decades: vmult_host overlaps upload / cell chunks / download on three streams (16 cell chunks)."""
    P = Pair(ctx, n=(10, 8, 8))
    P.residual_both()
    m = P.mesh
    for seed in (91, 92):   # second call reuses the staging vectors and events
        su, sp = seeded_uniform(m.u.global_ids, 3, seed), seeded_uniform(m.p.global_ids, 1, seed + 10)
        ou, op = P.O.vmult(su, sp)
        hu, hp = torch.from_numpy(su).pin_memory(), torch.from_numpy(sp).pin_memory()
        du, dp = torch.full_like(hu, 7.0).pin_memory(), torch.full_like(hp, 7.0).pin_memory()
        P.A.vmult_host(du, dp, hu, hp)
        assert rel_err(du.numpy(), ou) < TOL and rel_err(dp.numpy(), op) < TOL
        # and the device path afterwards (shared work counter, staging untouched)
        (_, _), (gu, gp) = P.vmult_both(seed=seed)
    P.close()


@pytest.mark.parametrize("var", [False, True])
def test_arbitrary_cell_numbering(ctx, var):
    """randomly permuted cells: nothing in the kernels, the decade blocking or the host pipeline relies on a
    lexicographic cell order (deal.II numbers cells along the p4est space-filling curve)"""
    P = Pair(ctx, n=(6, 5, 7), var=var, cell_perm_seed=3, damping=-0.2, tau_grad_div=0.3)
    (ru, rp), (gu, gp) = P.residual_both()
    assert rel_err(gu, ru) < TOL and rel_err(gp, rp) < TOL
    (ou, op), (gu, gp) = P.vmult_both(seed=41)
    assert rel_err(gu, ou) < TOL and rel_err(gp, op) < TOL
    m = P.mesh
    su, sp = seeded_uniform(m.u.global_ids, 3, 81), seeded_uniform(m.p.global_ids, 1, 82)
    ou, op = P.O.vmult(su, sp)
    hu, hp = torch.from_numpy(su).pin_memory(), torch.from_numpy(sp).pin_memory()
    du, dp = torch.empty_like(hu).pin_memory(), torch.empty_like(hp).pin_memory()
    P.A.vmult_host(du, dp, hu, hp)     # the monotone prefixes degenerate to upload - compute - download
    assert rel_err(du.numpy(), ou) < TOL and rel_err(dp.numpy(), op) < TOL
    P.close()


def test_host_entry_point_pipelined_hanging_nodes(ctx):
    from adaflo_b200.mesh import RefinedBoxMesh
    mesh = RefinedBoxMesh((3, 3, 3), ((1, 3), (0, 2), (1, 2)))
    dprm, oprm = flow_params(damping=-0.2, tau_grad_div=0.3)
    dtw, otw = time_weights()
    mf = MatrixFree(ctx, mesh, "cartesian")
    old, old_old = BlockVector.like(mf), BlockVector.like(mf)
    A = NavierStokesMatrix(dprm, 0, 1, 0, 1, old, old_old)
    A.initialize(mf, dtw, True)
    O = orc.OracleNavierStokesMatrix(mesh, oprm, otw, None, True, mesh.u_lists, mesh.p_lists)
    u0 = beltrami_velocity(mesh.u.coords, 0.1).reshape(-1) + 0.1 * seeded_uniform(mesh.u.global_ids, 3, 5)
    p0 = seeded_uniform(mesh.p.global_ids, 1, 6)
    ru, rp = np.zeros(O.n_u), np.zeros(O.n_p)
    O.residual(ru, rp, u0, p0, np.zeros(O.n_u), np.zeros(O.n_p))
    A.residual(BlockVector.like(mf), BlockVector(dev(ctx, u0), dev(ctx, p0)), BlockVector.like(mf))
    su, sp = seeded_uniform(mesh.u.global_ids, 3, 93), seeded_uniform(mesh.p.global_ids, 1, 94)
    ou, op = O.vmult(su, sp)
    hu, hp = torch.from_numpy(su).pin_memory(), torch.from_numpy(sp).pin_memory()
    du, dp = torch.empty_like(hu).pin_memory(), torch.empty_like(hp).pin_memory()
    A.vmult_host(du, dp, hu, hp)
    assert rel_err(du.numpy(), ou) < TOL and rel_err(dp.numpy(), op) < TOL


@pytest.mark.parametrize("refine", [((1, 3), (0, 2), (1, 2)), ((0, 1), (0, 1), (0, 1)), ((1, 2), (1, 2), (0, 3))])
@pytest.mark.parametrize("var", [False, True])
def test_hanging_nodes(ctx, refine, var):
    """2:1 refined meshes: in-kernel interpolation (compressed masks) vs the oracle's plain constraint
    expansion; residual, vmult, Tvmult identity and the sub-block operators."""
    from adaflo_b200.mesh import RefinedBoxMesh
    mesh = RefinedBoxMesh((3, 3, 3), refine)
    kw = dict(density_diff=9.0, viscosity_diff=0.5) if var else {}
    dprm, oprm = flow_params(damping=-0.2, tau_grad_div=0.3, **kw)
    dtw, otw = time_weights()
    mf = MatrixFree(ctx, mesh, "cartesian")
    old, old_old = BlockVector.like(mf), BlockVector.like(mf)
    A = NavierStokesMatrix(dprm, 0, 1, 0, 1, old, old_old)
    A.initialize(mf, dtw, True)
    O = orc.OracleNavierStokesMatrix(mesh, oprm, otw, None, True, mesh.u_lists, mesh.p_lists)
    if var:
        rho, mu, damp = variable_coefficients(mesh)
        O.rho[:], O.muq[:], O.damp[:] = rho, mu, damp
        A.set_coefficient(capi.COEF_DENSITY, dev(ctx, rho))
        A.set_coefficient(capi.COEF_VISCOSITY, dev(ctx, mu))
        A.set_coefficient(capi.COEF_DAMPING, dev(ctx, damp))
    m = mesh
    u = beltrami_velocity(m.u.coords, 0.1).reshape(-1) + 0.1 * seeded_uniform(m.u.global_ids, 3, 7)
    p = seeded_uniform(m.p.global_ids, 1, 8)
    uo, uoo = seeded_uniform(m.u.global_ids, 3, 9), seeded_uniform(m.u.global_ids, 3, 10)
    O.old_u, O.old_old_u = uo.copy(), uoo.copy()
    old.block(0).copy_(dev(ctx, uo))
    old_old.block(0).copy_(dev(ctx, uoo))
    ru, rp = np.zeros(O.n_u), np.zeros(O.n_p)
    O.residual(ru, rp, u, p, np.zeros(O.n_u), np.zeros(O.n_p))
    rhs = BlockVector.like(mf)
    A.residual(rhs, BlockVector(dev(ctx, u), dev(ctx, p)), BlockVector.like(mf))
    assert rel_err(rhs.block(0).cpu().numpy(), ru) < TOL and rel_err(rhs.block(1).cpu().numpy(), rp) < TOL
    su, sp = seeded_uniform(m.u.global_ids, 3, 31), seeded_uniform(m.p.global_ids, 1, 32)
    ou, op = O.vmult(su, sp)
    dst = BlockVector.like(mf)
    src = BlockVector(dev(ctx, su), dev(ctx, sp))
    A.vmult(dst, src)
    assert rel_err(dst.block(0).cpu().numpy(), ou) < TOL and rel_err(dst.block(1).cpu().numpy(), op) < TOL
    # transpose identity
    y = BlockVector(dev(ctx, seeded_uniform(m.u.global_ids, 3, 63)), dev(ctx, seeded_uniform(m.p.global_ids, 1, 64)))
    Ay, Atx = BlockVector.like(mf), BlockVector.like(mf)
    A.vmult(Ay, y)
    A.Tvmult(Atx, src)
    lhs = float(Atx.block(0) @ y.block(0) + Atx.block(1) @ y.block(1))
    rhs_ = float(src.block(0) @ Ay.block(0) + src.block(1) @ Ay.block(1))
    assert abs(lhs - rhs_) <= 1e-11 * max(abs(lhs), abs(rhs_), 1.0)
    # sub-block operators
    du = mf.initialize_dof_vector(0)
    A.velocity_vmult(du, dev(ctx, su))
    assert rel_err(du.cpu().numpy(), O.velocity_vmult(su)) < TOL
    pre = seeded_uniform(m.p.global_ids, 1, 43)
    o = pre.copy()
    O.divergence_vmult_add(o, su, var)
    g = dev(ctx, pre)
    A.divergence_vmult_add(g, dev(ctx, su), var)
    assert rel_err(g.cpu().numpy(), o) < TOL
    for name in ("pressure_mass_vmult", "pressure_poisson_vmult", "pressure_convdiff_vmult"):
        g = mf.initialize_dof_vector(1)
        getattr(A, name)(g, dev(ctx, sp))
        assert rel_err(g.cpu().numpy(), getattr(O, name)(sp)) < TOL, name
    A.clear()
    mf.close()


@pytest.mark.parametrize("hanging", [False, True])
def test_compute_diagonal(ctx, hanging):
    """diag(vmult) on the velocity block equals e_i^T A e_i of the oracle (sampled dofs); pressure diagonal is 0 / -1"""
    from adaflo_b200.mesh import RefinedBoxMesh
    import ctypes as C
    mesh = RefinedBoxMesh((3, 3, 2), ((1, 3), (0, 2), (0, 1))) if hanging else BoxMesh((4, 3, 3))
    dprm, oprm = flow_params(damping=-0.2, tau_grad_div=0.3)
    dtw, otw = time_weights()
    mf = MatrixFree(ctx, mesh, "cartesian")
    old, old_old = BlockVector.like(mf), BlockVector.like(mf)
    A = NavierStokesMatrix(dprm, 0, 1, 0, 1, old, old_old)
    A.initialize(mf, dtw, True)
    lists = (mesh.u_lists, mesh.p_lists) if hanging else (None, None)
    O = orc.OracleNavierStokesMatrix(mesh, oprm, otw, None, True, *lists)
    m = mesh
    u = beltrami_velocity(m.u.coords, 0.1).reshape(-1)
    p = seeded_uniform(m.p.global_ids, 1, 8)
    O.old_u, O.old_old_u = u.copy(), u.copy()
    old.block(0).copy_(dev(ctx, u))
    old_old.block(0).copy_(dev(ctx, u))
    z = np.zeros
    O.residual(z(O.n_u), z(O.n_p), u, p, z(O.n_u), z(O.n_p))
    A.residual(BlockVector.like(mf), BlockVector(dev(ctx, u), dev(ctx, p)), BlockVector.like(mf))
    du, dp = mf.initialize_dof_vector(0), mf.initialize_dof_vector(1)
    capi.check(ctx.lib.adaflo_b200_ns_compute_diagonal(A.h, C.c_void_p(du.data_ptr()), C.c_void_p(dp.data_ptr())))
    du, dp = du.cpu().numpy(), dp.cpu().numpy()
    assert np.all(dp == 0.0)          # no pressure Dirichlet dofs here
    rng = np.random.default_rng(1)
    sample = rng.choice(O.n_u, size=60, replace=False)
    for i in sample:
        e = np.zeros(O.n_u)
        e[i] = 1.0
        col_u, _ = O.vmult(e, z(O.n_p))
        assert abs(du[i] - col_u[i]) <= 1e-12 * max(1.0, abs(col_u[i])), (i, du[i], col_u[i])
    A.clear()
    mf.close()


def test_device_poiseuille_is_a_discrete_stokes_solution(ctx):
    """analytic pin of the CUDA path (no oracle involved): u = (0.5/nu)(1-y^2) e_x, p = 2 - x lie in Q2/Q1
    (tests/poiseuille.cc:74-112), so the Stokes residual computed on the device vanishes on every row whose test
    function does not touch the open in/outflow faces, and the flow is discretely divergence free"""
    nu = 0.8
    mesh = BoxMesh((4, 4, 2), dirichlet_faces=(2, 3, 4, 5))
    dprm, _ = flow_params(physical_type=capi.STOKES, viscosity=nu)
    dtw, _ = time_weights()
    mf = MatrixFree(ctx, mesh, "cartesian")
    A = NavierStokesMatrix(dprm, 0, 1, 0, 1, BlockVector.like(mf), BlockVector.like(mf))
    A.initialize(mf, dtw, False)
    x = mesh.u.coords
    u = np.zeros((len(x), 3))
    u[:, 0] = 0.5 / nu * (1 - x[:, 1] ** 2)
    p = 2.0 - mesh.p.coords[:, 0]
    rhs = BlockVector.like(mf)
    A.residual(rhs, BlockVector(dev(ctx, u.reshape(-1)), dev(ctx, p)), BlockVector.like(mf))
    ru = rhs.block(0).cpu().numpy().reshape(-1, 3)
    rp = rhs.block(1).cpu().numpy()
    interior = np.abs(np.abs(x[:, 0]) - 1) > 1e-12
    assert np.abs(ru[interior]).max() < 1e-12 and np.abs(rp).max() < 1e-12
    assert np.abs(ru[~interior]).max() > 1e-3      # the traction integral of the exact solution on the open faces
    A.clear()
    mf.close()


def test_amr_cycle_reupload(ctx):
    """What adaflo does after refine_grid (navier_stokes.cc:414-451 runs again: MatrixFree::reinit, operator
    re-initialised): the device mirror and the operator of the old mesh are destroyed, the refined mesh is uploaded on the
    same context, and the solution, transferred by the host (SolutionTransfer), is used as the linearisation point.  Three
    cycles; every cycle is checked against the oracle and the device memory of the destroyed cycles is returned."""
    from adaflo_b200.mesh import RefinedBoxMesh
    dprm, oprm = flow_params(damping=0.1, tau_grad_div=0.2)
    dtw, otw = time_weights()
    free_after = []
    for cycle, mesh in enumerate([BoxMesh((4, 4, 4)), RefinedBoxMesh((4, 4, 4), ((1, 3), (1, 3), (0, 2))),
                                  RefinedBoxMesh((4, 4, 4), ((0, 2), (2, 4), (1, 4)))]):
        mf = MatrixFree(ctx, mesh, "cartesian")
        old, old_old = BlockVector.like(mf), BlockVector.like(mf)
        A = NavierStokesMatrix(dprm, 0, 1, 0, 1, old, old_old)
        A.initialize(mf, dtw, True)
        lists = (mesh.u_lists, mesh.p_lists) if hasattr(mesh, "u_lists") else (None, None)
        O = orc.OracleNavierStokesMatrix(mesh, oprm, otw, None, True, *lists)
        u = beltrami_velocity(mesh.u.coords, 0.1).reshape(-1)  # the "transferred" solution: the same field on every mesh
        p = beltrami_pressure(mesh.p.coords, 0.1).reshape(-1)
        uo, uoo = seeded_uniform(mesh.u.global_ids, 3, 9), seeded_uniform(mesh.u.global_ids, 3, 10)
        O.old_u, O.old_old_u = uo.copy(), uoo.copy()
        old.block(0).copy_(dev(ctx, uo))
        old_old.block(0).copy_(dev(ctx, uoo))
        ru, rp = np.zeros(O.n_u), np.zeros(O.n_p)
        O.residual(ru, rp, u, p, np.zeros(O.n_u), np.zeros(O.n_p))
        rhs = BlockVector.like(mf)
        A.residual(rhs, BlockVector(dev(ctx, u), dev(ctx, p)), BlockVector.like(mf))
        assert rel_err(rhs.block(0).cpu().numpy(), ru) < TOL and rel_err(rhs.block(1).cpu().numpy(), rp) < TOL, cycle
        su, sp = seeded_uniform(mesh.u.global_ids, 3, 31), seeded_uniform(mesh.p.global_ids, 1, 32)
        ou, op = O.vmult(su, sp)
        dst = BlockVector.like(mf)
        A.vmult(dst, BlockVector(dev(ctx, su), dev(ctx, sp)))
        assert rel_err(dst.block(0).cpu().numpy(), ou) < TOL and rel_err(dst.block(1).cpu().numpy(), op) < TOL, cycle
        A.clear()
        mf.close()
        ctx.synchronize()
        free_after.append(torch.cuda.mem_get_info()[0])
    # nothing of a destroyed cycle stays allocated by the library (cudaMalloc'ed mirrors are freed; slack for the work pool)
    assert abs(free_after[2] - free_after[0]) < 64 << 20
