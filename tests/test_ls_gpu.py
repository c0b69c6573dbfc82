"""GPU parity of the level-set operators vs the dense CPU oracle,
tolerance 1e-12 relative (fp64)."""
import numpy as np
import pytest
import torch

from adaflo_b200 import capi
from adaflo_b200.mesh import BoxMesh, beltrami_velocity, seeded_uniform
from adaflo_b200.operators import LevelSetOperators, MatrixFree
from oracle import oracle as orc

from helpers import deformation, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-12


def dev(ctx, a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(ctx.device)


class LSPair:
    def __init__(self, ctx, n=(4, 3, 3), n_sub=3, geometry="cartesian", dirichlet_faces=()):
        self.ctx = ctx
        self.mesh = mesh = BoxMesh(n, lower=(-1, -1, -1), upper=(1.0, 0.7, 0.5))
        mesh.build_ls_space(n_sub, dirichlet_faces)
        deform = deformation(geometry)
        self.mf = MatrixFree(ctx, mesh, geometry, deform)
        h_min = float(mesh.h.min())
        eps, eps_used = 1.5, 1.5 * h_min / n_sub
        self.dprm = capi.LSParams(3, n_sub, 0, 0, eps, eps_used, 4.0, h_min)
        self.oprm = orc.LSParams(eps, eps_used, 4.0, h_min, 0.02, 3, 0)
        cd = mesh.cell_diameters()
        self.A = LevelSetOperators(self.mf, self.dprm, cd)
        self.gamma, self.w_old, self.w_oo, self.dt = 75.0, -100.0, 25.0, 0.02
        self.A.set_time(self.gamma, self.w_old, self.w_oo, self.dt)
        self.O = orc.OracleLevelSet(mesh, self.oprm, deform, general=(geometry == "general"))
        x = mesh.ls.coords if deform is None else deform(mesh.ls.coords)[0]
        r = np.linalg.norm(x - np.array([0.1, -0.05, -0.2]), axis=1)
        self.c = -np.tanh((r - 0.45) / (2 * 0.12)) + 0.01 * seeded_uniform(mesh.ls.global_ids, 1, 5)
        self.n = mesh.ls.n_owned
        # Jacobi diagonal for constrained rows
        self.diag = self.O.mass_diagonal()
        self.O.diag = self.diag
        self.A.set_diagonal(dev(ctx, self.diag))

    def vec(self, seed):
        return seeded_uniform(self.mesh.ls.global_ids, 1, seed)

    def close(self):
        self.A.close()
        self.mf.close()


@pytest.mark.parametrize("geometry", ["cartesian", "affine", "general"])
@pytest.mark.parametrize("n_sub", [3, 4])
@pytest.mark.parametrize("dirichlet", [(), (0, 3)])
def test_level_set_operators(ctx, geometry, n_sub, dirichlet):
    P = LSPair(ctx, n_sub=n_sub, geometry=geometry, dirichlet_faces=dirichlet)
    A, O, n = P.A, P.O, P.n
    # mass diagonal (a18)
    assert rel_err(A.initialize_mass_matrix_diagonal().cpu().numpy(), P.diag) < TOL
    # advection: rhs writes the convection velocity, vmult uses it (a11, a12)
    m = P.mesh
    vel = beltrami_velocity(m.u.coords, 0.0).reshape(-1) + 0.1 * seeded_uniform(m.u.global_ids, 3, 6)
    c_old, c_oo = P.c + 0.05 * P.vec(7), P.c + 0.05 * P.vec(8)
    pre = P.vec(9)
    o = pre.copy()
    O.advance_concentration_rhs(o, vel, P.c, c_old, c_oo, P.gamma, P.w_old, P.w_oo)
    g = dev(ctx, pre)
    A.advance_concentration_rhs(g, dev(ctx, vel), dev(ctx, P.c), dev(ctx, c_old), dev(ctx, c_oo))
    assert rel_err(g.cpu().numpy(), o) < TOL
    assert rel_err(A.pointdata_canonical(0).cpu().numpy(), O.conv) < TOL
    src = P.vec(10)
    g = P.mf.initialize_dof_vector(2)
    g.fill_(5.0)
    A.advance_concentration_vmult(g, dev(ctx, src))
    assert rel_err(g.cpu().numpy(), O.advance_concentration_vmult(src, P.gamma)) < TOL
    # normal (a13)
    rhs_o = [np.zeros(n) for _ in range(3)]
    O.compute_normal_rhs(rhs_o, P.c)
    rhs_g = [P.mf.initialize_dof_vector(2) for _ in range(3)]
    A.compute_normal_rhs(rhs_g, dev(ctx, P.c))
    for d in range(3):
        assert rel_err(rhs_g[d].cpu().numpy(), rhs_o[d]) < TOL
    nsrc = [P.vec(11 + d) for d in range(3)]
    ndst = [P.mf.initialize_dof_vector(2) for _ in range(3)]
    A.compute_normal_vmult(ndst, [dev(ctx, v) for v in nsrc])
    for d, ref in enumerate(O.compute_normal_vmult(nsrc)):
        assert rel_err(ndst[d].cpu().numpy(), ref) < TOL
    # curvature (a14): a smooth normal field = projected gradient stand-in
    x = m.ls.coords
    normal = [x[:, d] - [0.1, -0.05, -0.2][d] + 0.02 * P.vec(20 + d) for d in range(3)]
    normal[0][:5] = 0.0
    normal[1][:5] = 0.0
    normal[2][:5] = 1e-3     # |n| <= 1e-2 -> normalised to zero
    o = np.zeros(n)
    O.compute_curvature_rhs(o, normal)
    g = P.mf.initialize_dof_vector(2)
    A.compute_curvature_rhs(g, [dev(ctx, v) for v in normal])
    assert rel_err(g.cpu().numpy(), o) < TOL
    for diff in (True, False):
        g = P.mf.initialize_dof_vector(2)
        A.compute_curvature_vmult(g, dev(ctx, src), diff)
        assert rel_err(g.cpu().numpy(), O.compute_curvature_vmult(src, diff)) < TOL
    curv = 3.0 * P.vec(30)
    co = curv.copy()
    O.curvature_correction(co, P.c)
    cg = dev(ctx, curv)
    A.curvature_correction(cg, dev(ctx, P.c))
    assert rel_err(cg.cpu().numpy(), co) < TOL
    # reinitialisation (a15): first step freezes the normals, later steps reuse them
    for first in (True, False):
        o = np.zeros(n)
        O.reinitialization_rhs(o, P.c, normal, False, first)
        g = P.mf.initialize_dof_vector(2)
        A.reinitialization_rhs(g, dev(ctx, P.c), [dev(ctx, v) for v in normal], False, first)
        assert rel_err(g.cpu().numpy(), o) < TOL
    assert rel_err(A.pointdata_canonical(1).cpu().numpy(), O.normal_q) < TOL
    o = np.zeros(n)
    O.reinitialization_rhs(o, P.c, None, True, False)
    g = P.mf.initialize_dof_vector(2)
    A.reinitialization_rhs(g, dev(ctx, P.c), None, True, False)
    assert rel_err(g.cpu().numpy(), o) < TOL
    for diffuse_only in (False, True):
        g = P.mf.initialize_dof_vector(2)
        A.reinitialization_vmult(g, dev(ctx, src), diffuse_only)
        assert rel_err(g.cpu().numpy(), O.reinitialization_vmult(src, diffuse_only)) < TOL
    P.close()


def test_level_set_vmults_are_symmetric_where_the_weak_form_is(ctx):
    """normal / curvature / reinit operators are symmetric: <A x, y> = <x, A y>"""
    P = LSPair(ctx)
    x, y = dev(ctx, P.vec(41)), dev(ctx, P.vec(42))
    for f in (lambda d, s: P.A.compute_curvature_vmult(d, s, True), lambda d, s: P.A.reinitialization_vmult(d, s, True)):
        ax, ay = P.mf.initialize_dof_vector(2), P.mf.initialize_dof_vector(2)
        f(ax, x)
        f(ay, y)
        assert abs(float(ax @ y) - float(x @ ay)) < 1e-11 * abs(float(ax @ y))
    P.close()


@pytest.mark.parametrize("geometry", ["cartesian", "general"])
@pytest.mark.parametrize("interpolate", [False, True])
def test_heaviside_and_force(ctx, geometry, interpolate):
    """compute_heaviside + local_compute_force (row a16): Heaviside, rho_q/mu_q written into the NS operator, force vector"""
    from adaflo_b200.operators import BlockVector, NavierStokesMatrix
    from helpers import flow_params, time_weights
    n_sub = 3
    mesh = BoxMesh((4, 3, 3))
    mesh.build_ls_space(n_sub)
    deform = deformation(geometry)
    mf = MatrixFree(ctx, mesh, geometry, deform)
    dprm, _ = flow_params(density=1.0, viscosity=0.1, density_diff=9.0, viscosity_diff=0.4)
    dtw, _ = time_weights()
    A = NavierStokesMatrix(dprm, 0, 1, 0, 1, BlockVector.like(mf), BlockVector.like(mf))
    A.initialize(mf, dtw, True)
    eps = 1.5
    L = LevelSetOperators(mf, capi.LSParams(3, n_sub, 0, 0, eps, eps * float(mesh.h.min()) / n_sub, 4.0, float(mesh.h.min())), mesh.cell_diameters())
    x = mesh.ls.coords
    r = np.linalg.norm(x - np.array([0.1, -0.05, -0.2]), axis=1)
    c = -np.tanh((r - 0.5) / (2 * 0.08))
    H_ref = orc.compute_heaviside(mesh.ls.cell_nodes.astype(np.int64), c, eps, n_sub)
    Hg = mf.initialize_dof_vector(2)
    L.compute_heaviside(Hg, dev(ctx, c))
    assert np.abs(Hg.cpu().numpy() - H_ref).max() < 1e-14
    assert 0.0 < H_ref.mean() < 1.0 and ((H_ref > 0) & (H_ref < 1)).sum() > 10
    kappa = 2.0 / np.maximum(r, 0.2) + 0.1 * seeded_uniform(mesh.ls.global_ids, 1, 3)
    interp = None
    if interpolate:  # nodal interpolation onto the pressure space (level_set_base.cc:65-137): corner values
        n1 = n_sub + 1
        interp = np.zeros((8, n1 ** 3))
        for i in range(8):
            interp[i, (n_sub * (i & 1)) + n1 * ((n_sub * ((i >> 1) & 1)) + n1 * (n_sub * (i >> 2)))] = 1.0
    pre = seeded_uniform(mesh.u.global_ids, 3, 4)
    o = pre.copy()
    rho_q, mu_q = orc.compute_force(mesh, H_ref, kappa, 0.07, 0.98, 1.0, 9.0, 0.1, 0.4, True, interp, o, deform)
    g = dev(ctx, pre)
    L.compute_force(A, g, dev(ctx, H_ref), dev(ctx, kappa), 0.07, 0.98, interp)
    assert rel_err(g.cpu().numpy(), o) < TOL
    assert rel_err(A.get_coefficient(capi.COEF_DENSITY).cpu().numpy(), rho_q) < TOL
    assert rel_err(A.get_coefficient(capi.COEF_VISCOSITY).cpu().numpy(), mu_q) < TOL
    L.close()
    A.clear()
    mf.close()


@pytest.mark.parametrize("refine", [((1, 3), (0, 2), (0, 1)), ((0, 1), (0, 1), (0, 1))])
def test_level_set_operators_with_hanging_nodes(ctx, refine):
    """FE_Q_iso_Q1(3) on a 2:1 refined mesh: on-the-fly constraint resolution in the sub-cell kernels vs the
    oracle's plain expansion lists"""
    from adaflo_b200.mesh import RefinedBoxMesh
    n_sub = 3
    mesh = RefinedBoxMesh((3, 3, 2), refine)
    mesh.build_ls_space(n_sub, dirichlet_faces=(0,))
    mf = MatrixFree(ctx, mesh, "cartesian")
    h_min = float(mesh.h_cell.min())
    eps, eps_used = 1.5, 1.5 * h_min / n_sub
    A = LevelSetOperators(mf, capi.LSParams(3, n_sub, 0, 0, eps, eps_used, 4.0, h_min), mesh.cell_diameters(), dof_index_vel=0)
    A.set_time(75.0, -100.0, 25.0, 0.02)
    O = orc.OracleLevelSet(mesh, orc.LSParams(eps, eps_used, 4.0, h_min, 0.02, 3, 0))
    n = mesh.ls.n_owned
    # advection rhs with a velocity space that has hanging nodes (read_dof_values_plain of a distributed vector):
    # the oracle gets the plainly expanded nodal values, one private entry per (cell, node)
    vel = beltrami_velocity(mesh.u.coords, 0.0).reshape(-1) + 0.1 * seeded_uniform(mesh.u.global_ids, 3, 6)
    ptr, idx, wts = mesh.u_lists
    expanded = np.add.reduceat(wts * vel[idx], ptr[:-1]).reshape(mesh.n_cells, 3, 27)   # [cell][comp][node]
    O.cell_u = np.arange(mesh.n_cells * 27, dtype=np.int64).reshape(mesh.n_cells, 27)
    vel_ext = np.ascontiguousarray(expanded.transpose(0, 2, 1).reshape(-1))              # [(cell, node)][comp]
    x = mesh.ls.coords
    c_adv = -np.tanh((np.linalg.norm(x - np.array([0.1, -0.05, -0.2]), axis=1) - 0.45) / 0.3)
    c_old, c_oo = c_adv + 0.05 * seeded_uniform(mesh.ls.global_ids, 1, 7), c_adv + 0.05 * seeded_uniform(mesh.ls.global_ids, 1, 8)
    pre = seeded_uniform(mesh.ls.global_ids, 1, 9)
    o = pre.copy()
    O.advance_concentration_rhs(o, vel_ext, c_adv, c_old, c_oo, 75.0, -100.0, 25.0)
    g = dev(ctx, pre)
    A.advance_concentration_rhs(g, dev(ctx, vel), dev(ctx, c_adv), dev(ctx, c_old), dev(ctx, c_oo))
    assert rel_err(g.cpu().numpy(), o) < TOL
    assert rel_err(A.pointdata_canonical(0).cpu().numpy(), O.conv) < TOL
    diag = O.mass_diagonal()
    O.diag = diag
    A.set_diagonal(dev(ctx, diag))
    assert rel_err(A.initialize_mass_matrix_diagonal().cpu().numpy(), diag) < TOL
    x = mesh.ls.coords
    r = np.linalg.norm(x - np.array([0.1, -0.05, -0.2]), axis=1)
    c = -np.tanh((r - 0.45) / 0.3)
    src = seeded_uniform(mesh.ls.global_ids, 1, 10)
    g = mf.initialize_dof_vector(2)
    A.advance_concentration_vmult(g, dev(ctx, src))
    assert rel_err(g.cpu().numpy(), O.advance_concentration_vmult(src, 75.0)) < TOL
    rhs_o = [np.zeros(n) for _ in range(3)]
    O.compute_normal_rhs(rhs_o, c)
    rhs_g = [mf.initialize_dof_vector(2) for _ in range(3)]
    A.compute_normal_rhs(rhs_g, dev(ctx, c))
    for d in range(3):
        assert rel_err(rhs_g[d].cpu().numpy(), rhs_o[d]) < TOL
    normal = [x[:, d] - [0.1, -0.05, -0.2][d] + 0.02 * seeded_uniform(mesh.ls.global_ids, 1, 20 + d) for d in range(3)]
    ndst = [mf.initialize_dof_vector(2) for _ in range(3)]
    A.compute_normal_vmult(ndst, [dev(ctx, v) for v in normal])
    for d, ref in enumerate(O.compute_normal_vmult(normal)):
        assert rel_err(ndst[d].cpu().numpy(), ref) < TOL
    o = np.zeros(n)
    O.compute_curvature_rhs(o, normal)
    g = mf.initialize_dof_vector(2)
    A.compute_curvature_rhs(g, [dev(ctx, v) for v in normal])
    assert rel_err(g.cpu().numpy(), o) < TOL
    g = mf.initialize_dof_vector(2)
    A.compute_curvature_vmult(g, dev(ctx, src), True)
    assert rel_err(g.cpu().numpy(), O.compute_curvature_vmult(src, True)) < TOL
    o = np.zeros(n)
    O.reinitialization_rhs(o, c, normal, False, True)
    g = mf.initialize_dof_vector(2)
    A.reinitialization_rhs(g, dev(ctx, c), [dev(ctx, v) for v in normal], False, True)
    assert rel_err(g.cpu().numpy(), o) < TOL
    for diffuse_only in (False, True):
        g = mf.initialize_dof_vector(2)
        A.reinitialization_vmult(g, dev(ctx, src), diffuse_only)
        assert rel_err(g.cpu().numpy(), O.reinitialization_vmult(src, diffuse_only)) < TOL
    A.close()
    mf.close()


@pytest.mark.parametrize("refine", [((1, 3), (0, 2), (0, 1)), ((0, 1), (0, 1), (0, 1))])
@pytest.mark.parametrize("interpolate", [False, True])
def test_heaviside_and_force_with_hanging_nodes(ctx, refine, interpolate):
    """compute_heaviside / local_compute_force read the Heaviside function PLAINLY (level_set_okz.cc:341,497-536), so the
    hanging dofs of the level-set space carry their own values H(interpolated c), last cell wins.  The oracle runs
    on an extended dof space (real dofs + one dof per hanging lattice point) built from the mesh's plain expansion
    lists; the velocity test functions are distributed with the plain velocity lists."""
    from adaflo_b200.mesh import RefinedBoxMesh
    from adaflo_b200.operators import BlockVector, NavierStokesMatrix
    from helpers import flow_params, time_weights
    n_sub = 3
    mesh = RefinedBoxMesh((3, 3, 2), refine)
    mesh.build_ls_space(n_sub)
    mf = MatrixFree(ctx, mesh, "cartesian")
    dprm, _ = flow_params(density=1.0, viscosity=0.1, density_diff=9.0, viscosity_diff=0.4)
    dtw, _ = time_weights()
    A = NavierStokesMatrix(dprm, 0, 1, 0, 1, BlockVector.like(mf), BlockVector.like(mf))
    A.initialize(mf, dtw, True)
    h_min = float(mesh.h_cell.min())
    eps = 1.5
    L = LevelSetOperators(mf, capi.LSParams(3, n_sub, 0, 0, eps, eps * h_min / n_sub, 4.0, h_min), mesh.cell_diameters())
    n, N = mesh.ls.n_owned, (n_sub + 1) ** 3
    # ---- extended dof space from the plain expansion lists
    ptr, idx, wts = mesh.ls_lists
    ext = np.zeros((mesh.n_cells, N), dtype=np.int64)
    hanging = {}
    for c in range(mesh.n_cells):
        for j in range(N):
            k = c * N + j
            a, b = ptr[k], ptr[k + 1]
            if b - a == 1:
                ext[c, j] = idx[a]
            else:
                key = tuple(sorted(idx[a:b].tolist()))
                if key not in hanging:
                    hanging[key] = (len(hanging), idx[a:b].copy(), wts[a:b].copy())
                ext[c, j] = n + hanging[key][0]
    assert len(hanging) > 0

    def extend(v):
        out = np.zeros(n + len(hanging))
        out[:n] = v
        for hid, ii, ww in hanging.values():
            out[n + hid] = float(ww @ v[ii])
        return out

    x = mesh.ls.coords
    r = np.linalg.norm(x - np.array([0.1, -0.05, -0.2]), axis=1)
    c = -np.tanh((r - 0.5) / (2 * 0.08))
    H_ext = orc.compute_heaviside(ext, extend(c), eps, n_sub)
    Hg = mf.initialize_dof_vector(2)
    L.compute_heaviside(Hg, dev(ctx, c))
    assert np.abs(Hg.cpu().numpy() - H_ext[:n]).max() < 1e-14
    assert ((H_ext > 0) & (H_ext < 1)).sum() > 10
    kappa = 2.0 / np.maximum(r, 0.2) + 0.1 * seeded_uniform(mesh.ls.global_ids, 1, 3)
    interp = None
    if interpolate:
        n1 = n_sub + 1
        interp = np.zeros((8, n1 ** 3))
        for i in range(8):
            interp[i, (n_sub * (i & 1)) + n1 * ((n_sub * ((i >> 1) & 1)) + n1 * (n_sub * (i >> 2)))] = 1.0
    # ---- oracle force: private velocity entries per (cell, node), then the plain distribution
    K, jxw = mesh.general_geometry(3, None)
    K, jxw = np.ascontiguousarray(K), np.ascontiguousarray(jxw)
    cu = np.arange(mesh.n_cells * 27, dtype=np.int64).reshape(mesh.n_cells, 27)
    local = np.zeros(mesh.n_cells * 27 * 3)
    rho_q, mu_q = np.zeros((mesh.n_cells, 27)), np.zeros((mesh.n_cells, 27))
    ip = None if interp is None else np.ascontiguousarray(interp)
    import ctypes as C
    orc.ls_lib().ls_oracle_force(mesh.n_cells, n_sub, orc._p(np.ascontiguousarray(ext)), orc._p(cu), None, orc._p(K), orc._p(jxw),
                                 orc._p(H_ext), orc._p(extend(kappa)), C.c_double(0.07), C.c_double(0.98), C.c_double(1.0),
                                 C.c_double(9.0), C.c_double(0.1), C.c_double(0.4), 1, orc._p(ip), orc._p(local), orc._p(rho_q),
                                 orc._p(mu_q))
    pre = seeded_uniform(mesh.u.global_ids, 3, 4)
    o = pre.copy()
    uptr, uidx, uw = mesh.u_lists                      # per (cell, comp, node): entries node * 3 + comp and weights
    loc = local.reshape(mesh.n_cells, 27, 3).transpose(0, 2, 1).reshape(-1)   # -> (cell, comp, node)
    con = mesh.u.dirichlet.reshape(-1)
    contrib = np.repeat(loc, np.diff(uptr)) * uw
    keep = ~con[uidx]
    np.add.at(o, uidx[keep], contrib[keep])
    g = dev(ctx, pre)
    L.compute_force(A, g, Hg, dev(ctx, kappa), 0.07, 0.98, interp)
    assert rel_err(g.cpu().numpy(), o) < TOL
    assert rel_err(A.get_coefficient(capi.COEF_DENSITY).cpu().numpy(), rho_q) < TOL
    assert rel_err(A.get_coefficient(capi.COEF_VISCOSITY).cpu().numpy(), mu_q) < TOL
    L.close()
    A.clear()
    mf.close()
