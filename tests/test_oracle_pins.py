"""Pins of the CPU oracle against the reference's own known answers (CPU only).

The reference holds no operator-level golden vectors; what it does
hold are end-to-end logs and analytic solutions:
  * tests/beltrami_3d.output:5-7  -- interpolation errors of the Beltrami field on the 16^3 mesh
  * tests/beltrami_3d.output:13   -- norms of the first nonlinear residual (pure output of
                                     NavierStokesMatrix::residual + mean projection)
  * tests/poiseuille.cc:74-112    -- steady Poiseuille flow, exact in the Q2/Q1 space
plus the defining property vmult == Jacobian of residual, which ties the vmult
restatement to the pinned residual.
"""
import numpy as np
import pytest

from adaflo_b200 import capi
from adaflo_b200.mesh import BoxMesh, beltrami_pressure, beltrami_velocity, seeded_uniform, sine_deformation
from oracle import oracle as orc

from helpers import flow_params, time_weights, variable_coefficients

import json
import os

with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "beltrami_3d_output.json")) as _f:
    GOLDEN = json.load(_f)   # numbers of the reference's golden log tests/beltrami_3d.output


@pytest.fixture(scope="module")
def beltrami16():
    mesh = BoxMesh((16, 16, 16))
    _, oprm = flow_params()  # Newton, beta = 0.5 (skew-symmetric), nu = rho = 1: tests/beltrami_3d.prm
    dt = 0.05
    # time step #1 of BDF-2 uses implicit Euler weights and no extrapolation (time_stepping.cc:158-198)
    tw = orc.TimeWeights(1 / dt, -1 / dt, 0.0, 1.0, 1.0, 0.0)
    return mesh, orc.OracleNavierStokesMatrix(mesh, oprm, tw), dt


def test_dof_counts_match_golden_log(beltrami16):
    mesh, _, _ = beltrami16
    # tests/beltrami_3d.output:2-3
    assert mesh.n_cells == 4096
    assert mesh.u.n_dofs_owned == 107811 and mesh.p.n_dofs_owned == 4913


def test_interpolation_errors_match_golden_log(beltrami16):
    mesh, ns, _ = beltrami16
    u0 = beltrami_velocity(mesh.u.coords, 0.0).reshape(-1)
    p0 = beltrami_pressure(mesh.p.coords, 0.0).reshape(-1)
    xq = mesh.quadrature_points(4)  # QGauss(v_degree + 2), tests/beltrami.cc:254
    e_u = ns.l2_error(3, u0, beltrami_velocity(xq, 0.0), n_q_1d=4)
    e_p = ns.l2_error(1, p0, beltrami_pressure(xq, 0.0)[..., None], n_q_1d=4)
    # tests/beltrami_3d.output:5 prints 4 significant digits
    assert f"{e_p:.4g}" == GOLDEN["l2_errors_t0"]["e_p"] == "0.02383"
    assert f"{e_u:.4g}" == GOLDEN["l2_errors_t0"]["e_u"] == "0.0001993"


def test_first_nonlinear_residual_matches_golden_log(beltrami16):
    mesh, ns, dt = beltrami16
    u0 = beltrami_velocity(mesh.u.coords, 0.0).reshape(-1)
    p0 = beltrami_pressure(mesh.p.coords, 0.0).reshape(-1)
    # init_time_advance (navier_stokes.cc:659-685): solution = extrapolate(cur, old) = cur in step 1,
    # then Dirichlet values of the new time level (apply_boundary_conditions, :1216-1252)
    u = u0.copy()
    d = mesh.u.dirichlet.reshape(-1)
    u[d] = beltrami_velocity(mesh.u.coords, dt).reshape(-1)[d]
    ns.old_u, ns.old_old_u = u0.copy(), np.zeros_like(u0)
    ru, rp = np.zeros(ns.n_u), np.zeros(ns.n_p)          # system_rhs = const_rhs = 0
    ns.residual(ru, rp, u, p0, np.zeros(ns.n_u), np.zeros(ns.n_p))
    ns.apply_pressure_average_projection(rp)              # compute_residual, navier_stokes.cc:781-790
    # tests/beltrami_3d.output:13: "2.590e+00   6.423e-02"
    assert f"{np.linalg.norm(ru):.3e}" == GOLDEN["first_nonlinear_residual"]["1"]["u"] == "2.590e+00"
    assert f"{np.linalg.norm(rp):.3e}" == GOLDEN["first_nonlinear_residual"]["1"]["p"] == "6.423e-02"


def test_poiseuille_is_a_discrete_stokes_solution():
    """u = (0.5/nu)(1-y^2) e_x, p = 2 - x lie in Q2/Q1 (tests/poiseuille.cc:74-112): the Stokes
    residual vanishes on all rows whose test function does not touch the open boundary."""
    nu = 0.8
    mesh = BoxMesh((4, 4, 2), dirichlet_faces=(2, 3, 4, 5))
    _, oprm = flow_params(physical_type=capi.STOKES, viscosity=nu)
    _, tw = time_weights()
    ns = orc.OracleNavierStokesMatrix(mesh, oprm, tw, pressure_average_fix=False)
    x = mesh.u.coords
    u = np.zeros((len(x), 3))
    u[:, 0] = 0.5 / nu * (1 - x[:, 1] ** 2)
    p = 2.0 - mesh.p.coords[:, 0]
    ru, rp = np.zeros(ns.n_u), np.zeros(ns.n_p)
    ns.residual(ru, rp, u.reshape(-1), p, np.zeros(ns.n_u), np.zeros(ns.n_p))
    interior = (np.abs(np.abs(x[:, 0]) - 1) > 1e-12)       # rows away from the in/outflow faces
    ru = ru.reshape(-1, 3)
    assert np.abs(ru[interior]).max() < 1e-12
    assert np.abs(rp).max() < 1e-12                          # divergence free everywhere
    # on the open faces the residual is the traction integral -int (p - nu du/dn.n) v.n of the exact solution
    assert np.abs(ru[~interior]).max() > 1e-3


@pytest.mark.parametrize("lin", [capi.NEWTON])
@pytest.mark.parametrize("var", [False, True])
def test_vmult_is_the_jacobian_of_the_residual(lin, var):
    mesh = BoxMesh((3, 3, 2))
    deform = sine_deformation(0.04)
    kw = dict(density_diff=9.0, viscosity_diff=0.5) if var else {}
    _, oprm = flow_params(linearization=lin, damping=-0.2, tau_grad_div=0.3, **kw)
    _, tw = time_weights()
    ns = orc.OracleNavierStokesMatrix(mesh, oprm, tw, deform)
    if var:
        ns.rho[:], ns.muq[:], ns.damp[:] = variable_coefficients(mesh, deform)
    u = beltrami_velocity(deform(mesh.u.coords)[0], 0.1).reshape(-1)
    p = seeded_uniform(mesh.p.global_ids, 1, 3)
    du = seeded_uniform(mesh.u.global_ids, 3, 4)
    dp = seeded_uniform(mesh.p.global_ids, 1, 5)
    # the Jacobian acts on the unconstrained space
    du[mesh.u.dirichlet.reshape(-1)] = 0.0
    ns.old_u = seeded_uniform(mesh.u.global_ids, 3, 6)
    ns.old_old_u = seeded_uniform(mesh.u.global_ids, 3, 7)
    zero_u, zero_p = np.zeros(ns.n_u), np.zeros(ns.n_p)

    def R(uu, pp):
        ru, rp = np.zeros(ns.n_u), np.zeros(ns.n_p)
        ns.residual(ru, rp, uu, pp, zero_u, zero_p)
        return -ru, -rp       # residual() flips the sign (navier_stokes_matrix.cc:288-292)

    eps = 1e-6
    rpu, rpp = R(u + eps * du, p + eps * dp)
    rmu, rmp = R(u - eps * du, p - eps * dp)
    R(u, p)                   # leaves the linearisation point at (u, p)
    au, ap = np.zeros(ns.n_u), np.zeros(ns.n_p)
    ns._loop(orc.OP_VMULT, du, dp, au, ap)
    fd_u, fd_p = (rpu - rmu) / (2 * eps), (rpp - rmp) / (2 * eps)
    assert np.linalg.norm(fd_u - au) / np.linalg.norm(au) < 1e-7
    assert np.linalg.norm(fd_p - ap) / np.linalg.norm(ap) < 1e-7


def test_oracle_shape_tables_are_consistent():
    """partition of unity / derivative sums, and the literals hard-wired in the kernels."""
    for deg, nq in [(2, 3), (1, 3), (1, 2)]:
        v, d = orc.shape_1d(deg, nq)
        assert np.allclose(v.sum(axis=1), 1.0, atol=1e-15) and np.allclose(d.sum(axis=1), 0.0, atol=1e-14)
    v, d = orc.shape_1d(2, 3)
    assert v[0, 0] == 0.6872983346207416 and d[0, 1] == 3.0983866769659336
    x, w = orc.gauss01(3)
    assert abs(w.sum() - 1) < 1e-15 and w[1] == 0.4444444444444444


# ------------------------------------------------------------------------------------------------------------------
# tests/beltrami_3d_augp.output: the same flow on a locally refined mesh (2:1 hanging nodes) with augmented Taylor-Hood
# elements Q2/(Q1+DG0) -- pins the hanging-node constraints, the element-constant pressure and the second null-space mode
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "beltrami_3d_augp_output.json")) as _f:
    GOLDEN_AUGP = json.load(_f)


def augp_mesh():
    from adaflo_b200.mesh import RefinedBoxMesh
    # tests/beltrami.cc:385-412, see the "_mesh" entry of the fixture
    return RefinedBoxMesh((8, 8, 8), ((4, 8), (0, 2), (0, 2)), augmented=True)


def count_hanging(lists):
    ptr, idx, w = lists
    keys = set()
    for k in range(len(ptr) - 1):
        if ptr[k + 1] - ptr[k] > 1:
            keys.add((tuple(idx[ptr[k]:ptr[k + 1]].tolist()), tuple(np.round(w[ptr[k]:ptr[k + 1]], 12).tolist())))
    return len(keys)


def test_augp_mesh_and_dof_counts_match_golden_log():
    mesh = augp_mesh()
    assert mesh.n_cells == GOLDEN_AUGP["n_active_cells"]["value"]
    # deal.II counts constrained dofs: vertices of fine cells on coarse edges / faces (= the hanging Q1 nodes) carry
    # their own Q2 dofs as well, identified with the coarse edge / face dofs by constraints
    hanging_p = count_hanging(mesh.p_lists)
    hanging_u = count_hanging(mesh.u_lists) // 3
    assert mesh.p.n_owned + hanging_p == GOLDEN_AUGP["n_dofs"]["pressure"]
    assert 3 * (mesh.u.n_owned + hanging_u + hanging_p) == GOLDEN_AUGP["n_dofs"]["velocity"]


def test_augp_interpolation_errors_match_golden_log():
    """VectorTools::interpolate gives hanging nodes the exact values and the error is integrated before any
    constraint is distributed: plain nodal interpolation on every cell, DG0 part zero (navier_stokes.cc:1414-1430)"""
    import ctypes as C
    mesh = augp_mesh()
    nc = mesh.n_cells

    def local_coords(k):
        a = np.arange(k + 1) / k
        lz, ly, lx = np.meshgrid(a, a, a, indexing="ij")
        loc = np.stack([lx.reshape(-1), ly.reshape(-1), lz.reshape(-1)], axis=1)
        return (mesh.origin[:, None, :] + mesh.h_cell[:, None, :] * loc[None, :, :]).reshape(-1, 3)
    uvals = beltrami_velocity(local_coords(2), 0.0)
    pvals = np.concatenate([beltrami_pressure(local_coords(1), 0.0).reshape(nc, 8), np.zeros((nc, 1))], axis=1).reshape(-1)
    rows = np.arange(nc * 81)
    u_idx = (3 * ((rows // 81) * 27 + rows % 27) + (rows % 81) // 27).astype(np.int64)
    u_lists = (np.arange(nc * 81 + 1, dtype=np.int64), u_idx, np.ones(nc * 81))
    p_lists = (np.arange(nc * 9 + 1, dtype=np.int64), np.arange(nc * 9, dtype=np.int64), np.ones(nc * 9))
    om = orc.OracleMesh(mesh, 4, None, u_lists, p_lists)
    xq = mesh.quadrature_points(4)
    e_u = orc.lib().oracle_l2_error(C.byref(om.c), 3, orc._p(np.ascontiguousarray(uvals.reshape(-1))),
                                    orc._p(np.ascontiguousarray(beltrami_velocity(xq, 0.0))))
    e_p = orc.lib().oracle_l2_error(C.byref(om.c), 1, orc._p(pvals), orc._p(np.ascontiguousarray(beltrami_pressure(xq, 0.0)[..., None])))
    assert f"{e_p:.4g}" == GOLDEN_AUGP["l2_errors_t0"]["e_p"] == "0.09298"
    assert f"{e_u:.4g}" == GOLDEN_AUGP["l2_errors_t0"]["e_u"] == "0.001572"


def augp_first_residual_inputs(mesh, dt=0.05):
    u0 = beltrami_velocity(mesh.u.coords, 0.0).reshape(-1)
    p0 = beltrami_pressure(mesh.p.coords, 0.0).reshape(-1)
    p0[mesh.p.n_q1_owned:] = 0.0                                  # DG0 part of the interpolated pressure
    u = u0.copy()
    d = mesh.u.dirichlet.reshape(-1)
    u[d] = beltrami_velocity(mesh.u.coords, dt).reshape(-1)[d]
    return u0, p0, u


def test_augp_first_nonlinear_residual_matches_golden_log():
    mesh = augp_mesh()
    _, oprm = flow_params(augmented=True)
    dt = 0.05
    ns = orc.OracleNavierStokesMatrix(mesh, oprm, orc.TimeWeights(1 / dt, -1 / dt, 0.0, 1.0, 1.0, 0.0), None, True, mesh.u_lists, mesh.p_lists)
    u0, p0, u = augp_first_residual_inputs(mesh, dt)
    ns.old_u, ns.old_old_u = u0.copy(), np.zeros_like(u0)
    ru, rp = np.zeros(ns.n_u), np.zeros(ns.n_p)
    ns.residual(ru, rp, u, p0, np.zeros(ns.n_u), np.zeros(ns.n_p))
    ns.apply_pressure_average_projection(rp)                      # both null-space modes (Q1 constant, DG0 constant)
    assert f"{np.linalg.norm(ru):.3e}" == GOLDEN_AUGP["first_nonlinear_residual"]["1"]["u"] == "2.953e+00"
    assert f"{np.linalg.norm(rp):.3e}" == GOLDEN_AUGP["first_nonlinear_residual"]["1"]["p"] == "1.886e-01"


def test_golden_fixtures_match_the_reference_logs():
    """every number of tests/golden/*.json stands on the cited line of the reference's log (tests/golden/check_golden.py);
    runs only where the reference tree exists (the development container) -- the fixtures are what travels"""
    import importlib.util
    import os
    root = os.environ.get("ADAFLO_REFERENCE_ROOT", "/root/reference")
    if not os.path.isdir(os.path.join(root, "tests")):
        pytest.skip("reference tree not present")
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("check_golden", os.path.join(here, "golden", "check_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    report, bad = mod.check_all(root)
    assert len(report) == 7 and not bad, "\n".join(bad)
