"""ctypes front end of the CPU oracle (oracle/ns_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of ns_oracle.c.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The operator-level logic composed here (zeroing,
constrained rows, mean projection, coefficient rules of the pressure
sub-operators) restates source/navier_stokes_matrix.cc:85-483.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

OP_VMULT, OP_RESIDUAL, OP_VMULT_VELOCITY = 0, 1, 2
INCOMPRESSIBLE, STATIONARY, STOKES = 0, 1, 2
NEWTON, PICARD, SEMI_IMPLICIT, EXPLICIT, PROJECTION = 0, 1, 2, 3, 4


class _Mesh(C.Structure):
    _fields_ = [("n_cells", C.c_int32), ("n_q_1d", C.c_int32), ("u_degree", C.c_int32), ("p_degree", C.c_int32),
                ("p_dg0", C.c_int32), ("pad_", C.c_int32),
                ("u_ptr", C.c_void_p), ("u_idx", C.c_void_p), ("u_w", C.c_void_p),
                ("p_ptr", C.c_void_p), ("p_idx", C.c_void_p), ("p_w", C.c_void_p),
                ("u_constrained", C.c_void_p), ("p_constrained", C.c_void_p),
                ("inv_jac", C.c_void_p), ("jxw", C.c_void_p)]


class FlowParams(C.Structure):
    _fields_ = [("physical_type", C.c_int32), ("linearization", C.c_int32), ("augmented_taylor_hood", C.c_int32),
                ("variable_coefficients", C.c_int32), ("density", C.c_double), ("viscosity", C.c_double),
                ("damping", C.c_double), ("tau_grad_div", C.c_double), ("beta", C.c_double)]


class TimeWeights(C.Structure):
    _fields_ = [("weight", C.c_double), ("weight_old", C.c_double), ("weight_old_old", C.c_double),
                ("tau1", C.c_double), ("extrapolate_old", C.c_double), ("extrapolate_old_old", C.c_double)]


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "ns_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-std=c11", "-shared", "-o", so, src, "-lm"])
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.oracle_l2_error.restype = C.c_double
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def gauss01(n):
    x = np.zeros(n)
    w = np.zeros(n)
    lib().oracle_gauss01(n, _p(x), _p(w))
    return x, w


def shape_1d(degree, nq):
    v = np.zeros((nq, degree + 1))
    d = np.zeros((nq, degree + 1))
    lib().oracle_shape_1d(degree, nq, _p(v), _p(d))
    return v, d


class OracleMesh:
    """Uncompressed mesh description for the oracle, built from a BoxMesh-like
    object (spaces with cell_nodes / dirichlet, general geometry arrays)."""

    def __init__(self, mesh, n_q_1d=3, deform=None, u_lists=None, p_lists=None):
        self.mesh = mesh
        self.n_q_1d = n_q_1d
        nc = mesh.n_cells
        u, p = mesh.u, mesh.p
        self.p_dg0 = 1 if getattr(p, "n_dg0", 0) else 0
        # velocity gather lists: row = cell*81 + c*27 + n  ->  entry node*3 + c
        if u_lists is None:
            cn = u.cell_nodes.astype(np.int64)
            idx = (cn[:, None, :] * 3 + np.arange(3)[None, :, None]).reshape(-1)
            self.u_idx = np.ascontiguousarray(idx)
            self.u_ptr = np.arange(len(idx) + 1, dtype=np.int64)
            self.u_w = np.ones(len(idx))
        else:
            self.u_ptr, self.u_idx, self.u_w = u_lists
        if p_lists is None:
            idx = p.cell_nodes.astype(np.int64).reshape(-1)
            self.p_idx = np.ascontiguousarray(idx)
            self.p_ptr = np.arange(len(idx) + 1, dtype=np.int64)
            self.p_w = np.ones(len(idx))
        else:
            self.p_ptr, self.p_idx, self.p_w = p_lists
        self.u_con = np.ascontiguousarray(u.dirichlet.reshape(-1).astype(np.uint8))
        self.p_con = np.ascontiguousarray(p.dirichlet.reshape(-1).astype(np.uint8))
        K, jxw = mesh.general_geometry(n_q_1d, deform)
        self.inv_jac = np.ascontiguousarray(K)
        self.jxw = np.ascontiguousarray(jxw)
        self.c = _Mesh(nc, n_q_1d, 2, 1, self.p_dg0, 0, _p(self.u_ptr), _p(self.u_idx), _p(self.u_w),
                       _p(self.p_ptr), _p(self.p_idx), _p(self.p_w), _p(self.u_con), _p(self.p_con),
                       _p(self.inv_jac), _p(self.jxw))
        self.n_q = n_q_1d ** 3


class OracleNavierStokesMatrix:
    """CPU restatement of adaflo::NavierStokesMatrix<3> for one rank (no ghosts
    exchange: use with nranks == 1, or exchange by hand in the gloo tests)."""

    def __init__(self, mesh, prm: FlowParams, tw: TimeWeights, deform=None, pressure_average_fix=True,
                 u_lists=None, p_lists=None):
        self.mesh = mesh
        self.prm = prm
        self.tw = tw
        self._u_lists, self._p_lists = u_lists, p_lists
        self.mu = OracleMesh(mesh, 3, deform, u_lists, p_lists)   # quad_index_u: QGauss(3)
        self.mp = OracleMesh(mesh, 2, deform, u_lists, p_lists)   # quad_index_p: QGauss(2)
        nc, nq = mesh.n_cells, 27
        self.n_u = mesh.u.n_dofs_local
        self.n_p = mesh.p.n_dofs_local
        self.lin = np.zeros((nc, 12, nq))
        self.rho = self.muq = self.damp = None
        if prm.variable_coefficients:
            # initialize(): navier_stokes_matrix.cc:100-106
            self.rho = np.full((nc, nq), prm.density)
            self.muq = np.full((nc, nq), prm.viscosity)
            self.damp = np.zeros((nc, nq))
        self.old_u = np.zeros(self.n_u)
        self.old_old_u = np.zeros(self.n_u)
        self.lin_prec = None
        # pressure null-space data, navier_stokes_matrix.cc:110-168
        self.modes = [None, None]
        self.mode_weights = [None, None]
        self.inv_weights = [0.0, 0.0]
        if pressure_average_fix or prm.augmented_taylor_hood:
            pres_mass = np.zeros(self.n_p)
            lib().oracle_pressure_op(C.byref(self.mp.c), 2, None, 0, None, _p(pres_mass))
            n_own = mesh.p.n_owned
            n_dg0 = getattr(mesh.p, "n_dg0", 0)
            is_dg0 = np.zeros(self.n_p, bool)
            if n_dg0:
                is_dg0[mesh.p.n_q1_owned:mesh.p.n_q1_owned + n_dg0] = True
            const_modes = [~is_dg0, is_dg0] if n_dg0 else [np.ones(self.n_p, bool), None]
            for mode in range(2):
                if mode == 0 and not pressure_average_fix:
                    continue
                if mode == 1 and not prm.augmented_taylor_hood:
                    continue
                m = np.zeros(self.n_p)
                w = np.zeros(self.n_p)
                sel = const_modes[mode].copy()
                sel[n_own:] = False
                m[sel] = 1.0
                w[sel] = pres_mass[sel]
                m[mesh.p.dirichlet.reshape(-1)] = 0.0
                self.modes[mode] = m
                self.mode_weights[mode] = w
                self.inv_weights[mode] = 1.0 / float(m[:n_own] @ w[:n_own])

    # ------------------------------------------------------------------ helpers
    def _loop(self, op, src_u, src_p, dst_u, dst_p):
        lib().oracle_ns_cell_loop(C.byref(self.mu.c), C.byref(self.prm), C.byref(self.tw), op, _p(src_u), _p(src_p),
                                  _p(self.old_u), _p(self.old_old_u), _p(dst_u), _p(dst_p), _p(self.lin),
                                  _p(self.rho), _p(self.muq), _p(self.damp))

    def apply_pressure_average_projection(self, p):
        """navier_stokes_matrix.cc:191-205"""
        if self.prm.linearization != PROJECTION and self.prm.physical_type != STATIONARY:
            n_own = self.mesh.p.n_owned
            for mode in range(2):
                if self.mode_weights[mode] is not None:
                    product = float(self.mode_weights[mode][:n_own] @ p[:n_own])
                    p += (-product * self.inv_weights[mode]) * self.modes[mode]

    # --------------------------------------------------------------- operators
    def vmult(self, src_u, src_p):
        """navier_stokes_matrix.cc:221-262"""
        dst_u = np.zeros(self.n_u)
        dst_p = np.zeros(self.n_p)
        self._loop(OP_VMULT, src_u, src_p, dst_u, dst_p)
        cu = self.mesh.u.constrained_dofs()
        cp = self.mesh.p.constrained_dofs()
        dst_u[cu] = src_u[cu]
        dst_p[cp] = -src_p[cp]
        self.apply_pressure_average_projection(dst_p)
        return dst_u, dst_p

    def residual(self, rhs_u, rhs_p, src_u, src_p, user_u, user_p):
        """navier_stokes_matrix.cc:266-293 (rhs is read-modify-write)"""
        self._loop(OP_RESIDUAL, src_u, src_p, rhs_u, rhs_p)
        rhs_u *= -1.0
        rhs_u += user_u
        rhs_p *= -1.0
        rhs_p += user_p

    def velocity_vmult(self, src_u):
        """navier_stokes_matrix.cc:337-382"""
        dst_u = np.zeros(self.n_u)
        lin_keep = self.lin
        if self.lin_prec is not None:
            self.lin = self.lin_prec
        self._loop(OP_VMULT_VELOCITY, src_u, None, dst_u, None)
        self.lin = lin_keep
        cu = self.mesh.u.constrained_dofs()
        dst_u[cu] = src_u[cu]
        return dst_u

    def fix_linearization_point(self):
        self.lin_prec = self.lin.copy()

    def divergence_vmult_add(self, dst_p, src_u, weight_by_viscosity=False):
        """navier_stokes_matrix.cc:300-332"""
        lib().oracle_ns_divergence(C.byref(self.mu.c), C.byref(self.prm), int(weight_by_viscosity), _p(src_u),
                                   _p(dst_p), _p(self.muq))

    def _mu_center(self):
        # begin_viscosities(cell)[n_q_points(0)/2]: the centre point of QGauss(3)^3
        return self.muq[:, 27 // 2] if self.muq is not None else np.full(self.mesh.n_cells, self.prm.viscosity)

    def pressure_mass_vmult(self, src_p):
        """navier_stokes_matrix.cc:421-455, kernel :1036-1071"""
        if self.prm.linearization == PROJECTION or self.prm.physical_type == STATIONARY:
            coeff = np.ones(self.mesh.n_cells)
        else:
            coeff = 1.0 / (self._mu_center() + self.prm.tau_grad_div)
        coeff = np.ascontiguousarray(coeff)
        dst = np.zeros(self.n_p)
        lib().oracle_pressure_op(C.byref(self.mp.c), 0, _p(coeff), 0, _p(src_p), _p(dst))
        cp = self.mesh.p.constrained_dofs()
        dst[cp] = src_p[cp]
        if self.mode_weights[1] is not None and self.prm.linearization != PROJECTION:
            n_own = self.mesh.p.n_owned
            w = float(self.mode_weights[1][:n_own] @ dst[:n_own])
            dst += (-w * self.inv_weights[1]) * self.modes[1]
        return dst

    def pressure_poisson_vmult(self, src_p):
        """navier_stokes_matrix.cc:386-417, kernel :965-1032"""
        var = self.rho is not None and self.prm.linearization != PROJECTION
        dst = np.zeros(self.n_p)
        if var and self.prm.physical_type != STATIONARY:
            coeff = np.ascontiguousarray(1.0 / (self.tw.weight * self.rho))
            lib().oracle_pressure_op(C.byref(self.mu.c), 1, _p(coeff), 1, _p(src_p), _p(dst))
        else:
            if var:
                rho = self.rho[:, 27 // 2]
            else:
                rho = np.full(self.mesh.n_cells, min(self.prm.density, self.prm.density + getattr(self, "density_diff", 0.0)))
            coeff = np.ones(self.mesh.n_cells) if self.prm.physical_type == STATIONARY else 1.0 / (self.tw.weight * rho)
            coeff = np.ascontiguousarray(coeff)
            lib().oracle_pressure_op(C.byref(self.mp.c), 1, _p(coeff), 0, _p(src_p), _p(dst))
        cp = self.mesh.p.constrained_dofs()
        dst[cp] = src_p[cp]
        return dst

    def pressure_convdiff_vmult(self, src_p):
        """navier_stokes_matrix.cc:459-483, kernel :1099-1140 (diffusive part only)"""
        coeff = np.ascontiguousarray(self._mu_center())
        dst = np.zeros(self.n_p)
        lib().oracle_pressure_op(C.byref(self.mu.c), 1, _p(coeff), 0, _p(src_p), _p(dst))
        cp = self.mesh.p.constrained_dofs()
        dst[cp] = src_p[cp]
        return dst

    def l2_error(self, n_comp, vec, exact_q, n_q_1d=3, deform=None):
        om = self.mu if n_q_1d == 3 else OracleMesh(self.mesh, n_q_1d, deform, self._u_lists, self._p_lists)
        exact_q = np.ascontiguousarray(exact_q)
        return lib().oracle_l2_error(C.byref(om.c), n_comp, _p(vec), _p(exact_q))


# --------------------------------------------------------------------------
# level-set operators (oracle/ls_oracle.c)
# --------------------------------------------------------------------------
class _LSMesh(C.Structure):
    _fields_ = [("n_cells", C.c_int32), ("n_sub", C.c_int32), ("cell_nodes", C.c_void_p), ("constrained", C.c_void_p),
                ("inv_jac", C.c_void_p), ("jxw", C.c_void_p), ("cell_diameters", C.c_void_p),
                ("ptr", C.c_void_p), ("idx", C.c_void_p), ("w", C.c_void_p)]


class LSParams(C.Structure):
    _fields_ = [("epsilon", C.c_double), ("epsilon_used", C.c_double), ("damping_scale_factor", C.c_double),
                ("minimal_edge_length", C.c_double), ("step_size", C.c_double), ("dim", C.c_int32), ("pad_", C.c_int32)]


_ls_lib = None


def ls_lib():
    global _ls_lib
    if _ls_lib is None:
        so = os.path.join(_HERE, "liboracle_ls.so")
        src = os.path.join(_HERE, "ls_oracle.c")
        if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
            subprocess.check_call(["gcc", "-O2", "-fPIC", "-std=c11", "-shared", "-o", so, src, "-lm"])
        _ls_lib = C.CDLL(so)
    return _ls_lib


def iterated_gauss_geometry(mesh, n_sub, deform=None, general=False):
    """inverse Jacobians and JxW on QIterated(QGauss(2), n_sub) (q = qx + nq qy + nq^2 qz); affine cells unless
    general (then per point, mesh.iterated_geometry)"""
    if general:
        K, jxw = mesh.iterated_geometry(n_sub, deform)
        return np.ascontiguousarray(K), np.ascontiguousarray(jxw)
    K1, _ = mesh.general_geometry(1, deform)                   # cell-centre Jacobian (exact for affine cells)
    nq = (2 * n_sub) ** 3
    det = 1.0 / np.linalg.det(K1[:, 0])
    K = np.repeat(K1[:, :1], nq, axis=1)
    jxw = np.repeat((det / (8.0 * n_sub ** 3))[:, None], nq, axis=1)
    return np.ascontiguousarray(K), np.ascontiguousarray(jxw)


class OracleLevelSet:
    """Dense CPU restatement of the matrix-free level-set operators on one rank."""

    def __init__(self, mesh, prm: LSParams, deform=None, general=False):
        self.mesh = mesh
        self.prm = prm
        ls = mesh.ls
        self.n_sub = ls.n_sub
        self.n = ls.n_owned
        self.NQ = (2 * ls.n_sub) ** 3
        self.cell_nodes = np.ascontiguousarray(ls.cell_nodes.astype(np.int64))
        self.con = np.ascontiguousarray(ls.dirichlet.reshape(-1).astype(np.uint8))
        self.K, self.jxw = iterated_gauss_geometry(mesh, ls.n_sub, deform, general)
        self.cd = np.ascontiguousarray(mesh.cell_diameters())
        self.lists = getattr(mesh, "ls_lists", None)   # plain constraint expansion on meshes with hanging nodes
        lp = [_p(a) for a in self.lists] if self.lists is not None else [None, None, None]
        self.c = _LSMesh(mesh.n_cells, ls.n_sub, _p(self.cell_nodes), _p(self.con), _p(self.K), _p(self.jxw), _p(self.cd), *lp)
        self.conv = np.zeros((mesh.n_cells, self.NQ, 3))
        self.normal_q = np.zeros((mesh.n_cells, self.NQ, 3))
        self.diag = None
        self.cell_u = np.ascontiguousarray(mesh.u.cell_nodes.astype(np.int64))

    def _constrained_rows(self, dst, src):
        con = self.mesh.ls.constrained_dofs()
        if len(con):
            dst[con] = self.diag[con] * src[con]

    def mass_diagonal(self):
        d = np.zeros(self.n)
        ls_lib().ls_oracle_mass_diagonal(C.byref(self.c), _p(d))
        return d

    def advance_concentration_vmult(self, src, gamma):
        dst = np.zeros(self.n)
        ls_lib().ls_oracle_advect(C.byref(self.c), C.c_double(gamma), _p(self.conv), _p(src), _p(dst))
        self._constrained_rows(dst, src)
        return dst

    def advance_concentration_rhs(self, rhs, vel, c, c_old, c_oo, gamma, w_old, w_oo):
        ls_lib().ls_oracle_advect_rhs(C.byref(self.c), _p(self.cell_u), _p(vel), _p(c), _p(c_old), _p(c_oo), C.c_double(gamma),
                                      C.c_double(w_old), C.c_double(w_oo), _p(rhs), _p(self.conv))

    def _scalar(self, kind, scale, src):
        dst = np.zeros(self.n)
        ls_lib().ls_oracle_scalar_op(C.byref(self.c), C.byref(self.prm), kind, C.c_double(scale), _p(self.normal_q), _p(src), _p(dst))
        self._constrained_rows(dst, src)
        return dst

    def compute_normal_vmult(self, src):
        return [self._scalar(0, self.prm.damping_scale_factor, s) for s in src]

    def compute_normal_rhs(self, rhs, level_set):
        ls_lib().ls_oracle_normal_rhs(C.byref(self.c), _p(level_set), _p(rhs[0]), _p(rhs[1]), _p(rhs[2]))

    def compute_curvature_vmult(self, src, apply_diffusion=True):
        return self._scalar(0, 1.0 if apply_diffusion else 0.0, src)

    def compute_curvature_rhs(self, rhs, normal):
        ls_lib().ls_oracle_curvature_rhs(C.byref(self.c), _p(normal[0]), _p(normal[1]), _p(normal[2]), _p(rhs))

    def curvature_correction(self, curvature, level_set):
        ls_lib().ls_oracle_curvature_correction(C.c_int64(self.n), C.c_double(self.prm.epsilon_used), self.prm.dim, _p(level_set), _p(curvature))

    def reinitialization_vmult(self, src, diffuse_only=False):
        return self._scalar(2 if diffuse_only else 1, 0.0, src)

    def reinitialization_rhs(self, rhs, solution, normal, diffuse_only=False, first_reinit_step=True):
        z = np.zeros(self.n)
        nb = normal if normal is not None else [z, z, z]
        ls_lib().ls_oracle_reinit_rhs(C.byref(self.c), C.byref(self.prm), int(diffuse_only), int(first_reinit_step), _p(solution),
                                      _p(nb[0]), _p(nb[1]), _p(nb[2]), _p(self.normal_q), _p(rhs))


def discrete_heaviside(x):
    """include/adaflo/level_set_base.h:122-145"""
    if x > 0:
        return 1.0 - discrete_heaviside(-x)
    if x < -2.0:
        return 0.0
    if x < -1.0:
        return (1. / 8. * (5. * x + x * x) + 1. / 32. * (-3. - 2. * x) * np.sqrt(-7. - 12. * x - 4. * x * x)
                - 1. / 16 * np.arcsin(np.sqrt(2.) * (x + 3. / 2.)) + 23. / 32. - np.pi / 64.)
    return (1. / 8. * (3. * x + x * x) - 1. / 32. * (-1. - 2. * x) * np.sqrt(1. - 4. * x - 4. * x * x)
            + 1. / 16 * np.arcsin(np.sqrt(2.) * (x + 1. / 2.)) + 15. / 32. - np.pi / 64.)


def compute_heaviside(cell_nodes, level_set, epsilon, n_sub):
    """LevelSetOKZSolver::compute_heaviside, source/level_set_okz.cc:479-540 (cell loop, last writer wins)"""
    cutoff = np.tanh(2.0)
    H = np.zeros_like(level_set)
    for nodes in cell_nodes:
        loc = level_set[nodes]
        if np.any(np.abs(loc) < cutoff):
            for i, c in zip(nodes, loc):
                d = -3.0 if c < -cutoff else (3.0 if c > cutoff else np.log((1 + c) / (1 - c)))
                H[i] = discrete_heaviside(d * epsilon * 2.0 / n_sub)
        else:
            H[nodes] = 0.0 if loc[0] < 0 else 1.0
    return H


def compute_force(mesh, heaviside, curvature, sigma, gravity, rho0, drho, mu0, dmu, variable, interp, dst_u, deform=None):
    """dense LevelSetOKZSolver::local_compute_force; returns (rho_q, mu_q) [n_cells, 27]"""
    n_sub = mesh.ls.n_sub
    K, jxw = mesh.general_geometry(3, deform)
    K, jxw = np.ascontiguousarray(K), np.ascontiguousarray(jxw)
    cl = np.ascontiguousarray(mesh.ls.cell_nodes.astype(np.int64))
    cu = np.ascontiguousarray(mesh.u.cell_nodes.astype(np.int64))
    ucon = np.ascontiguousarray(mesh.u.dirichlet.reshape(-1).astype(np.uint8))
    rho_q, mu_q = np.zeros((mesh.n_cells, 27)), np.zeros((mesh.n_cells, 27))
    ip = None if interp is None else np.ascontiguousarray(interp)
    ls_lib().ls_oracle_force(mesh.n_cells, n_sub, _p(cl), _p(cu), _p(ucon), _p(K), _p(jxw), _p(heaviside), _p(curvature),
                             C.c_double(sigma), C.c_double(gravity), C.c_double(rho0), C.c_double(drho), C.c_double(mu0),
                             C.c_double(dmu), int(variable), _p(ip), _p(dst_u), _p(rho_q), _p(mu_q))
    return rho_q, mu_q
