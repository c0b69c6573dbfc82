"""ls_oracle_2d.py -- dense 2-D restatement of adaflo's matrix-free level-set operators (FE_Q_iso_Q1(n_sub) with
QIterated(QGauss(2), n_sub), Cartesian cells).

TEST INFRASTRUCTURE ONLY.  Same weak forms as oracle/ls_oracle.c (which cites the reference lines), written for
dim = 2 as numpy contractions over all cells.  The library runs 2-D level-set problems through its 3-D kernels on a
one-cell-thick extrusion (adaflo_b200.operators.LevelSetOperators2D); this file is the independent check.
Parity status: pinned against the reference's golden log tests/rising_bubble_ls.output:6,13 -- the CG iteration counts
"reinitialize (8 + 8)" and "(7 + 7)" of LevelSetOKZSolverReinitialization::reinitialize are reproduced with these
operators (normal projection, reinitialisation operator / right-hand side, mass diagonal), and line 14 -- the first
Navier-Stokes residual 0.0198 = norm of the force vector (curvature, Heaviside, compute_force) -- as well; and against
tests/spurious_currents_ls.output:9,10 -- "reinitialize (8 + 7)" and the first residual 0.365 of a surface-tension-only
problem, which pins the curvature (normalised divergence, projection, distance correction) to three digits
(tests/test_2d_cpu.py).  The 3-D oracle is tied to this file through the extrusion identity in the same test file.
"""
import numpy as np


class OracleLevelSet2D:
    def __init__(self, mesh2d, n_sub, epsilon, epsilon_used, minimal_edge_length, step_size, cell_diameters):
        m = self.m = mesh2d
        self.n_sub, self.eps, self.eps_used = n_sub, epsilon, epsilon_used
        self.min_edge, self.step = minimal_edge_length, step_size
        self.cd = np.asarray(cell_diameters, dtype=np.float64)
        n1, nq1 = n_sub + 1, 2 * n_sub
        g = [0.5 - 0.5 / np.sqrt(3.0), 0.5 + 0.5 / np.sqrt(3.0)]
        V, G = np.zeros((nq1, n1)), np.zeros((nq1, n1))
        L2 = np.zeros((nq1, 3))
        for q in range(nq1):
            s, xi = q // 2, g[q % 2]
            V[q, s], V[q, s + 1] = 1 - xi, xi
            G[q, s], G[q, s + 1] = -float(n_sub), float(n_sub)
            x = (s + xi) / n_sub
            L2[q] = [2 * (x - 0.5) * (x - 1), 4 * x * (1 - x), 2 * x * (x - 0.5)]
        self.phi = np.einsum("bj,ai->baji", V, V).reshape(nq1 * nq1, n1 * n1)          # [q = qx + nq1 qy, n = ix + n1 iy]
        self.gphi = np.stack([np.einsum("bj,ai->baji", V, G).reshape(nq1 * nq1, n1 * n1),
                              np.einsum("bj,ai->baji", G, V).reshape(nq1 * nq1, n1 * n1)], axis=-1)
        self.phi_u = np.einsum("bj,ai->baji", L2, L2).reshape(nq1 * nq1, 9)
        self.K = 1.0 / np.asarray(m.h)
        self.jxw = np.full(nq1 * nq1, np.prod(m.h) / (4.0 * n_sub * n_sub))
        self.cells = m.cell_ls
        self.con = m.ls_dirichlet
        self.n = m.n_ls_nodes
        self.conv = np.zeros((m.n_cells, nq1 * nq1, 2))
        self.normal_q = np.zeros((m.n_cells, nq1 * nq1, 2))
        self.diag = None

    # ---- building blocks
    def _gather(self, vec, resolve):
        v = np.asarray(vec)[self.cells]
        return np.where(self.con[self.cells], 0.0, v) if resolve else v

    def _eval(self, loc):
        return np.einsum("qn,cn->cq", self.phi, loc), np.einsum("qnf,cn->cqf", self.gphi, loc) * self.K

    def _integ(self, tv, tg):
        out = np.zeros((self.m.n_cells, self.phi.shape[1]))
        if tv is not None:
            out += np.einsum("qn,cq->cn", self.phi, tv * self.jxw)
        if tg is not None:
            out += np.einsum("qnf,cqf->cn", self.gphi, tg * self.K * self.jxw[None, :, None])
        return out

    def _scatter(self, out, dst, skip_constrained=True):
        keep = ~self.con[self.cells] if skip_constrained else np.ones_like(self.cells, dtype=bool)
        np.add.at(dst, self.cells[keep], out[keep])

    def _constrained_rows(self, dst, src):
        dst[self.con] = self.diag[self.con] * src[self.con]

    # ---- operators (names of oracle.OracleLevelSet)
    def mass_diagonal(self):
        d = np.zeros(self.n)
        out = np.einsum("qn,q->n", self.phi ** 2, self.jxw)
        self._scatter(np.broadcast_to(out, (self.m.n_cells, len(out))), d, skip_constrained=False)
        return d

    def advance_concentration_vmult(self, src, gamma):
        v, g = self._eval(self._gather(src, True))
        dst = np.zeros(self.n)
        self._scatter(self._integ(gamma * v + np.einsum("cqf,cqf->cq", g, self.conv), None), dst)
        self._constrained_rows(dst, src)
        return dst

    def advance_concentration_rhs(self, rhs, vel, c, c_old, c_oo, gamma, w_old, w_oo):
        u = np.einsum("qn,cnd->cqd", self.phi_u, np.asarray(vel).reshape(-1, 2)[self.m.cell_u])
        self.conv[:] = u
        v, g = self._eval(self._gather(c, False))
        vo, _ = self._eval(self._gather(c_old, False))
        voo, _ = self._eval(self._gather(c_oo, False))
        res = -(gamma * v + np.einsum("cqf,cqf->cq", u, g) + w_old * vo + w_oo * voo)
        self._scatter(self._integ(res, None), rhs)

    def _delta(self, scale):
        d = np.maximum(self.eps_used / self.eps, self.cd / self.n_sub)
        return scale * d * d

    def _scalar(self, kind, scale, src):
        v, g = self._eval(self._gather(src, True))
        dtau_inv = max(0.95 / (1.0 / 4.0 * self.min_edge / self.n_sub), 1.0 / (5.0 * self.step))   # dim = 2
        D = np.maximum(self.eps_used, self.cd / self.n_sub)[:, None, None]
        if kind == 0:
            out = self._integ(v, self._delta(scale)[:, None, None] * g)
        elif kind == 1:
            ng = np.einsum("cqf,cqf->cq", self.normal_q, g)
            out = self._integ(dtau_inv * v, D * ng[:, :, None] * self.normal_q)
        else:
            out = self._integ(dtau_inv * v, D * g)
        dst = np.zeros(self.n)
        self._scatter(out, dst)
        self._constrained_rows(dst, src)
        return dst

    def compute_normal_vmult(self, src, damping_scale_factor):
        return [self._scalar(0, damping_scale_factor, s) for s in src]

    def compute_normal_rhs(self, rhs, level_set):
        _, g = self._eval(self._gather(level_set, False))
        for d in range(2):
            self._scatter(self._integ(g[:, :, d], None), rhs[d])

    def compute_curvature_vmult(self, src, apply_diffusion=True):
        return self._scalar(0, 1.0 if apply_diffusion else 0.0, src)

    def compute_curvature_rhs(self, rhs, normal):
        loc = np.stack([self._gather(n, False) for n in normal], axis=-1)            # [c, n, 2]
        nn = np.linalg.norm(loc, axis=-1)
        ok = nn > 1e-2
        loc = np.where(ok[:, :, None], loc / np.where(ok, nn, 1.0)[:, :, None], 0.0)
        div = sum(self._eval(loc[:, :, d])[1][:, :, d] for d in range(2))
        out = self._integ(-div, None)
        out[~ok.any(axis=1)] = 0.0
        self._scatter(out, rhs)

    def reinitialization_vmult(self, src, diffuse_only=False):
        return self._scalar(2 if diffuse_only else 1, 0.0, src)

    def reinitialization_rhs(self, rhs, solution, normal, diffuse_only=False, first_reinit_step=True):
        v, g = self._eval(self._gather(solution, False))
        D = np.maximum(self.eps_used, self.cd / self.n_sub)[:, None]
        if not diffuse_only:
            if first_reinit_step:
                nv = np.stack([self._eval(self._gather(n, False))[0] for n in normal], axis=-1)
                self.normal_q[:] = nv / np.maximum(1e-4, np.linalg.norm(nv, axis=-1))[:, :, None]
            ng = np.einsum("cqf,cqf->cq", self.normal_q, g)
            tg = self.normal_q * (0.5 * (1 - v * v) - ng * D)[:, :, None]
        else:
            tg = -D[:, :, None] * g
        self._scatter(self._integ(None, tg), rhs)


def compute_force_2d(mesh2d, n_sub, heaviside, curvature, sigma, gravity, rho0, drho, mu0, dmu, variable, dst_u, interpolate=False,
                     u_dirichlet=None):
    """dense LevelSetOKZSolver::local_compute_force (level_set_okz.cc:317-409) for dim = 2 on QGauss(3)^2, without the
    level-set -> pressure interpolation; returns (rho_q, mu_q) [n_cells, 9]; gravity acts on component dim - 1 = y"""
    m = mesh2d
    n1 = n_sub + 1
    s35 = np.sqrt(0.6)
    gp = np.array([0.5 - 0.5 * s35, 0.5, 0.5 + 0.5 * s35])
    gw = np.array([5.0, 8.0, 5.0]) / 18.0
    hv, hd, L2 = np.zeros((3, n1)), np.zeros((3, n1)), np.zeros((3, 3))
    for p in range(3):
        s = min(int(gp[p] * n_sub), n_sub - 1)
        xi = gp[p] * n_sub - s
        hv[p, s], hv[p, s + 1] = 1 - xi, xi
        hd[p, s], hd[p, s + 1] = -float(n_sub), float(n_sub)
        x = gp[p]
        L2[p] = [2 * (x - 0.5) * (x - 1), 4 * x * (1 - x), 2 * x * (x - 0.5)]
    phi = np.einsum("bj,ai->baji", hv, hv).reshape(9, n1 * n1)
    gx = np.einsum("bj,ai->baji", hv, hd).reshape(9, n1 * n1)
    gy = np.einsum("bj,ai->baji", hd, hv).reshape(9, n1 * n1)
    phiu = np.einsum("bj,ai->baji", L2, L2).reshape(9, 9)
    jxw = np.einsum("b,a->ba", gw, gw).reshape(9) * np.prod(m.h)
    Hl, kl = np.asarray(heaviside)[m.cell_ls], np.asarray(curvature)[m.cell_ls]
    H = np.einsum("qn,cn->cq", phi, Hl)
    kap = np.einsum("qn,cn->cq", phi, kl)
    g = np.stack([np.einsum("qn,cn->cq", gx, Hl) / m.h[0], np.einsum("qn,cn->cq", gy, Hl) / m.h[1]], axis=-1)
    if interpolate:   # gradient of the Q1 interpolant of H (interpolate_grad_onto_pressure, level_set_okz.cc:367-378)
        corners = [0, n_sub, n1 * n_sub, n1 * n_sub + n_sub]
        Hc = Hl[:, corners]                                                  # [c, (ix, iy)]
        x = gp
        g1x = np.stack([-(1 - x), (1 - x), -x, x], axis=-1)                   # d/dxi of the 4 bilinear shapes at eta = x[qy]
        gxq = np.zeros((9, 4)); gyq = np.zeros((9, 4))
        for qy in range(3):
            for qx in range(3):
                q = qx + 3 * qy
                gxq[q] = [-(1 - x[qy]), (1 - x[qy]), -x[qy], x[qy]]
                gyq[q] = [-(1 - x[qx]), -x[qx], (1 - x[qx]), x[qx]]
        g = np.stack([np.einsum("qn,cn->cq", gxq, Hc) / m.h[0], np.einsum("qn,cn->cq", gyq, Hc) / m.h[1]], axis=-1)
    rho = rho0 + drho * H if variable else np.full_like(H, rho0)
    mu = mu0 + dmu * H
    f = sigma * kap[:, :, None] * g
    f[:, :, 1] -= gravity * rho
    out = np.einsum("qn,cqd->cnd", phiu, f * jxw[None, :, None])
    dirichlet = m.u_dirichlet if u_dirichlet is None else u_dirichlet
    keep = ~dirichlet[m.cell_u]
    idx = 2 * m.cell_u[:, :, None] + np.arange(2)[None, None, :]
    np.add.at(dst_u, idx[keep], out[keep])
    return rho, mu
