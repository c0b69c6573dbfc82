"""ns_oracle_2d.py -- dense 2-D restatement of adaflo's Navier-Stokes operators (Q2/Q1, Cartesian cells).

TEST INFRASTRUCTURE ONLY (same rules as ns_oracle.c: only tests/ may import it).

It restates, for dim = 2, the same per-quadrature-point formulas of source/navier_stokes_matrix.cc that
oracle/ns_oracle.c restates for dim = 3 (vmult :702-893, residual :601-700, constrained rows :247-256, mean
projection :191-205), as dense numpy contractions over all cells -- no sum factorisation, no extrusion.  The library
runs 2-D problems through its 3-D kernels on a one-cell-thick extrusion (adaflo_b200.operators.NavierStokesMatrix2D);
this file is the independent check of that construction.

Parity status: pinned against the reference's golden log tests/couette.output:10 -- the first nonlinear residual
1.601e+01 of the 2-D lid-driven channel (Q2/Q1, 64 x 16 cells, BDF-2 start-up weights, Dirichlet lid, constrained
tangential velocity on the open boundaries) is reproduced to all printed digits -- and cross-validated against the pinned
3-D oracle through the extrusion identity A_2d = R A_3d E (both in tests/test_2d_cpu.py).  tests/beltrami_2d is Q4/Q3
on an adaptive mesh and stays out of reach.
"""
import numpy as np

NEWTON, PICARD = 0, 1
INCOMPRESSIBLE, STATIONARY, STOKES = 0, 1, 2


def _gauss3():
    s = np.sqrt(0.6)
    return np.array([0.5 - 0.5 * s, 0.5, 0.5 + 0.5 * s]), np.array([5.0, 8.0, 5.0]) / 18.0


def _q2(x):
    v = np.stack([2 * (x - 0.5) * (x - 1), 4 * x * (1 - x), 2 * x * (x - 0.5)], axis=-1)
    d = np.stack([4 * x - 3, 4 - 8 * x, 4 * x - 1], axis=-1)
    return v, d


def _q1(x):
    return np.stack([1 - x, x], axis=-1), np.stack([-np.ones_like(x), np.ones_like(x)], axis=-1)


class OracleNavierStokes2D:
    def __init__(self, mesh2d, physical_type, linearization, density, viscosity, damping, tau_grad_div, beta, tw,
                 pressure_average_fix=True):
        """tw: (weight, weight_old, weight_old_old, tau1); damping with the sign of FlowParameters::damping"""
        self.m = mesh2d
        self.physical_type, self.linearization = physical_type, linearization
        self.rho, self.mu, self.damping, self.tau_gd, self.beta = density, viscosity, damping, tau_grad_div, beta
        self.weight, self.w_old, self.w_oo, self.tau1 = tw
        x, w = _gauss3()
        vu, du = _q2(x)                       # [q1d, node1d]
        vp, _ = _q1(x)
        # points q = qx + 3 qy, nodes n = ix + 3 iy (x fastest)
        self.phi_u = np.einsum("bj,ai->baji", vu, vu).reshape(9, 9)            # [q, n]
        gx = np.einsum("bj,ai->baji", vu, du).reshape(9, 9)
        gy = np.einsum("bj,ai->baji", du, vu).reshape(9, 9)
        self.dphi_u = np.stack([gx, gy], axis=-1)                              # [q, n, f]
        self.phi_p = np.einsum("bj,ai->baji", vp, vp).reshape(9, 4)
        self.wq = np.einsum("b,a->ba", w, w).reshape(9)
        self.K = 1.0 / np.asarray(mesh2d.h)                                    # diagonal inverse Jacobian
        self.jxw = self.wq * np.prod(mesh2d.h)
        self.n_u, self.n_p = mesh2d.n_u_nodes * 2, mesh2d.n_p_nodes
        self.lin = np.zeros((mesh2d.n_cells, 6, 9))                            # u (2), grad u (4) at every point
        self.old_u = np.zeros(self.n_u)
        self.old_old_u = np.zeros(self.n_u)
        self.mode_w = None
        if pressure_average_fix:
            # weights = pressure mass matrix applied to 1 (QGauss(2) is exact for bilinears: h^2 / 4 per vertex and cell)
            wv = np.zeros(self.n_p)
            np.add.at(wv, mesh2d.cell_p.reshape(-1), np.prod(mesh2d.h) / 4.0)
            self.mode_w = wv

    # -------------------------------------------------------------------------------------------------
    def _gather(self, vec, cells, n_comp, dirichlet=None):
        v = vec.reshape(-1, n_comp)[cells]                                     # [cell, node, comp]
        if dirichlet is not None:
            v = np.where(dirichlet[cells], 0.0, v)
        return v

    def _loop(self, op, src_u, src_p, dst_u, dst_p):
        """op: 'vmult', 'vmult_velocity' or 'residual'; adds the cell integrals into dst"""
        m = self.m
        resolve = op != "residual"
        ul = self._gather(src_u, m.cell_u, 2, m.u_dirichlet if resolve else None)
        val = np.einsum("qn,cnd->cqd", self.phi_u, ul)
        grad = np.einsum("qnf,cnd->cqdf", self.dphi_u, ul) * self.K[None, None, None, :]   # [c, q, d, e]
        pres = 0.0
        if op != "vmult_velocity":
            pl = self._gather(src_p, m.cell_p, 1)[:, :, 0]
            pres = np.einsum("qn,cn->cq", self.phi_p, pl)
        div = grad[:, :, 0, 0] + grad[:, :, 1, 1]
        stokes = self.physical_type == STOKES
        gamma = self.weight if self.physical_type == INCOMPRESSIBLE else 0.0
        conv = np.zeros_like(val)
        if not stokes:
            conv = gamma * val
            if op == "residual":
                if self.physical_type != STATIONARY:
                    uo = np.einsum("qn,cnd->cqd", self.phi_u, self._gather(self.old_u, m.cell_u, 2))
                    uoo = np.einsum("qn,cnd->cqd", self.phi_u, self._gather(self.old_old_u, m.cell_u, 2))
                    conv = conv + self.w_old * uo + self.w_oo * uoo
                res = self.beta * div[:, :, None] * val + np.einsum("cqe,cqde->cqd", val, grad)
                conv = conv + self.tau1 * res
                self.lin[:, 0:2, :] = val.transpose(0, 2, 1)
                if self.linearization == NEWTON:
                    self.lin[:, 2:6, :] = grad.reshape(grad.shape[0], 9, 4).transpose(0, 2, 1)
                else:
                    self.lin[:, 2, :] = div
            elif self.linearization == NEWTON:
                lu = self.lin[:, 0:2, :].transpose(0, 2, 1)
                lg = self.lin[:, 2:6, :].transpose(0, 2, 1).reshape(-1, 9, 2, 2)
                ltr = lg[:, :, 0, 0] + lg[:, :, 1, 1]
                res = self.beta * div[:, :, None] * lu + self.beta * ltr[:, :, None] * val \
                    + np.einsum("cqe,cqde->cqd", lu, grad) + np.einsum("cqe,cqde->cqd", val, lg)
                conv = conv + self.tau1 * res
            else:
                lu = self.lin[:, 0:2, :].transpose(0, 2, 1)
                ldiv = self.lin[:, 2, :]
                conv = conv + self.tau1 * (self.beta * ldiv[:, :, None] * val + np.einsum("cqe,cqde->cqd", lu, grad))
            conv = conv * self.rho - self.damping * val
        tmu = self.mu * self.tau1
        S = tmu * (grad + grad.transpose(0, 1, 3, 2))
        for d in range(2):
            S[:, :, d, d] = 2 * tmu * grad[:, :, d, d] + self.tau_gd * div - (pres if op != "vmult_velocity" else 0.0)
        Sref = S * self.K[None, None, None, :] * self.jxw[None, :, None, None]               # [c, q, d, f]
        out_u = np.einsum("qnf,cqdf->cnd", self.dphi_u, Sref)
        if not stokes:
            out_u = out_u + np.einsum("qn,cqd->cnd", self.phi_u, conv * self.jxw[None, :, None])
        keep = ~m.u_dirichlet[m.cell_u]                                                     # constrained rows are not written
        idx = (2 * m.cell_u[:, :, None] + np.arange(2)[None, None, :])
        np.add.at(dst_u, idx[keep], out_u[keep])
        if op != "vmult_velocity":
            out_p = np.einsum("qn,cq->cn", self.phi_p, -div * self.jxw[None, :])
            np.add.at(dst_p, m.cell_p.reshape(-1), out_p.reshape(-1))

    def _project(self, p):
        if self.mode_w is not None and self.physical_type != STATIONARY:
            p -= float(self.mode_w @ p) / float(self.mode_w.sum())

    # ------------------------------------------------------------------------------------------------- operators
    def vmult(self, src_u, src_p):
        dst_u, dst_p = np.zeros(self.n_u), np.zeros(self.n_p)
        self._loop("vmult", src_u, src_p, dst_u, dst_p)
        c = self.m.u_dirichlet.reshape(-1)
        dst_u[c] = src_u[c]
        self._project(dst_p)
        return dst_u, dst_p

    def velocity_vmult(self, src_u):
        dst_u = np.zeros(self.n_u)
        self._loop("vmult_velocity", src_u, None, dst_u, None)
        c = self.m.u_dirichlet.reshape(-1)
        dst_u[c] = src_u[c]
        return dst_u

    def residual(self, rhs_u, rhs_p, src_u, src_p, user_u, user_p):
        """rhs = user - (cell integrals + rhs preload), navier_stokes_matrix.cc:266-293"""
        self._loop("residual", src_u, src_p, rhs_u, rhs_p)
        rhs_u *= -1.0
        rhs_u += user_u
        rhs_p *= -1.0
        rhs_p += user_p


class OracleNavierStokes2DVariable(OracleNavierStokes2D):
    """The same operator with density and viscosity given per quadrature point (rho_q, mu_q: [cell, 9], the arrays
    LevelSetOKZSolver::local_compute_force fills, level_set_okz.cc:333-334; variable_densities / variable_viscosities of
    navier_stokes_matrix.h:293-395), time-dependent Newton linearisation.  Parity status: pinned through the second time step
    of tests/rising_bubble_ls.output (tests/test_2d_cpu.py: the Newton solve of step 1 with this operator leads to the
    printed advection residual, iteration counts and first Navier-Stokes residual of step 2)."""

    def __init__(self, mesh2d, rho_q, mu_q, damping, tau_grad_div, beta, tw):
        super().__init__(mesh2d, INCOMPRESSIBLE, NEWTON, 1.0, 1.0, damping, tau_grad_div, beta, tw, pressure_average_fix=False)
        self.rho_q, self.mu_q = np.asarray(rho_q), np.asarray(mu_q)

    def _loop(self, op, src_u, src_p, dst_u, dst_p):
        assert op in ("vmult", "residual")
        m = self.m
        resolve = op != "residual"
        ul = self._gather(src_u, m.cell_u, 2, m.u_dirichlet if resolve else None)
        val = np.einsum("qn,cnd->cqd", self.phi_u, ul)
        grad = np.einsum("qnf,cnd->cqdf", self.dphi_u, ul) * self.K[None, None, None, :]
        pres = np.einsum("qn,cn->cq", self.phi_p, self._gather(src_p, m.cell_p, 1)[:, :, 0])
        div = grad[:, :, 0, 0] + grad[:, :, 1, 1]
        conv = self.weight * val
        if op == "residual":
            uo = np.einsum("qn,cnd->cqd", self.phi_u, self._gather(self.old_u, m.cell_u, 2))
            uoo = np.einsum("qn,cnd->cqd", self.phi_u, self._gather(self.old_old_u, m.cell_u, 2))
            conv = conv + self.w_old * uo + self.w_oo * uoo
            conv = conv + self.tau1 * (self.beta * div[:, :, None] * val + np.einsum("cqe,cqde->cqd", val, grad))
            self.lin[:, 0:2, :] = val.transpose(0, 2, 1)
            self.lin[:, 2:6, :] = grad.reshape(grad.shape[0], 9, 4).transpose(0, 2, 1)
        else:
            lu = self.lin[:, 0:2, :].transpose(0, 2, 1)
            lg = self.lin[:, 2:6, :].transpose(0, 2, 1).reshape(-1, 9, 2, 2)
            ltr = lg[:, :, 0, 0] + lg[:, :, 1, 1]
            conv = conv + self.tau1 * (self.beta * div[:, :, None] * lu + self.beta * ltr[:, :, None] * val
                                       + np.einsum("cqe,cqde->cqd", lu, grad) + np.einsum("cqe,cqde->cqd", val, lg))
        conv = conv * self.rho_q[:, :, None] - self.damping * val
        tmu = self.mu_q * self.tau1                                              # [cell, q]
        S = tmu[:, :, None, None] * (grad + grad.transpose(0, 1, 3, 2))
        for d in range(2):
            S[:, :, d, d] = 2 * tmu * grad[:, :, d, d] + self.tau_gd * div - pres
        Sref = S * self.K[None, None, None, :] * self.jxw[None, :, None, None]
        out_u = np.einsum("qnf,cqdf->cnd", self.dphi_u, Sref) + np.einsum("qn,cqd->cnd", self.phi_u, conv * self.jxw[None, :, None])
        keep = ~m.u_dirichlet[m.cell_u]
        idx = (2 * m.cell_u[:, :, None] + np.arange(2)[None, None, :])
        np.add.at(dst_u, idx[keep], out_u[keep])
        np.add.at(dst_p, m.cell_p.reshape(-1), np.einsum("qn,cq->cn", self.phi_p, -div * self.jxw[None, :]).reshape(-1))
