"""ns_oracle_generic.py -- dense restatement of adaflo's Navier-Stokes operators for Taylor-Hood Q_k / Q_(k-1), dim = 2 or 3.

TEST INFRASTRUCTURE ONLY (same rules as ns_oracle.c: only tests/ may import it).

It restates the per-quadrature-point formulas of source/navier_stokes_matrix.cc (vmult :702-893, residual :601-700,
constrained rows :247-256, mean projection :191-205) for any dimension and degree as dense numpy contractions over all
cells -- no sum factorisation -- exactly like oracle/ns_oracle_2d.py does for 2-D Q2/Q1.  FE_Q(k) is the Lagrange basis on
the Gauss-Lobatto points (deal.II's choice; equidistant for k <= 2), the quadrature QGauss(k + 1).

Parity status: no golden log of the reference exercises a degree other than 2/1 on a conforming mesh (tests/beltrami_2d is
Q4/Q3 on a mesh with hanging nodes), so this oracle is pinned indirectly: tests/test_generic_cpu.py proves it equal to the
pinned oracles for Q2/Q1 (oracle/ns_oracle.c in 3-D: beltrami_3d.output; oracle/ns_oracle_2d.py in 2-D: couette.output) on
the same inputs to 1e-13; for higher degrees it shares every line of code with those cases (only the 1-D tables change),
and tests/test_generic_cpu.py checks the degree-specific parts: exact integration of polynomials and vmult = Jacobian of the
residual.
"""
import numpy as np
from numpy.polynomial import legendre as npleg

NEWTON, PICARD = 0, 1
INCOMPRESSIBLE, STATIONARY, STOKES = 0, 1, 2


def gauss(n):
    x, w = npleg.leggauss(n)
    return 0.5 * (x + 1.0), 0.5 * w


def gauss_lobatto(n):
    """n Gauss-Lobatto points on [0, 1] (support points of FE_Q(n - 1))"""
    if n == 1:
        return np.array([0.5])
    if n == 2:
        return np.array([0.0, 1.0])
    c = np.zeros(n)
    c[-1] = 1.0                                       # P_{n-1}
    interior = npleg.legroots(npleg.legder(c))
    return 0.5 * (np.concatenate([[-1.0], np.sort(interior), [1.0]]) + 1.0)


def lagrange_tables(nodes, x):
    """values and derivatives [point, node] of the Lagrange basis on `nodes` at the points x"""
    n = len(nodes)
    v, d = np.ones((len(x), n)), np.zeros((len(x), n))
    for j in range(n):
        for m in range(n):
            if m != j:
                v[:, j] *= (x - nodes[m]) / (nodes[j] - nodes[m])
        for i in range(n):
            if i == j:
                continue
            t = np.full(len(x), 1.0 / (nodes[j] - nodes[i]))
            for m in range(n):
                if m != j and m != i:
                    t *= (x - nodes[m]) / (nodes[j] - nodes[m])
            d[:, j] += t
    return v, d


def tensor(dim, tabs):
    """tabs: one [q1d, n1d] table per direction (x first) -> [q, n] with x fastest in both indices"""
    if dim == 2:
        return np.einsum("bj,ai->baji", tabs[1], tabs[0]).reshape(tabs[0].shape[0] * tabs[1].shape[0], -1)
    return np.einsum("ck,bj,ai->cbakji", tabs[2], tabs[1], tabs[0]).reshape(tabs[0].shape[0] * tabs[1].shape[0] * tabs[2].shape[0], -1)


class OracleNavierStokesGeneric:
    def __init__(self, mesh, physical_type, linearization, density, viscosity, damping, tau_grad_div, beta, tw, pressure_average_fix=True):
        """mesh: adaflo_b200.mesh.TensorMesh (dim, degree, cell_u, cell_p, u_dirichlet, h); tw: (weight, weight_old, weight_old_old, tau1)"""
        self.m, self.dim, k = mesh, mesh.dim, mesh.degree
        dim = self.dim
        self.physical_type, self.linearization = physical_type, linearization
        self.rho, self.mu, self.damping, self.tau_gd, self.beta = density, viscosity, damping, tau_grad_div, beta
        self.weight, self.w_old, self.w_oo, self.tau1 = tw
        xq, wq = gauss(k + 1)
        vu, du = lagrange_tables(gauss_lobatto(k + 1), xq)
        vp, _ = lagrange_tables(gauss_lobatto(k), xq)
        self.phi_u = tensor(dim, [vu] * dim)                                                   # [q, n]
        self.dphi_u = np.stack([tensor(dim, [du if f == d else vu for d in range(dim)]) for f in range(dim)], axis=-1)   # [q, n, f]
        self.phi_p = tensor(dim, [vp] * dim)
        self.wq = tensor(dim, [wq[:, None]] * dim)[:, 0]
        self.K = 1.0 / np.asarray(mesh.h)
        self.jxw = self.wq * np.prod(mesh.h)
        self.n_u, self.n_p = mesh.n_u_nodes * dim, mesh.n_p_nodes
        self.n_q = len(self.wq)
        self.lin = np.zeros((mesh.n_cells, dim + dim * dim, self.n_q))
        self.old_u = np.zeros(self.n_u)
        self.old_old_u = np.zeros(self.n_u)
        self.mode_w = None
        if pressure_average_fix:
            wv = np.zeros(self.n_p)                                                            # pressure mass matrix applied to 1
            np.add.at(wv, mesh.cell_p.reshape(-1), np.tile(self.phi_p.T @ self.jxw, mesh.n_cells))
            self.mode_w = wv

    def _gather(self, vec, cells, n_comp, dirichlet=None):
        v = vec.reshape(-1, n_comp)[cells]
        if dirichlet is not None:
            v = np.where(dirichlet[cells], 0.0, v)
        return v

    def _loop(self, op, src_u, src_p, dst_u, dst_p):
        m, dim = self.m, self.dim
        resolve = op != "residual"
        ul = self._gather(src_u, m.cell_u, dim, m.u_dirichlet if resolve else None)
        val = np.einsum("qn,cnd->cqd", self.phi_u, ul)
        grad = np.einsum("qnf,cnd->cqdf", self.dphi_u, ul) * self.K[None, None, None, :]          # [c, q, d, e]
        pres = 0.0
        if op != "vmult_velocity":
            pres = np.einsum("qn,cn->cq", self.phi_p, self._gather(src_p, m.cell_p, 1)[:, :, 0])
        div = np.einsum("cqdd->cq", grad)
        stokes = self.physical_type == STOKES
        gamma = self.weight if self.physical_type == INCOMPRESSIBLE else 0.0
        conv = np.zeros_like(val)
        nq = self.n_q
        if not stokes:
            conv = gamma * val
            if op == "residual":
                if self.physical_type != STATIONARY:
                    uo = np.einsum("qn,cnd->cqd", self.phi_u, self._gather(self.old_u, m.cell_u, dim))
                    uoo = np.einsum("qn,cnd->cqd", self.phi_u, self._gather(self.old_old_u, m.cell_u, dim))
                    conv = conv + self.w_old * uo + self.w_oo * uoo
                res = self.beta * div[:, :, None] * val + np.einsum("cqe,cqde->cqd", val, grad)
                conv = conv + self.tau1 * res
                self.lin[:, 0:dim, :] = val.transpose(0, 2, 1)
                if self.linearization == NEWTON:
                    self.lin[:, dim:, :] = grad.reshape(grad.shape[0], nq, dim * dim).transpose(0, 2, 1)
                else:
                    self.lin[:, dim, :] = div
            elif self.linearization == NEWTON:
                lu = self.lin[:, 0:dim, :].transpose(0, 2, 1)
                lg = self.lin[:, dim:, :].transpose(0, 2, 1).reshape(-1, nq, dim, dim)
                ltr = np.einsum("cqdd->cq", lg)
                res = self.beta * div[:, :, None] * lu + self.beta * ltr[:, :, None] * val \
                    + np.einsum("cqe,cqde->cqd", lu, grad) + np.einsum("cqe,cqde->cqd", val, lg)
                conv = conv + self.tau1 * res
            else:
                lu = self.lin[:, 0:dim, :].transpose(0, 2, 1)
                ldiv = self.lin[:, dim, :]
                conv = conv + self.tau1 * (self.beta * ldiv[:, :, None] * val + np.einsum("cqe,cqde->cqd", lu, grad))
            conv = conv * self.rho - self.damping * val
        tmu = self.mu * self.tau1
        S = tmu * (grad + grad.transpose(0, 1, 3, 2))
        for d in range(dim):
            S[:, :, d, d] = 2 * tmu * grad[:, :, d, d] + self.tau_gd * div - (pres if op != "vmult_velocity" else 0.0)
        Sref = S * self.K[None, None, None, :] * self.jxw[None, :, None, None]
        out_u = np.einsum("qnf,cqdf->cnd", self.dphi_u, Sref)
        if not stokes:
            out_u = out_u + np.einsum("qn,cqd->cnd", self.phi_u, conv * self.jxw[None, :, None])
        keep = ~m.u_dirichlet[m.cell_u]
        idx = (dim * m.cell_u[:, :, None] + np.arange(dim)[None, None, :])
        np.add.at(dst_u, idx[keep], out_u[keep])
        if op != "vmult_velocity":
            out_p = np.einsum("qn,cq->cn", self.phi_p, -div * self.jxw[None, :])
            np.add.at(dst_p, m.cell_p.reshape(-1), out_p.reshape(-1))

    def apply_pressure_average_projection(self, p):
        if self.mode_w is not None and self.physical_type != STATIONARY:
            p -= float(self.mode_w @ p) / float(self.mode_w.sum())

    def vmult(self, src_u, src_p):
        dst_u, dst_p = np.zeros(self.n_u), np.zeros(self.n_p)
        self._loop("vmult", src_u, src_p, dst_u, dst_p)
        c = self.m.u_dirichlet.reshape(-1)
        dst_u[c] = src_u[c]
        self.apply_pressure_average_projection(dst_p)
        return dst_u, dst_p

    def velocity_vmult(self, src_u):
        dst_u = np.zeros(self.n_u)
        self._loop("vmult_velocity", src_u, None, dst_u, None)
        c = self.m.u_dirichlet.reshape(-1)
        dst_u[c] = src_u[c]
        return dst_u

    def residual(self, rhs_u, rhs_p, src_u, src_p, user_u, user_p):
        """rhs = user - cell integrals, navier_stokes_matrix.cc:266-293"""
        self._loop("residual", src_u, src_p, rhs_u, rhs_p)
        rhs_u *= -1.0
        rhs_u += user_u
        rhs_p *= -1.0
        rhs_p += user_p
