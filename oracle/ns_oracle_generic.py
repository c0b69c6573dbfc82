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
        h_cell = getattr(mesh, "h_cell", None)                                                 # meshes with two cell sizes
        if h_cell is None:
            h_cell = np.tile(np.asarray(mesh.h, dtype=np.float64)[None, :], (mesh.n_cells, 1))
        self.K = 1.0 / h_cell                                                                  # [c, e]
        self.jxw = self.wq[None, :] * np.prod(h_cell, axis=1)[:, None]                         # [c, q]
        self.n_u, self.n_p = mesh.n_u_nodes * dim, mesh.n_p_nodes
        self.n_q = len(self.wq)
        # hanging nodes: constraint rows (row_ptr, parents, weights) of the last nodes of each space, applied exactly as
        # FEEvaluation::read_dof_values / distribute_local_to_global apply an AffineConstraints row: expansion matrices
        # E [cell-local slot, node] with E = identity pick for plain nodes
        self.Eu = self._expansion(mesh.cell_u, mesh.n_u_nodes, getattr(mesh, "u_rows", None))
        self.Ep = self._expansion(mesh.cell_p, mesh.n_p_nodes, getattr(mesh, "p_rows", None))
        self.p_constrained = np.zeros(self.n_p, dtype=bool)
        if getattr(mesh, "p_rows", None) is not None:
            self.p_constrained[mesh.n_p_nodes - (len(mesh.p_rows[0]) - 1):] = True
        self.u_constrained = mesh.u_dirichlet.copy()
        if getattr(mesh, "u_rows", None) is not None:
            self.u_constrained[mesh.n_u_nodes - (len(mesh.u_rows[0]) - 1):, :] = True
        self.lin = np.zeros((mesh.n_cells, dim + dim * dim, self.n_q))
        self.old_u = np.zeros(self.n_u)
        self.old_old_u = np.zeros(self.n_u)
        self.mode_w = None
        if pressure_average_fix:
            # integrals of the pressure shape functions, distributed through the constraints (navier_stokes_matrix.cc:117-133);
            # the constant mode is 1 on the unconstrained dofs (:158-163)
            wv = self._scatter(self.Ep, np.einsum("qn,cq->cn", self.phi_p, self.jxw)[:, :, None], None)[:, 0]
            self.mode_w = wv
            self.mode = (~self.p_constrained).astype(np.float64)

    @staticmethod
    def _expansion(cells, n_nodes, rows):
        import scipy.sparse as sps
        n_slots = cells.size
        flat = cells.reshape(-1)
        if rows is None:
            return sps.csr_matrix((np.ones(n_slots), (np.arange(n_slots), flat)), shape=(n_slots, n_nodes))
        ptr, cols, w = (np.asarray(a) for a in rows)
        first = n_nodes - (len(ptr) - 1)
        ri, ci, vi = [], [], []
        for slot, node in enumerate(flat):
            if node < first:
                ri.append(slot), ci.append(node), vi.append(1.0)
            else:
                for k in range(ptr[node - first], ptr[node - first + 1]):
                    ri.append(slot), ci.append(int(cols[k])), vi.append(float(w[k]))
        return sps.csr_matrix((vi, (ri, ci)), shape=(n_slots, n_nodes))

    def _gather(self, vec, cells, n_comp, E=None, dirichlet=None):
        """plain read of the vector entries (E None: read_dof_values_plain), or the constraint-resolving read (read_dof_values):
        hanging rows expanded, homogeneous Dirichlet entries read as zero"""
        v = vec.reshape(-1, n_comp)
        if E is None:
            return v[cells]
        if dirichlet is not None:
            v = np.where(dirichlet, 0.0, v)
        return (E @ v).reshape(cells.shape[0], cells.shape[1], n_comp)

    def _scatter(self, E, local, dirichlet):
        """distribute_local_to_global: E^T local, rows of homogeneous Dirichlet entries dropped; returns [n_nodes, n_comp]"""
        out = E.T @ local.reshape(-1, local.shape[-1])
        if dirichlet is not None:
            out = np.where(dirichlet, 0.0, out)
        return out

    def _loop(self, op, src_u, src_p, dst_u, dst_p):
        m, dim = self.m, self.dim
        resolve = op != "residual"
        ul = self._gather(src_u, m.cell_u, dim, self.Eu if resolve else None, m.u_dirichlet)
        val = np.einsum("qn,cnd->cqd", self.phi_u, ul)
        grad = np.einsum("qnf,cnd->cqdf", self.dphi_u, ul) * self.K[:, None, None, :]             # [c, q, d, e]
        pres = 0.0
        if op != "vmult_velocity":
            pres = np.einsum("qn,cn->cq", self.phi_p, self._gather(src_p, m.cell_p, 1, self.Ep if resolve else None)[:, :, 0])
        div = np.einsum("cqdd->cq", grad)
        stokes = self.physical_type == STOKES
        gamma = self.weight if self.physical_type == INCOMPRESSIBLE else 0.0
        conv = np.zeros_like(val)
        nq = self.n_q
        if not stokes:
            conv = gamma * val
            if op == "residual":
                if self.physical_type != STATIONARY:
                    uo = np.einsum("qn,cnd->cqd", self.phi_u, self._gather(self.old_u, m.cell_u, dim))       # plain reads
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
        Sref = S * self.K[:, None, None, :] * self.jxw[:, :, None, None]
        out_u = np.einsum("qnf,cqdf->cnd", self.dphi_u, Sref)
        if not stokes:
            out_u = out_u + np.einsum("qn,cqd->cnd", self.phi_u, conv * self.jxw[:, :, None])
        dst_u += self._scatter(self.Eu, out_u, m.u_dirichlet).reshape(-1)
        if op != "vmult_velocity":
            out_p = np.einsum("qn,cq->cn", self.phi_p, -div * self.jxw)
            dst_p += self._scatter(self.Ep, out_p[:, :, None], None)[:, 0]

    def l2_error(self, space, vec, exact, n_q_1d):
        """sqrt(int |u_h - exact|^2) with QGauss(n_q_1d) per direction (VectorTools::integrate_difference + L2 norm,
        tests/beltrami.cc:254-283: QGauss(degree + 2)); space 0: velocity, 1: pressure; exact(x[c, q, dim]) -> [c, q, n_comp];
        the vector entries are read plainly (hanging nodes hold their distributed / interpolated values)"""
        m, dim = self.m, self.dim
        k = m.degree if space == 0 else m.degree - 1
        xq, wq = gauss(n_q_1d)
        v1, _ = lagrange_tables(gauss_lobatto(k + 1), xq)
        phi = tensor(dim, [v1] * dim)
        w = tensor(dim, [wq[:, None]] * dim)[:, 0]
        grids = np.meshgrid(*([xq] * dim), indexing="ij")
        pts = np.stack([g.T.reshape(-1) if dim == 2 else g.transpose(2, 1, 0).reshape(-1) for g in grids], axis=1)   # x fastest
        h_cell = getattr(m, "h_cell", None)
        if h_cell is None:
            h_cell = np.tile(np.asarray(m.h, dtype=np.float64)[None, :], (m.n_cells, 1))
        origin = getattr(m, "origin", None)
        if origin is None:
            idx = np.stack(np.unravel_index(np.arange(m.n_cells), m.n[::-1]), axis=1)[:, ::-1]
            origin = np.asarray(m.lower)[None, :] + idx * np.asarray(m.h)[None, :]
        x = origin[:, None, :] + h_cell[:, None, :] * pts[None, :, :]
        cells, n_comp = (m.cell_u, dim) if space == 0 else (m.cell_p, 1)
        uh = np.einsum("qn,cnd->cqd", phi, vec.reshape(-1, n_comp)[cells])
        diff = uh - exact(x).reshape(uh.shape)
        return float(np.sqrt(np.einsum("cqd,cqd,q,c->", diff, diff, w, np.prod(h_cell, axis=1))))

    def apply_pressure_average_projection(self, p):
        if self.mode_w is not None and self.physical_type != STATIONARY:
            p -= float(self.mode_w @ p) / float(self.mode_w @ self.mode) * self.mode

    def vmult(self, src_u, src_p):
        dst_u, dst_p = np.zeros(self.n_u), np.zeros(self.n_p)
        self._loop("vmult", src_u, src_p, dst_u, dst_p)
        c = self.u_constrained.reshape(-1)                     # constrained rows, :247-256
        dst_u[c] = src_u[c]
        dst_p[self.p_constrained] = -src_p[self.p_constrained]
        self.apply_pressure_average_projection(dst_p)
        return dst_u, dst_p

    def velocity_vmult(self, src_u):
        dst_u = np.zeros(self.n_u)
        self._loop("vmult_velocity", src_u, None, dst_u, None)
        c = self.u_constrained.reshape(-1)
        dst_u[c] = src_u[c]
        return dst_u

    # ---- sub-block operators of the block preconditioner (navier_stokes_matrix.cc:300-483, cell kernels :920-1073), constant
    # coefficients
    def divergence_vmult_add(self, dst_p, src_u, weight_by_viscosity=False):
        """dst_p += int q (weight div u), weight = -1 or -viscosity, velocity quadrature, constraints resolved (:920-962)"""
        m, dim = self.m, self.dim
        ul = self._gather(src_u, m.cell_u, dim, self.Eu, m.u_dirichlet)
        grad = np.einsum("qnf,cnd->cqdf", self.dphi_u, ul) * self.K[:, None, None, :]
        div = np.einsum("cqdd->cq", grad)
        w = -self.mu if weight_by_viscosity else -1.0
        out_p = np.einsum("qn,cq->cn", self.phi_p, w * div * self.jxw)
        dst_p += self._scatter(self.Ep, out_p[:, :, None], None)[:, 0]

    def _pressure_quadrature(self):
        """pressure shape data on the pressure quadrature QGauss(degree_p + 1) (quad_index_p)"""
        k, dim = self.m.degree, self.dim
        xq, wq = gauss(k)
        vp, dp = lagrange_tables(gauss_lobatto(k), xq)
        phi = tensor(dim, [vp] * dim)
        dphi = np.stack([tensor(dim, [dp if f == d else vp for d in range(dim)]) for f in range(dim)], axis=-1)
        w = tensor(dim, [wq[:, None]] * dim)[:, 0]
        jxw = w[None, :] / np.prod(self.K, axis=1)[:, None]
        return phi, dphi, jxw

    def pressure_mass_vmult(self, src_p):
        """int phi (p / (mu + tau_grad_div)) on the pressure quadrature, constrained rows = identity (:420-455, :1041-1073)"""
        phi, _, jxw = self._pressure_quadrature()
        c = 1.0 if self.physical_type == STATIONARY else 1.0 / (self.mu + self.tau_gd)
        pl = self._gather(src_p, self.m.cell_p, 1, self.Ep)[:, :, 0]
        out = np.einsum("qn,cq->cn", phi, c * np.einsum("qn,cn->cq", phi, pl) * jxw)
        dst = self._scatter(self.Ep, out[:, :, None], None)[:, 0]
        dst[self.p_constrained] = src_p[self.p_constrained]
        return dst

    def pressure_poisson_vmult(self, src_p):
        """int grad phi . grad p / (weight rho) on the pressure quadrature, constrained rows = identity (:386-415, :966-1038)"""
        phi, dphi, jxw = self._pressure_quadrature()
        c = 1.0 if self.physical_type == STATIONARY else 1.0 / (self.weight * self.rho)
        pl = self._gather(src_p, self.m.cell_p, 1, self.Ep)[:, :, 0]
        g = np.einsum("qnf,cn->cqf", dphi, pl) * self.K[:, None, :]
        out = np.einsum("qnf,cqf->cn", dphi, c * g * self.K[:, None, :] * jxw[:, :, None])
        dst = self._scatter(self.Ep, out[:, :, None], None)[:, 0]
        dst[self.p_constrained] = src_p[self.p_constrained]
        return dst

    def residual(self, rhs_u, rhs_p, src_u, src_p, user_u, user_p):
        """rhs = user - cell integrals, navier_stokes_matrix.cc:266-293"""
        self._loop("residual", src_u, src_p, rhs_u, rhs_p)
        rhs_u *= -1.0
        rhs_u += user_u
        rhs_p *= -1.0
        rhs_p += user_p
