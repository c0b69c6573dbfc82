"""Synthetic hexahedral box meshes with Taylor-Hood Q2/Q1 DoF spaces.

This module stands in for the host side that deal.II / p4est own in the
reference (mesh, DoF numbering, constraints, partition).  It produces exactly
the arrays a deal.II adaptor would extract from ``MatrixFree`` and hand to the
C ABI (``include/adaflo_b200.h``): per-cell node index lists with Dirichlet
flag bits, hanging-node masks, geometry in compressed form and the ghost
exchange plan of the vector partitioner.

Reference conventions restated here (no reference code is used):
  * velocity FE = FESystem(FE_Q(QGaussLobatto(3)), dim), pressure FE = FE_Q(1)
    or FE_Q_DG0(1)                      -- source/navier_stokes.cc:92-106
  * quadrature QGauss(3) for the coupled operator, QGauss(2) for the pressure
    sub-operators                       -- source/navier_stokes.cc:447-448
  * cell-local DoF order is lexicographic (x fastest) -- MatrixFree's
    ``lexicographic_numbering``         -- SURVEY.md 8(c) item (1)
  * mesh partition: contiguous ranges of cells per rank, interface DoFs owned
    by the lower rank (p4est/deal.II rule) -- tests/beltrami.cc:196,209

Device vector layout chosen by this framework (not the reference's):
velocity vectors are node-major with the ``dim`` components interleaved
(``u[node*3 + c]``), so that one 32-bit node index serves three DoFs.
"""
from __future__ import annotations

import numpy as np

FLAG_SHIFT = 29
NODE_MASK = (1 << FLAG_SHIFT) - 1


# --------------------------------------------------------------------------
# 1-D building blocks
# --------------------------------------------------------------------------
def gauss_legendre_01(n: int):
    """Gauss-Legendre points/weights on [0,1] (deal.II ``QGauss<1>(n)``)."""
    x, w = np.polynomial.legendre.leggauss(n)
    # polish with two Newton steps in long double for last-bit reproducibility
    x = x.astype(np.longdouble)
    for _ in range(2):
        p0 = np.ones_like(x)
        p1 = x.copy()
        for k in range(2, n + 1):
            p0, p1 = p1, ((2 * k - 1) * x * p1 - (k - 1) * p0) / k
        dp = n * (x * p1 - p0) / (x * x - 1)
        x = x - p1 / dp
    p0 = np.ones_like(x)
    p1 = x.copy()
    for k in range(2, n + 1):
        p0, p1 = p1, ((2 * k - 1) * x * p1 - (k - 1) * p0) / k
    dp = n * (x * p1 - p0) / (x * x - 1)
    w = 2.0 / ((1 - x * x) * dp * dp)
    return np.asarray(0.5 * (x + 1), dtype=np.float64), np.asarray(0.5 * w, dtype=np.float64)


def gauss_lobatto_01(n: int):
    """Gauss-Lobatto points on [0,1] (support points of FE_Q(degree n-1))."""
    if n == 2:
        return np.array([0.0, 1.0])
    if n == 3:
        return np.array([0.0, 0.5, 1.0])
    # interior points = roots of P'_{n-1}
    k = n - 1
    c = np.zeros(k + 1)
    c[k] = 1
    r = np.polynomial.legendre.Legendre(c).deriv().roots()
    x = np.concatenate([[-1.0], np.sort(r.real), [1.0]])
    return 0.5 * (x + 1)


def lagrange_values(nodes: np.ndarray, x: np.ndarray):
    """Values and first derivatives of the Lagrange basis on ``nodes`` at ``x``.

    Returns (V, D) with V[q, i] = l_i(x_q), D[q, i] = l_i'(x_q).
    """
    nodes = np.asarray(nodes, dtype=np.longdouble)
    x = np.asarray(x, dtype=np.longdouble)
    n = len(nodes)
    V = np.ones((len(x), n), dtype=np.longdouble)
    D = np.zeros((len(x), n), dtype=np.longdouble)
    for i in range(n):
        denom = np.longdouble(1)
        for j in range(n):
            if j != i:
                denom *= nodes[i] - nodes[j]
        for q in range(len(x)):
            val = np.longdouble(1)
            for j in range(n):
                if j != i:
                    val *= x[q] - nodes[j]
            der = np.longdouble(0)
            for m in range(n):
                if m == i:
                    continue
                t = np.longdouble(1)
                for j in range(n):
                    if j != i and j != m:
                        t *= x[q] - nodes[j]
                der += t
            V[q, i] = val / denom
            D[q, i] = der / denom
    return np.asarray(V, dtype=np.float64), np.asarray(D, dtype=np.float64)


# --------------------------------------------------------------------------
# DoF space description (what the C ABI's space_upload consumes)
# --------------------------------------------------------------------------
class DofSpace:
    """One DoF space of one rank: local numbering = owned nodes, then ghosts."""

    def __init__(self, name, n_comp, nodes_per_cell):
        self.name = name
        self.n_comp = n_comp
        self.nodes_per_cell = nodes_per_cell
        self.n_owned = 0          # nodes
        self.n_ghost = 0          # nodes
        self.cell_nodes = None    # [n_cells, nodes_per_cell] uint32 local node index (no flags)
        self.dirichlet = None     # [n_local, n_comp] bool
        self.coords = None        # [n_local, 3] float64 support points
        self.global_ids = None    # [n_local] int64 global node id (partition independent)
        # ghost exchange plan: lists of (peer, local indices)
        self.send = []            # owned nodes that peer ghosts: (peer, idx int64[])
        self.recv = []            # (peer, first ghost local index, count)
        # hanging nodes (None on conforming meshes)
        self.hanging_mask = None  # [n_cells] uint32 refinement configuration
        self.n_dg0 = 0            # FE_Q_DG0: number of element-constant dofs (owned, appended)

    @property
    def n_local(self):
        return self.n_owned + self.n_ghost

    @property
    def n_dofs_owned(self):
        return self.n_owned * self.n_comp

    @property
    def n_dofs_local(self):
        return self.n_local * self.n_comp

    def encoded_cell_nodes(self):
        """cell->node list with the per-component Dirichlet flags in bits 29..31."""
        idx = self.cell_nodes.astype(np.uint32)
        assert self.n_local <= NODE_MASK
        flags = np.zeros(self.n_local, dtype=np.uint32)
        for c in range(self.n_comp):
            flags |= (self.dirichlet[:, c].astype(np.uint32) << np.uint32(FLAG_SHIFT + c))
        return idx | flags[self.cell_nodes]

    def constrained_dofs(self):
        """Locally owned constrained dof indices (vector entries), sorted.

        Mirrors ``MatrixFree::get_constrained_dofs`` used at
        source/navier_stokes_matrix.cc:250-255.
        """
        d = self.dirichlet[: self.n_owned].reshape(-1)
        return np.nonzero(d)[0].astype(np.uint32)


class BoxMesh:
    """Uniform Cartesian box of n = (nx, ny, nz) hexahedra, z-slab partition.

    ``rank``/``nranks`` select the slab; ``deform`` (callable X->x with
    jacobian) switches the geometry to the general per-quadrature-point path.
    """

    def __init__(self, n, lower=(-1.0, -1.0, -1.0), upper=(1.0, 1.0, 1.0),
                 rank=0, nranks=1, dirichlet_faces=(0, 1, 2, 3, 4, 5),
                 augmented=False, build_coords=True):
        self.dim = 3
        self.n = tuple(int(v) for v in n)
        self.lower = np.asarray(lower, dtype=np.float64)
        self.upper = np.asarray(upper, dtype=np.float64)
        self.h = (self.upper - self.lower) / np.asarray(self.n, dtype=np.float64)
        self.rank, self.nranks = rank, nranks
        self.augmented = augmented
        nx, ny, nz = self.n
        assert nz >= nranks
        # contiguous z-slabs (cells are lexicographic, so a slab is a contiguous cell range)
        bounds = [(nz * r) // nranks for r in range(nranks + 1)]
        self.z0, self.z1 = bounds[rank], bounds[rank + 1]
        self.n_cells = nx * ny * (self.z1 - self.z0)
        self.n_cells_global = nx * ny * nz
        self.u = self._build_space("u", 2, 3, dirichlet_faces, build_coords)
        self.p = self._build_space("p", 1, 1, (), build_coords)
        if augmented:
            self._append_dg0()

    # ---------------------------------------------------------------- spaces
    def _build_space(self, name, k, n_comp, dirichlet_faces, build_coords, equidistant=False):
        nx, ny, nz = self.n
        NX, NY, NZ = k * nx + 1, k * ny + 1, k * nz + 1
        z0, z1 = self.z0, self.z1
        # owned node planes: interface plane belongs to the lower rank
        zp0 = 0 if self.rank == 0 else k * z0 + 1
        zp1 = k * z1                     # inclusive
        n_planes = zp1 - zp0 + 1
        plane = NX * NY
        sp = DofSpace(name, n_comp, (k + 1) ** 3)
        sp.grid = (NX, NY, NZ)
        sp.n_owned = plane * n_planes
        sp.n_ghost = plane if self.rank > 0 else 0
        sp.n_owned_global = NX * NY * NZ

        # local index of a global node (ix,iy,iz): owned -> lexicographic in owned range;
        # ghost plane (iz == k*z0, rank>0) -> appended
        def local_index(ix, iy, iz):
            owned = (iz - zp0) * plane + iy * NX + ix
            ghost = sp.n_owned + iy * NX + ix
            return np.where(iz >= zp0, owned, ghost)

        # cells, lexicographic x fastest
        cz, cy, cx = np.meshgrid(np.arange(z0, z1), np.arange(ny), np.arange(nx), indexing="ij")
        cx, cy, cz = cx.reshape(-1), cy.reshape(-1), cz.reshape(-1)
        a = np.arange(k + 1)
        lz, ly, lx = np.meshgrid(a, a, a, indexing="ij")
        lx, ly, lz = lx.reshape(-1), ly.reshape(-1), lz.reshape(-1)
        gx = k * cx[:, None] + lx[None, :]
        gy = k * cy[:, None] + ly[None, :]
        gz = k * cz[:, None] + lz[None, :]
        sp.cell_nodes = local_index(gx, gy, gz).astype(np.uint32)
        self.cell_coords = np.stack([cx, cy, cz], axis=1)

        # global ids / coords / Dirichlet flags of local nodes
        iz_o, iy_o, ix_o = np.meshgrid(np.arange(zp0, zp1 + 1), np.arange(NY), np.arange(NX), indexing="ij")
        ix_l, iy_l, iz_l = ix_o.reshape(-1), iy_o.reshape(-1), iz_o.reshape(-1)
        if sp.n_ghost:
            iy_g, ix_g = np.meshgrid(np.arange(NY), np.arange(NX), indexing="ij")
            ix_l = np.concatenate([ix_l, ix_g.reshape(-1)])
            iy_l = np.concatenate([iy_l, iy_g.reshape(-1)])
            iz_l = np.concatenate([iz_l, np.full(plane, k * z0)])
        sp.global_ids = (ix_l + NX * (iy_l + NY * iz_l)).astype(np.int64)
        if build_coords:
            # FE_Q support points: Gauss-Lobatto inside each cell (equidistant for k<=2)
            assert k <= 2 or equidistant
            sp.coords = np.stack([self.lower[0] + self.h[0] * ix_l / k,
                                  self.lower[1] + self.h[1] * iy_l / k,
                                  self.lower[2] + self.h[2] * iz_l / k], axis=1)
        on_face = [ix_l == 0, ix_l == NX - 1, iy_l == 0, iy_l == NY - 1, iz_l == 0, iz_l == NZ - 1]
        d = np.zeros(len(ix_l), dtype=bool)
        for f in dirichlet_faces:
            d |= on_face[f]
        sp.dirichlet = np.repeat(d[:, None], n_comp, axis=1)

        # ghost exchange plan (slab partition: one peer below, one above)
        if self.rank > 0:
            sp.recv.append((self.rank - 1, sp.n_owned, plane))
        if self.rank < self.nranks - 1:
            top = (zp1 - zp0) * plane + np.arange(plane, dtype=np.int64)
            sp.send.append((self.rank + 1, top))
        return sp

    def _append_dg0(self):
        """FE_Q_DG0(1): one element-constant pressure dof per cell, appended to
        the owned range (source/navier_stokes.cc:98-106)."""
        sp = self.p
        # re-number: owned Q1 nodes, owned DG0 dofs, ghosts
        n_q1_owned = sp.n_owned
        shift = self.n_cells
        cn = sp.cell_nodes.astype(np.int64)
        cn = np.where(cn >= n_q1_owned, cn + shift, cn)
        dg = n_q1_owned + np.arange(self.n_cells, dtype=np.int64)
        sp.cell_nodes = np.concatenate([cn, dg[:, None]], axis=1).astype(np.uint32)
        sp.nodes_per_cell = 9
        sp.n_dg0 = self.n_cells
        sp.n_q1_owned = n_q1_owned
        # global ids: DG0 dofs numbered after all Q1 nodes
        nx, ny = self.n[0], self.n[1]
        cell_gid = sp.n_owned_global + (self.cell_coords[:, 0] + nx * (self.cell_coords[:, 1] + ny * self.cell_coords[:, 2]))
        sp.global_ids = np.concatenate([sp.global_ids[:n_q1_owned], cell_gid, sp.global_ids[n_q1_owned:]])
        ctr = self.lower[None, :] + self.h[None, :] * (self.cell_coords + 0.5)
        if sp.coords is not None:
            sp.coords = np.concatenate([sp.coords[:n_q1_owned], ctr, sp.coords[n_q1_owned:]])
        sp.dirichlet = np.concatenate([sp.dirichlet[:n_q1_owned], np.zeros((self.n_cells, 1), bool),
                                       sp.dirichlet[n_q1_owned:]])
        sp.n_owned = n_q1_owned + self.n_cells
        sp.recv = [(peer, first + shift, cnt) for (peer, first, cnt) in sp.recv]

    def build_ls_space(self, n_sub=3, dirichlet_faces=(), build_coords=True):
        """Level-set space FE_Q_iso_Q1(n_sub) (source/level_set_base.cc:54-60): (n_sub+1)^3 nodes per
        cell on the uniform (n_sub n + 1)^3 lattice, lexicographic in the cell."""
        self.ls = self._build_space("ls", n_sub, 1, dirichlet_faces, build_coords, equidistant=True)
        self.ls.n_sub = n_sub
        return self.ls

    def cell_diameters(self):
        """compute_cell_diameters (include/adaflo/util.h:47-120): for Cartesian cells the largest edge"""
        return np.full(self.n_cells, float(self.h.max()))

    # -------------------------------------------------------------- geometry
    def cartesian_geometry(self):
        """Compressed Cartesian geometry: one table row (1/hx,1/hy,1/hz,det) and a
        per-cell row index (MatrixFree detects box meshes as Cartesian,
        source/navier_stokes.cc:60-63 / SURVEY appendix A)."""
        table = np.array([[1.0 / self.h[0], 1.0 / self.h[1], 1.0 / self.h[2], np.prod(self.h)]])
        index = np.zeros(self.n_cells, dtype=np.uint32)
        return table, index

    def general_geometry(self, n_q_1d=3, deform=None, rule=None):
        """Per-quadrature-point inverse Jacobian K[f][e]=d xi_f/d x_e and JxW.

        Returns invJ [n_cells, n_q, 3, 3] and JxW [n_cells, n_q]; ``deform`` is
        ``(X) -> (x, dx/dX)`` on physical box coordinates (identity if None).
        """
        qp, qw = gauss_legendre_01(n_q_1d) if rule is None else rule   # rule: 1-D points / weights on [0, 1]
        n_q_1d = len(qp)
        qz, qy, qx = np.meshgrid(np.arange(n_q_1d), np.arange(n_q_1d), np.arange(n_q_1d), indexing="ij")
        qx, qy, qz = qx.reshape(-1), qy.reshape(-1), qz.reshape(-1)
        w = qw[qx] * qw[qy] * qw[qz]
        xi = np.stack([qp[qx], qp[qy], qp[qz]], axis=1)                        # [nq,3]
        X = self.lower[None, None, :] + self.h[None, None, :] * (self.cell_coords[:, None, :] + xi[None, :, :])
        nq = len(w)
        J = np.zeros((self.n_cells, nq, 3, 3))
        if deform is None:
            for d in range(3):
                J[:, :, d, d] = self.h[d]
        else:
            _, F = deform(X.reshape(-1, 3))                                   # F[e][f] = dx_e/dX_f
            J = F.reshape(self.n_cells, nq, 3, 3) * self.h[None, None, None, :]  # dx_e/dxi_f
        det = np.linalg.det(J)
        K = np.linalg.inv(J)                                                  # K[f][e] = dxi_f/dx_e
        return np.ascontiguousarray(K), det * w[None, :]

    def iterated_geometry(self, n_sub, deform=None):
        """geometry on QIterated(QGauss(2), n_sub), the level-set quadrature (source/two_phase_base.cc:265-268)"""
        g, w = gauss_legendre_01(2)
        pts = np.concatenate([(s + np.asarray(g, dtype=np.float64)) / n_sub for s in range(n_sub)])
        wts = np.concatenate([np.asarray(w, dtype=np.float64) / n_sub for _ in range(n_sub)])
        return self.general_geometry(2 * n_sub, deform, rule=(pts, wts))

    def quadrature_points(self, n_q_1d=3, deform=None):
        qp, _ = gauss_legendre_01(n_q_1d)
        qz, qy, qx = np.meshgrid(np.arange(n_q_1d), np.arange(n_q_1d), np.arange(n_q_1d), indexing="ij")
        xi = np.stack([qp[qx.reshape(-1)], qp[qy.reshape(-1)], qp[qz.reshape(-1)]], axis=1)
        X = self.lower[None, None, :] + self.h[None, None, :] * (self.cell_coords[:, None, :] + xi[None, :, :])
        if deform is not None:
            x, _ = deform(X.reshape(-1, 3))
            X = x.reshape(X.shape)
        return X


def sine_deformation(amplitude=0.05):
    """Smooth invertible deformation of the box used to force the general-geometry path."""
    def deform(X):
        x = X.copy()
        F = np.zeros((X.shape[0], 3, 3))
        for d in range(3):
            F[:, d, d] = 1.0
        s = np.sin(np.pi * X)
        c = np.cos(np.pi * X) * np.pi
        # x_e += A * sin(pi X_{e+1}) * sin(pi X_{e+2})
        for e in range(3):
            a, b = (e + 1) % 3, (e + 2) % 3
            x[:, e] += amplitude * s[:, a] * s[:, b]
            F[:, e, a] += amplitude * c[:, a] * s[:, b]
            F[:, e, b] += amplitude * s[:, a] * c[:, b]
        return x, F
    return deform


# --------------------------------------------------------------------------
# deterministic, partition independent test vectors
# --------------------------------------------------------------------------
def seeded_uniform(global_ids: np.ndarray, n_comp: int, seed: int):
    """U(-1,1) values keyed by global dof id (counter based, so any partition
    reproduces the same global vector) -- SURVEY.md 8(d) 'Synthetic inputs'."""
    gid = (global_ids[:, None].astype(np.uint64) * np.uint64(n_comp) + np.arange(n_comp, dtype=np.uint64)[None, :])
    x = gid.reshape(-1) + np.uint64((int(seed) * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF)
    # splitmix64
    x = (x + np.uint64(0x9E3779B97F4A7C15))
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (2.0 / (1 << 53)) - 1.0


def beltrami_velocity(x: np.ndarray, t=0.0, nu=1.0):
    """Ethier-Steinman Beltrami flow, the analytic field of tests/beltrami.cc:82-115."""
    a = 0.25 * np.pi
    d = 2.0 * a
    if x.shape[-1] == 2:   # the 2-D Taylor vortex of the same file, tests/beltrami.cc:109-114
        e = np.exp(-2.0 * nu * a * a * t)
        return np.stack([-a * np.cos(a * x[..., 0]) * np.sin(a * x[..., 1]) * e, a * np.sin(a * x[..., 0]) * np.cos(a * x[..., 1]) * e], axis=-1)
    X, Y, Z = x[..., 0], x[..., 1], x[..., 2]
    e = np.exp(-nu * d * d * t)
    u0 = -a * (np.exp(a * X) * np.sin(a * Y + d * Z) + np.exp(a * Z) * np.cos(a * X + d * Y)) * e
    u1 = -a * (np.exp(a * Y) * np.sin(a * Z + d * X) + np.exp(a * X) * np.cos(a * Y + d * Z)) * e
    u2 = -a * (np.exp(a * Z) * np.sin(a * X + d * Y) + np.exp(a * Y) * np.cos(a * Z + d * X)) * e
    return np.stack([u0, u1, u2], axis=-1)


def beltrami_pressure(x: np.ndarray, t=0.0, nu=1.0):
    """Pressure of the Beltrami flow, tests/beltrami.cc:140-160."""
    a = 0.25 * np.pi
    d = 2.0 * a
    if x.shape[-1] == 2:   # tests/beltrami.cc:161-164
        return -a * a * 0.25 * (np.cos(2 * a * x[..., 0]) + np.cos(2 * a * x[..., 1])) * np.exp(-4.0 * nu * a * a * t)
    X, Y, Z = x[..., 0], x[..., 1], x[..., 2]
    val = -a * a * 0.5 * (np.exp(2 * a * X) + np.exp(2 * a * Y) + np.exp(2 * a * Z)
                          + 2 * np.sin(a * X + d * Y) * np.cos(a * Z + d * X) * np.exp(a * (Y + Z))
                          + 2 * np.sin(a * Y + d * Z) * np.cos(a * X + d * Y) * np.exp(a * (Z + X))
                          + 2 * np.sin(a * Z + d * X) * np.cos(a * Y + d * Z) * np.exp(a * (X + Y)))
    return val * np.exp(-2 * nu * d * d * t)


def linear_deformation(shear=0.15):
    """Affine (constant Jacobian) deformation used to exercise the affine geometry path."""
    A = np.array([[1.0, shear, 0.5 * shear], [0.0, 1.0 + shear, -shear], [0.25 * shear, 0.0, 1.0]])

    def deform(X):
        x = X @ A.T
        F = np.broadcast_to(A, (X.shape[0], 3, 3)).copy()
        return x, F
    return deform


# --------------------------------------------------------------------------
# two-level meshes with 2:1 hanging nodes (config spurious_currents_ls_3d uses AMR)
# --------------------------------------------------------------------------
# Refinement configuration of a cell (adaflo_b200_space_desc.hanging_masks), the information
# deal.II keeps per cell for its compressed hanging-node scheme: position of the cell in its
# parent per direction, which of its three outer faces are hanging, and which outer edges are
# hanging without an adjacent hanging face.
HN_SUBCELL_X, HN_SUBCELL_Y, HN_SUBCELL_Z = 1, 2, 4
HN_FACE_X, HN_FACE_Y, HN_FACE_Z = 8, 16, 32
HN_EDGE_X, HN_EDGE_Y, HN_EDGE_Z = 64, 128, 256


class BoxMesh2D:
    """2-D box of n = (nx, ny) Q2/Q1 cells and its one-cell-thick 3-D extrusion.

    The library has 3-D kernels only; a 2-D operator is applied exactly as A_2d = R A_3d E on the extruded mesh
    (`mesh3d`, thickness 1): E copies a 2-D field to the 3 (velocity) / 2 (pressure) node planes with w = 0, R sums
    the planes -- the z basis functions sum to 1, so every 3-D bilinear form reduces to the 2-D one for z-constant
    fields ("2-D problems").  The third velocity component is constrained everywhere.
    Vectors: velocity entry (node, c) = 2 node + c, nodes lexicographic (x fastest)."""

    def __init__(self, n, lower=(-1.0, -1.0), upper=(1.0, 1.0), dirichlet_edges=(0, 1, 2, 3)):
        self.dim = 2
        self.n = (int(n[0]), int(n[1]))
        self.lower = np.asarray(lower, dtype=np.float64)
        self.upper = np.asarray(upper, dtype=np.float64)
        self.h = (self.upper - self.lower) / np.asarray(self.n, dtype=np.float64)
        nx, ny = self.n
        self.n_cells = nx * ny
        self.mesh3d = m3 = BoxMesh((nx, ny, 1), lower=(self.lower[0], self.lower[1], 0.0), upper=(self.upper[0], self.upper[1], 1.0),
                                   dirichlet_faces=tuple(dirichlet_edges))
        m3.u.dirichlet[:, 2] = True                         # w = 0: never read, never written
        NXu, NYu, NXp, NYp = 2 * nx + 1, 2 * ny + 1, nx + 1, ny + 1
        self.n_u_nodes, self.n_p_nodes = NXu * NYu, NXp * NYp
        # 2-D cell lists = bottom plane of the extruded lists (node3d = plane * k + node2d)
        self.cell_u = m3.u.cell_nodes[:, :9].astype(np.int64)
        self.cell_p = m3.p.cell_nodes[:, :4].astype(np.int64)
        assert self.cell_u.max() < self.n_u_nodes and self.cell_p.max() < self.n_p_nodes
        self.u_dirichlet = m3.u.dirichlet[: self.n_u_nodes, :2].copy()
        self.u_coords = m3.u.coords[: self.n_u_nodes, :2].copy()
        self.p_coords = m3.p.coords[: self.n_p_nodes, :2].copy()
        self.u_ids = np.arange(self.n_u_nodes, dtype=np.int64)
        self.p_ids = np.arange(self.n_p_nodes, dtype=np.int64)
        # extension maps: 3-D vector entries of a 2-D entry
        nodes = np.arange(self.n_u_nodes)
        self.ext_u = (3 * (nodes[:, None, None] + self.n_u_nodes * np.arange(3)[None, :, None]) + np.arange(2)[None, None, :])
        self.ext_u = self.ext_u.transpose(0, 2, 1).reshape(self.n_u_nodes * 2, 3)       # [2 node + c][plane]
        self.ext_p = np.arange(self.n_p_nodes)[:, None] + self.n_p_nodes * np.arange(2)[None, :]

    def build_ls_space(self, n_sub=3, dirichlet_edges=()):
        """FE_Q_iso_Q1(n_sub) level-set space: (n_sub nx + 1)(n_sub ny + 1) nodes, extruded to n_sub + 1 node planes"""
        ls3 = self.mesh3d.build_ls_space(n_sub, dirichlet_faces=tuple(dirichlet_edges))
        nx, ny = self.n
        self.n_sub = n_sub
        self.n_ls_nodes = (n_sub * nx + 1) * (n_sub * ny + 1)
        n1 = n_sub + 1
        self.cell_ls = ls3.cell_nodes[:, : n1 * n1].astype(np.int64)
        assert self.cell_ls.max() < self.n_ls_nodes
        self.ls_dirichlet = ls3.dirichlet[: self.n_ls_nodes, 0].copy()
        self.ls_coords = ls3.coords[: self.n_ls_nodes, :2].copy()
        self.ls_ids = np.arange(self.n_ls_nodes, dtype=np.int64)
        self.ext_ls = np.arange(self.n_ls_nodes)[:, None] + self.n_ls_nodes * np.arange(n1)[None, :]
        return ls3

    def cell_diameters(self):
        return np.full(self.n_cells, float(self.h.max()))

    def extend_ls(self, v2d):
        v3 = np.zeros(self.mesh3d.ls.n_owned)
        v3[self.ext_ls] = np.asarray(v2d)[:, None]
        return v3

    def restrict_ls(self, v3d):
        return np.asarray(v3d)[self.ext_ls].sum(axis=1)

    def extend(self, u2d, p2d):
        m3 = self.mesh3d
        u3, p3 = np.zeros(m3.u.n_owned * 3), np.zeros(m3.p.n_owned)
        u3[self.ext_u] = np.asarray(u2d)[:, None]
        p3[self.ext_p] = np.asarray(p2d)[:, None]
        return u3, p3

    def restrict(self, u3d, p3d):
        return np.asarray(u3d)[self.ext_u].sum(axis=1), np.asarray(p3d)[self.ext_p].sum(axis=1)


class RefinedBoxMesh:
    """Box of n coarse cells where the coarse cells inside ``refine`` = ((i0,i1),(j0,j1),(k0,k1))
    are split into 8 children (one level, single rank).  Produces

      * the compressed description the C ABI takes: per-cell node lists in which the slots of a
        hanging face / edge hold the PARENT's face / edge nodes, plus the per-cell mask;
      * independently, the plain constraint expansion of every cell-local dof (CSR lists for the
        oracle), obtained by evaluating the coarse neighbour's Lagrange basis at the hanging node.

    Cells: unrefined coarse cells first (lexicographic), then the children (parent-major)."""

    def __init__(self, n, refine, lower=(-1.0, -1.0, -1.0), upper=(1.0, 1.0, 1.0), dirichlet_faces=(0, 1, 2, 3, 4, 5),
                 augmented=False):
        self.dim = 3
        self.n = tuple(int(v) for v in n)
        self.lower = np.asarray(lower, dtype=np.float64)
        self.upper = np.asarray(upper, dtype=np.float64)
        self.hc = (self.upper - self.lower) / np.asarray(self.n, dtype=np.float64)
        self.rank, self.nranks = 0, 1
        nx, ny, nz = self.n
        refined = np.zeros((nx, ny, nz), dtype=bool)
        (i0, i1), (j0, j1), (k0, k1) = refine
        refined[i0:i1, j0:j1, k0:k1] = True
        self.refined = refined
        # cells: (level, origin in units of the fine cell size, size in fine cells)
        cells = []
        for k in range(nz):
            for j in range(ny):
                for i in range(nx):
                    if not refined[i, j, k]:
                        cells.append((0, 2 * i, 2 * j, 2 * k, 0, 0, 0))
        for k in range(nz):
            for j in range(ny):
                for i in range(nx):
                    if refined[i, j, k]:
                        for c in range(8):
                            sx, sy, sz = c & 1, (c >> 1) & 1, c >> 2
                            cells.append((1, 2 * i + sx, 2 * j + sy, 2 * k + sz, sx, sy, sz))
        self.cells = np.array(cells, dtype=np.int64)
        self.n_cells = len(cells)
        self.n_cells_global = self.n_cells
        self.h_cell = np.where(self.cells[:, 0:1] == 0, self.hc[None, :], 0.5 * self.hc[None, :])
        self.origin = self.lower[None, :] + 0.5 * self.hc[None, :] * self.cells[:, 1:4]
        self.u, self.u_lists = self._build_space("u", 2, 3, dirichlet_faces)
        self.p, self.p_lists = self._build_space("p", 1, 1, ())
        self.augmented = augmented
        if augmented:
            self._append_dg0()

    def _append_dg0(self):
        """FE_Q_DG0(1): one element-constant pressure dof per cell behind the Q1 nodes (source/navier_stokes.cc:98-106);
        the hanging-node masks concern the 8 vertex slots only"""
        sp = self.p
        n_q1 = sp.n_owned
        dg = n_q1 + np.arange(self.n_cells, dtype=np.int64)
        sp.cell_nodes = np.concatenate([sp.cell_nodes.astype(np.int64), dg[:, None]], axis=1).astype(np.uint32)
        sp.nodes_per_cell = 9
        sp.n_dg0, sp.n_q1_owned = self.n_cells, n_q1
        sp.global_ids = np.concatenate([sp.global_ids, sp.global_ids.max() + 1 + np.arange(self.n_cells)])
        sp.coords = np.concatenate([sp.coords, self.origin + 0.5 * self.h_cell])
        sp.dirichlet = np.concatenate([sp.dirichlet, np.zeros((self.n_cells, 1), bool)])
        sp.n_owned = sp.n_owned_global = n_q1 + self.n_cells
        # plain expansion lists: 9 rows per cell, the last one is the cell's own constant
        ptr, idx, w = self.p_lists
        nptr, nidx, nw = [0], [], []
        for c in range(self.n_cells):
            for node in range(8):
                a, b = ptr[c * 8 + node], ptr[c * 8 + node + 1]
                nidx.extend(idx[a:b].tolist())
                nw.extend(w[a:b].tolist())
                nptr.append(len(nidx))
            nidx.append(int(dg[c]))
            nw.append(1.0)
            nptr.append(len(nidx))
        self.p_lists = (np.array(nptr, dtype=np.int64), np.array(nidx, dtype=np.int64), np.array(nw, dtype=np.float64))

    def _is_refined(self, i, j, k):
        nx, ny, nz = self.n
        if i < 0 or j < 0 or k < 0 or i >= nx or j >= ny or k >= nz:
            return None  # outside
        return bool(self.refined[i, j, k])

    def _build_space(self, name, deg, n_comp, dirichlet_faces, iso_q1=False):
        nx, ny, nz = self.n
        L = 2 * deg                      # lattice steps per coarse cell (fine nodes)
        NX, NY, NZ = L * nx + 1, L * ny + 1, L * nz + 1
        n1 = deg + 1
        npc = n1 ** 3
        lid = lambda x, y, z: x + NX * (y + NY * z)
        # lattice points of each cell's nodes
        step = np.where(self.cells[:, 0] == 0, 2, 1)               # lattice step between nodes of the cell
        a = np.arange(n1)
        lz, ly, lx = np.meshgrid(a, a, a, indexing="ij")
        lx, ly, lz = lx.reshape(-1), ly.reshape(-1), lz.reshape(-1)
        base = self.cells[:, 1:4] * deg                              # cell origin on the lattice
        px = base[:, 0:1] + step[:, None] * lx[None, :]
        py = base[:, 1:2] + step[:, None] * ly[None, :]
        pz = base[:, 2:3] + step[:, None] * lz[None, :]
        # a lattice point is hanging if it lies in the closure of an unrefined coarse cell without being its node
        hanging = np.zeros(NX * NY * NZ, dtype=bool)
        used = np.zeros(NX * NY * NZ, dtype=bool)
        used[lid(px, py, pz).reshape(-1)] = True
        for k in range(nz):
            for j in range(ny):
                for i in range(nx):
                    if self.refined[i, j, k]:
                        continue
                    xs, ys, zs = np.meshgrid(np.arange(L * i, L * i + L + 1), np.arange(L * j, L * j + L + 1),
                                             np.arange(L * k, L * k + L + 1), indexing="ij")
                    odd = (xs % 2 == 1) | (ys % 2 == 1) | (zs % 2 == 1)
                    hanging[lid(xs[odd], ys[odd], zs[odd])] = True
        active = used & ~hanging
        number = -np.ones(NX * NY * NZ, dtype=np.int64)
        number[active] = np.arange(active.sum())
        sp = DofSpace(name, n_comp, npc)
        sp.n_owned = int(active.sum())
        sp.n_ghost = 0
        sp.n_owned_global = sp.n_owned
        sp.global_ids = np.nonzero(active)[0].astype(np.int64)      # lattice id = partition independent id
        az, ay, ax = np.unravel_index(sp.global_ids, (NZ, NY, NX))
        hf = self.hc / L
        sp.coords = np.stack([self.lower[0] + hf[0] * ax, self.lower[1] + hf[1] * ay, self.lower[2] + hf[2] * az], axis=1)
        on_face = [ax == 0, ax == NX - 1, ay == 0, ay == NY - 1, az == 0, az == NZ - 1]
        d = np.zeros(sp.n_owned, dtype=bool)
        for f in dirichlet_faces:
            d |= on_face[f]
        sp.dirichlet = np.repeat(d[:, None], n_comp, axis=1)

        # ---- compressed lists + masks
        cell_nodes = number[lid(px, py, pz)]                         # -1 on hanging slots
        masks = np.zeros(self.n_cells, dtype=np.uint32)
        for c in range(self.n_cells):
            lvl, ox, oy, oz, sx, sy, sz = self.cells[c]
            if lvl == 0:
                continue
            ci, cj, ck = ox // 2, oy // 2, oz // 2                  # parent coarse cell
            s = (sx, sy, sz)
            out = [(-1 if s[d_] == 0 else 1) for d_ in range(3)]    # outward direction of the outer faces
            nb = lambda dx, dy, dz: self._is_refined(ci + dx, cj + dy, ck + dz)
            face = [nb(out[0], 0, 0) is False, nb(0, out[1], 0) is False, nb(0, 0, out[2]) is False]
            # an outer edge along d is hanging if any coarse cell around it is unrefined
            edge = []
            for d_ in range(3):
                e, f = (d_ + 1) % 3, (d_ + 2) % 3
                off = [0, 0, 0]
                around = []
                for se in (0, out[e]):
                    for sf in (0, out[f]):
                        off[e], off[f] = se, sf
                        around.append(nb(*off))
                edge.append(any(r is False for r in around) and not face[e] and not face[f])
            m = sx * HN_SUBCELL_X + sy * HN_SUBCELL_Y + sz * HN_SUBCELL_Z
            m |= face[0] * HN_FACE_X | face[1] * HN_FACE_Y | face[2] * HN_FACE_Z
            m |= edge[0] * HN_EDGE_X | edge[1] * HN_EDGE_Y | edge[2] * HN_EDGE_Z
            if not (any(face) or any(edge)):
                continue
            masks[c] = m
            # parent's nodes on the lattice (step 2)
            pbase = np.array([ci, cj, ck]) * L
            outer = [0 if s[d_] == 0 else deg for d_ in range(3)]   # local index of the outer face
            for node in range(npc):
                l = (node % n1, (node // n1) % n1, node // (n1 * n1))
                onf = [face[d_] and l[d_] == outer[d_] for d_ in range(3)]
                one = [edge[d_] and l[(d_ + 1) % 3] == outer[(d_ + 1) % 3] and l[(d_ + 2) % 3] == outer[(d_ + 2) % 3]
                       for d_ in range(3)]
                if not (any(onf) or any(one)):
                    continue
                # slot holds the parent's node with the same local index on that face / edge; directions
                # that are not interpolated keep the cell's own lattice position
                interp_dir = [False, False, False]
                for d_ in range(3):
                    if onf[d_]:
                        interp_dir[(d_ + 1) % 3] = interp_dir[(d_ + 2) % 3] = True
                    if one[d_]:
                        interp_dir[d_] = True
                pos = []
                for d_ in range(3):
                    if interp_dir[d_]:
                        pos.append(pbase[d_] + 2 * l[d_])            # parent's node l[d] along d
                    else:
                        pos.append(base[c, d_] + l[d_])              # own position (lies on the parent's boundary)
                nid = number[lid(pos[0], pos[1], pos[2])]
                assert nid >= 0, (name, c, node, pos)
                cell_nodes[c, node] = nid
        assert cell_nodes.min() >= 0
        sp.cell_nodes = cell_nodes.astype(np.uint32)
        sp.hanging_mask = masks

        # ---- independent plain expansion: hanging value = coarse neighbour's interpolant at the point
        gl = gauss_lobatto_01(n1) if not iso_q1 else None
        ptr = [0]
        idx, wts = [], []
        pts = lid(px, py, pz)
        for c in range(self.n_cells):
            for comp in range(n_comp):
                for node in range(npc):
                    p_l = pts[c, node]
                    if number[p_l] >= 0:
                        idx.append(number[p_l] * n_comp + comp)
                        wts.append(1.0)
                    else:
                        x, y, z = px[c, node], py[c, node], pz[c, node]
                        owner = None
                        for ci in {(x - 1) // L, min(x // L, nx - 1)}:
                            for cj in {(y - 1) // L, min(y // L, ny - 1)}:
                                for ck in {(z - 1) // L, min(z // L, nz - 1)}:
                                    if self._is_refined(ci, cj, ck) is False and L * ci <= x <= L * ci + L and \
                                            L * cj <= y <= L * cj + L and L * ck <= z <= L * ck + L:
                                        owner = (ci, cj, ck)
                        assert owner is not None
                        ref = [(x - L * owner[0]) / L, (y - L * owner[1]) / L, (z - L * owner[2]) / L]
                        if iso_q1:   # FE_Q_iso_Q1: piecewise linear hats on deg sub-intervals
                            V = [np.maximum(0.0, 1.0 - np.abs(r * deg - np.arange(n1))) for r in ref]
                        else:
                            V = [lagrange_values(gl, np.array([r]))[0][0] for r in ref]
                        for kz in range(n1):
                            for ky in range(n1):
                                for kx in range(n1):
                                    w = V[0][kx] * V[1][ky] * V[2][kz]
                                    if abs(w) > 1e-15:
                                        q = number[lid(L * owner[0] + 2 * kx, L * owner[1] + 2 * ky, L * owner[2] + 2 * kz)]
                                        assert q >= 0
                                        idx.append(q * n_comp + comp)
                                        wts.append(w)
                    ptr.append(len(idx))
        lists = (np.array(ptr, dtype=np.int64), np.array(idx, dtype=np.int64), np.array(wts, dtype=np.float64))
        return sp, lists

    def build_ls_space(self, n_sub=3, dirichlet_faces=()):
        """level-set space FE_Q_iso_Q1(n_sub) with its hanging-node masks and plain expansion lists"""
        self.ls, self.ls_lists = self._build_space("ls", n_sub, 1, dirichlet_faces, iso_q1=True)
        self.ls.n_sub = n_sub
        return self.ls

    def cell_diameters(self):
        return self.h_cell.max(axis=1)

    # ------------------------------------------------------------ geometry
    def cartesian_geometry(self):
        table = np.array([[1.0 / self.hc[0], 1.0 / self.hc[1], 1.0 / self.hc[2], np.prod(self.hc)],
                          [2.0 / self.hc[0], 2.0 / self.hc[1], 2.0 / self.hc[2], np.prod(self.hc) / 8.0]])
        return table, self.cells[:, 0].astype(np.uint32)

    def general_geometry(self, n_q_1d=3, deform=None):
        assert deform is None
        qp, qw = gauss_legendre_01(n_q_1d)
        qz, qy, qx = np.meshgrid(np.arange(n_q_1d), np.arange(n_q_1d), np.arange(n_q_1d), indexing="ij")
        w = (qw[qx] * qw[qy] * qw[qz]).reshape(-1)
        nq = len(w)
        K = np.zeros((self.n_cells, nq, 3, 3))
        for d_ in range(3):
            K[:, :, d_, d_] = 1.0 / self.h_cell[:, d_:d_ + 1]
        det = np.prod(self.h_cell, axis=1)
        return np.ascontiguousarray(K), det[:, None] * w[None, :]

    def quadrature_points(self, n_q_1d=3, deform=None):
        qp, _ = gauss_legendre_01(n_q_1d)
        qz, qy, qx = np.meshgrid(np.arange(n_q_1d), np.arange(n_q_1d), np.arange(n_q_1d), indexing="ij")
        xi = np.stack([qp[qx.reshape(-1)], qp[qy.reshape(-1)], qp[qz.reshape(-1)]], axis=1)
        return self.origin[:, None, :] + self.h_cell[:, None, :] * xi[None, :, :]


class TensorMesh:
    """Uniform box of n cells per direction in dim = 2 or 3 with Taylor-Hood Q_k / Q_(k-1) spaces (k = degree >= 2):
    the meshes of the degree / dimension dispatch (csrc/ns_generic.cu), e.g. Q4/Q3 in 2-D like tests/beltrami_2d.prm.
    Nodes are global tensor grids (lexicographic, x fastest), cell-local nodes lexicographic; FE_Q(k) nodes sit on the
    Gauss-Lobatto points of each cell (deal.II's FE_Q).  Velocity entry (node, c) = dim node + c; homogeneous Dirichlet flags on
    all boundary faces."""

    def __init__(self, dim, n, degree, lower=None, upper=None):
        assert dim in (2, 3) and degree >= 2
        self.dim, self.degree = dim, degree
        self.n = tuple(int(v) for v in n)
        assert len(self.n) == dim
        self.lower = np.asarray(lower if lower is not None else (-1.0,) * dim, dtype=np.float64)
        self.upper = np.asarray(upper if upper is not None else (1.0,) * dim, dtype=np.float64)
        self.h = (self.upper - self.lower) / np.asarray(self.n, dtype=np.float64)
        self.n_cells = int(np.prod(self.n))
        self.cell_u, self.n_u_nodes, self.u_coords, self.u_boundary = self._space(degree)
        self.cell_p, self.n_p_nodes, self.p_coords, _ = self._space(degree - 1)
        self.u_dirichlet = np.repeat(self.u_boundary[:, None], dim, axis=1)
        self.u_ids = np.arange(self.n_u_nodes, dtype=np.int64)
        self.p_ids = np.arange(self.n_p_nodes, dtype=np.int64)

    @staticmethod
    def _gll(n):
        from numpy.polynomial import legendre as npleg
        if n == 2:
            return np.array([0.0, 1.0])
        c = np.zeros(n)
        c[-1] = 1.0
        return 0.5 * (np.concatenate([[-1.0], np.sort(npleg.legroots(npleg.legder(c))), [1.0]]) + 1.0)

    def _space(self, k):
        dim, n = self.dim, self.n
        N = [n[d] * k + 1 for d in range(dim)]
        gll = self._gll(k + 1)
        # node index along direction d of local node a in cell i: i k + a
        local = np.arange(k + 1)
        axes = [np.arange(n[d])[:, None] * k + local[None, :] for d in range(dim)]          # [cell_d, a]
        x1d = []
        for d in range(dim):
            x = np.zeros(N[d])
            for i in range(n[d]):
                x[i * k: i * k + k + 1] = self.lower[d] + self.h[d] * (i + gll)
            x1d.append(x)
        if dim == 2:
            g = axes[0][None, :, None, :] + N[0] * axes[1][:, None, :, None]                # [cy, cx, b, a]
            cells = g.reshape(n[1] * n[0], (k + 1) ** 2)
            X, Y = np.meshgrid(x1d[0], x1d[1], indexing="xy")
            coords = np.stack([X.reshape(-1), Y.reshape(-1)], axis=1)
            ix, iy = np.meshgrid(np.arange(N[0]), np.arange(N[1]), indexing="xy")
            bnd = ((ix == 0) | (ix == N[0] - 1) | (iy == 0) | (iy == N[1] - 1)).reshape(-1)
        else:
            g = (axes[0][None, None, :, None, None, :] + N[0] * axes[1][None, :, None, None, :, None]
                 + N[0] * N[1] * axes[2][:, None, None, :, None, None])                     # [cz, cy, cx, c, b, a]
            cells = g.reshape(n[2] * n[1] * n[0], (k + 1) ** 3)
            Z, Y, X = np.meshgrid(x1d[2], x1d[1], x1d[0], indexing="ij")
            coords = np.stack([X.reshape(-1), Y.reshape(-1), Z.reshape(-1)], axis=1)
            iz, iy, ix = np.meshgrid(np.arange(N[2]), np.arange(N[1]), np.arange(N[0]), indexing="ij")
            bnd = ((ix == 0) | (ix == N[0] - 1) | (iy == 0) | (iy == N[1] - 1) | (iz == 0) | (iz == N[2] - 1)).reshape(-1)
        return cells.astype(np.int64), int(np.prod(N)), coords, bnd

    def encoded_cell_u(self):
        """cell lists with the constraint flag bits of the C ABI (bit 29 + c: component c constrained)"""
        flags = np.zeros(self.n_u_nodes, dtype=np.uint32)
        for c in range(self.dim):
            flags |= (self.u_dirichlet[:, c].astype(np.uint32) << np.uint32(29 + c))
        return (self.cell_u.astype(np.uint32) | flags[self.cell_u]).astype(np.uint32)

    def constrained_u(self):
        return np.nonzero(self.u_dirichlet.reshape(-1))[0].astype(np.uint32)

    def cartesian_geometry(self):
        inv = np.ones(3)
        inv[: self.dim] = 1.0 / self.h
        return np.array([[inv[0], inv[1], inv[2], float(np.prod(self.h))]]), np.zeros(self.n_cells, dtype=np.uint32)


class RefinedTensorMesh2D:
    """2-D box of n coarse cells per direction whose rectangular block `refine` = ((x0, x1), (y0, y1)) of coarse cells is refined
    once more (2:1 hanging nodes on the block's faces towards active coarse cells), with Taylor-Hood Q_k / Q_(k-1) spaces of any
    degree: the mesh of tests/beltrami_2d (tests/beltrami.cc:385-412: a 4 x 4 grid refined twice, active cells #2 and #3 flagged,
    one more global refinement = a 32 x 32 grid whose cells [0, 4) x [2, 4) are refined; Q4/Q3).

    Like deal.II, every cell keeps its own dofs on a hanging face (the coarse cell's k + 1 face nodes; the two fine cells' 2 k + 1
    nodes, of which the two end points are the coarse vertices): the 2 k - 1 remaining fine nodes are *constrained* dofs that exist
    in the vectors (the DoF counts of beltrami_2d.output:3 include them) and carry one constraint row each,
    value = sum_j L_j(s) * coarse face node j, L_j = the Lagrange basis of the coarse face on the Gauss-Lobatto points and s the
    node's position on it -- the rows AffineConstraints holds after DoFTools::make_hanging_node_constraints.  Constrained nodes
    are numbered last (adaflo_b200_space_desc::n_hanging_nodes).  Cells: coarse cells lexicographically (without the refined
    ones), then the fine cells of the block lexicographically."""

    def __init__(self, n, refine, degree, lower=(-1.0, -1.0), upper=(1.0, 1.0)):
        assert degree >= 2
        self.dim, self.degree = 2, degree
        self.n = (int(n[0]), int(n[1]))
        (x0, x1), (y0, y1) = refine
        assert 0 <= x0 < x1 <= self.n[0] and 0 <= y0 < y1 <= self.n[1]
        self.refine = ((x0, x1), (y0, y1))
        self.lower, self.upper = np.asarray(lower, dtype=np.float64), np.asarray(upper, dtype=np.float64)
        self.h = (self.upper - self.lower) / np.asarray(self.n, dtype=np.float64)      # coarse cell size
        # cells in units of the fine cell size: (origin x, origin y, size)
        cells = [(2 * i, 2 * j, 2) for j in range(self.n[1]) for i in range(self.n[0]) if not (x0 <= i < x1 and y0 <= j < y1)]
        self.n_coarse = len(cells)
        cells += [(i, j, 1) for j in range(2 * y0, 2 * y1) for i in range(2 * x0, 2 * x1)]
        self.cells = np.asarray(cells, dtype=np.int64)
        self.n_cells = len(cells)
        self.h_cell = self.cells[:, 2:3] * (0.5 * self.h)[None, :]
        self.origin = self.lower[None, :] + self.cells[:, 0:2] * (0.5 * self.h)[None, :]
        self.cell_u, self.n_u_nodes, self.u_coords, self.u_boundary, self.u_rows = self._space(degree)
        self.cell_p, self.n_p_nodes, self.p_coords, _, self.p_rows = self._space(degree - 1)
        self.u_dirichlet = np.repeat(self.u_boundary[:, None], 2, axis=1)
        self.u_ids = np.arange(self.n_u_nodes, dtype=np.int64)
        self.p_ids = np.arange(self.n_p_nodes, dtype=np.int64)

    @staticmethod
    def _gll(n):
        """n Gauss-Lobatto points on [0, 1], exactly symmetric about 1/2"""
        from numpy.polynomial import legendre as npleg
        if n == 2:
            return np.array([0.0, 1.0])
        c = np.zeros(n)
        c[-1] = 1.0
        z = np.concatenate([[-1.0], np.sort(npleg.legroots(npleg.legder(c))), [1.0]])
        z = 0.5 * (z - z[::-1])
        return 0.5 * (z + 1.0)

    def _space(self, k):
        gll = self._gll(k + 1)
        (x0, x1), (y0, y1) = self.refine
        nx, ny = self.n
        scale = float(2 ** 40)
        key = lambda px, py: (int(round(px * scale)), int(round(py * scale)))
        coarse, fine = {}, {}
        pos = []                                                    # node position in fine-cell units
        lists = np.zeros((self.n_cells, (k + 1) ** 2), dtype=np.int64)
        for c in range(self.n_coarse):
            ox, oy, s = self.cells[c]
            for b in range(k + 1):
                for a in range(k + 1):
                    kk = key(ox + s * gll[a], oy + s * gll[b])
                    if kk not in coarse:
                        coarse[kk] = len(pos)
                        pos.append((ox + s * gll[a], oy + s * gll[b]))
                    lists[c, a + (k + 1) * b] = coarse[kk]
        hanging = {}                                                # node -> (parents, weights)
        for c in range(self.n_coarse, self.n_cells):
            ox, oy, s = self.cells[c]
            for b in range(k + 1):
                for a in range(k + 1):
                    px, py = ox + gll[a], oy + gll[b]
                    kk = key(px, py)
                    at_coarse_vertex = (a in (0, k)) and (b in (0, k)) and (ox + a // k) % 2 == 0 and (oy + b // k) % 2 == 0
                    if at_coarse_vertex and kk in coarse:           # a vertex shared with an active coarse cell: one dof
                        lists[c, a + (k + 1) * b] = coarse[kk]
                        continue
                    if kk not in fine:
                        fine[kk] = len(pos)
                        pos.append((px, py))
                        # on a face of the block with an active coarse neighbour?  (never at a corner of the block: those are
                        # coarse vertices, and a block face on the domain boundary has no neighbour)
                        face = None
                        if a == 0 and ox == 2 * x0 and x0 > 0:
                            face = (0, 2 * x0)
                        elif a == k and ox + 1 == 2 * x1 and x1 < nx:
                            face = (0, 2 * x1)
                        elif b == 0 and oy == 2 * y0 and y0 > 0:
                            face = (1, 2 * y0)
                        elif b == k and oy + 1 == 2 * y1 and y1 < ny:
                            face = (1, 2 * y1)
                        if face is not None:
                            d, X = face                              # face normal direction, its coordinate
                            t = py if d == 0 else px                 # position along the face
                            jc = int(np.floor(t / 2.0 + 1e-12))
                            if abs(t - 2 * jc) < 1e-12 and abs(t - round(t)) < 1e-12 and int(round(t)) % 2 == 0:
                                raise AssertionError("coarse vertex on a hanging face must have been merged")
                            sloc = (t - 2 * jc) / 2.0
                            wts = lagrange_values(gll, np.array([sloc]))[0][0]
                            par = []
                            for j in range(k + 1):
                                tj = 2 * jc + 2 * gll[j]
                                par.append(coarse[key(X, tj) if d == 0 else key(tj, X)])
                            keep = np.abs(wts) > 1e-14
                            hanging[fine[kk]] = (np.asarray(par)[keep], wts[keep])
                    lists[c, a + (k + 1) * b] = fine[kk]
        n_nodes = len(pos)
        # constrained nodes last
        is_h = np.zeros(n_nodes, dtype=bool)
        is_h[list(hanging.keys())] = True
        order = np.concatenate([np.nonzero(~is_h)[0], np.nonzero(is_h)[0]])
        new_id = np.empty(n_nodes, dtype=np.int64)
        new_id[order] = np.arange(n_nodes)
        lists = new_id[lists]
        P = np.asarray(pos)[order]
        coords = self.lower[None, :] + P * (0.5 * self.h)[None, :]
        bnd = (np.abs(P[:, 0]) < 1e-12) | (np.abs(P[:, 0] - 2 * nx) < 1e-12) | (np.abs(P[:, 1]) < 1e-12) | (np.abs(P[:, 1] - 2 * ny) < 1e-12)
        n_h = int(is_h.sum())
        ptr, cols, wts = [0], [], []
        for old in order[n_nodes - n_h:]:
            par, w = hanging[int(old)]
            cols += [int(new_id[q]) for q in par]
            wts += [float(v) for v in w]
            ptr.append(len(cols))
        assert not bnd[n_nodes - n_h:].any()
        rows = (np.asarray(ptr, dtype=np.uint32), np.asarray(cols, dtype=np.uint32), np.asarray(wts, dtype=np.float64))
        return lists, n_nodes, coords, bnd, rows

    @property
    def n_hanging_u(self):
        return len(self.u_rows[0]) - 1

    @property
    def n_hanging_p(self):
        return len(self.p_rows[0]) - 1

    def encoded_cell_u(self):
        flags = np.zeros(self.n_u_nodes, dtype=np.uint32)
        for c in range(2):
            flags |= (self.u_dirichlet[:, c].astype(np.uint32) << np.uint32(29 + c))
        return (self.cell_u.astype(np.uint32) | flags[self.cell_u]).astype(np.uint32)

    def encoded_rows_u(self):
        """constraint rows with the Dirichlet flag bits on the parents"""
        ptr, cols, w = self.u_rows
        flags = np.zeros(self.n_u_nodes, dtype=np.uint32)
        for c in range(2):
            flags |= (self.u_dirichlet[:, c].astype(np.uint32) << np.uint32(29 + c))
        return ptr, (cols | flags[cols]).astype(np.uint32), w

    def constrained_u(self):
        """MatrixFree::get_constrained_dofs: Dirichlet entries and the entries of the hanging nodes, sorted"""
        con = self.u_dirichlet.copy()
        con[self.n_u_nodes - self.n_hanging_u:, :] = True
        return np.nonzero(con.reshape(-1))[0].astype(np.uint32)

    def constrained_p(self):
        return np.arange(self.n_p_nodes - self.n_hanging_p, self.n_p_nodes, dtype=np.uint32)

    def distribute(self, vec, space):
        """AffineConstraints::distribute for the hanging-node rows of a nodal vector ([n_nodes * n_comp], node-major)"""
        ptr, cols, w = self.u_rows if space == 0 else self.p_rows
        n_nodes = self.n_u_nodes if space == 0 else self.n_p_nodes
        v = vec.reshape(n_nodes, -1)
        first = n_nodes - (len(ptr) - 1)
        for r in range(len(ptr) - 1):
            v[first + r] = (w[ptr[r]:ptr[r + 1], None] * v[cols[ptr[r]:ptr[r + 1]]]).sum(axis=0)
        return vec

    def cartesian_geometry(self):
        table = np.array([[2.0 / self.h[0] * f, 2.0 / self.h[1] * f, 1.0, self.h[0] * self.h[1] / (4.0 * f * f)] for f in (0.5, 1.0)])
        return table, (self.cells[:, 2] == 1).astype(np.uint32)

    def quadrature_points(self, n_q_1d):
        qp, _ = gauss_legendre_01(n_q_1d)
        qx, qy = np.meshgrid(qp, qp, indexing="xy")
        xi = np.stack([qx.reshape(-1), qy.reshape(-1)], axis=1)
        return self.origin[:, None, :] + self.h_cell[:, None, :] * xi[None, :, :]
