"""Host-side mirror of the reference's operator interface on top of the C ABI.

Class and method names follow adaflo (include/adaflo/navier_stokes_matrix.h:57-204,
include/adaflo/diagonal_preconditioner.h) so that tests read like calls into the
reference.  PyTorch is used for device memory and streams only; every operation
runs in libadaflo_b200.so (hand-written CUDA) -- there is no PyTorch or CPU
fallback path.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import capi
from .capi import FlowParams, TimeWeights, check


def _ptr(t):
    if t is None:
        return None
    if isinstance(t, torch.Tensor):
        assert t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()
        return C.c_void_p(t.data_ptr())
    raise TypeError(type(t))


def _np(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class _DevView:
    """Wrap a raw device pointer so torch can view it (no ownership)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False), "version": 3}


def device_view(ptr, n, device):
    return torch.as_tensor(_DevView(ptr, n), device=device)


class Context:
    """One rank / one GPU (adaflo_b200_ctx_t)."""

    def __init__(self, device=0, stream=None):
        self.lib = capi.load()
        if not torch.cuda.is_available():
            raise capi.AdafloB200Error("no CUDA device: adaflo_b200 has no CPU fallback")
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        if stream is None:
            stream = torch.cuda.current_stream(self.device)
        self.stream = stream
        h = C.c_void_p()
        check(self.lib.adaflo_b200_ctx_create(C.byref(h), device, C.c_void_p(stream.cuda_stream)))
        self.h = h
        self.rank, self.nranks = 0, 1

    def init_comm(self, rank, nranks, broadcast_bytes):
        """Create the NCCL communicator; ``broadcast_bytes(bytes|None) -> bytes``
        distributes rank 0's unique id (e.g. over torch.distributed)."""
        if nranks <= 1:
            return
        idbuf = (C.c_ubyte * 128)()
        if rank == 0:
            check(self.lib.adaflo_b200_comm_unique_id(idbuf))
        data = broadcast_bytes(bytes(idbuf) if rank == 0 else None)
        idbuf = (C.c_ubyte * 128).from_buffer_copy(data)
        check(self.lib.adaflo_b200_ctx_init_comm(self.h, idbuf, rank, nranks))
        self.rank, self.nranks = rank, nranks

    def synchronize(self):
        check(self.lib.adaflo_b200_ctx_synchronize(self.h))

    def close(self):
        if self.h:
            self.lib.adaflo_b200_ctx_destroy(self.h)
            self.h = None

    # ---- vector kernels (deal.II LinearAlgebra::distributed::Vector operations)
    def axpy(self, y, a, x):
        check(self.lib.adaflo_b200_vec_axpy(self.h, _ptr(y), a, _ptr(x), y.numel()))

    def sadd(self, y, s, a, x):
        check(self.lib.adaflo_b200_vec_sadd(self.h, _ptr(y), s, a, _ptr(x), y.numel()))

    def scale(self, y, scaling):
        check(self.lib.adaflo_b200_vec_scale(self.h, _ptr(y), _ptr(scaling), y.numel()))

    def set(self, x, value):
        check(self.lib.adaflo_b200_vec_set(self.h, _ptr(x), value, x.numel()))

    def dot(self, x, y, n=None):
        r = C.c_double()
        check(self.lib.adaflo_b200_vec_dot(self.h, _ptr(x), _ptr(y), n if n is not None else x.numel(), C.byref(r)))
        return r.value

    def multi_dot(self, x, ys, n=None):
        k = len(ys)
        arr = (C.c_void_p * k)(*[y.data_ptr() for y in ys])
        res = (C.c_double * k)()
        check(self.lib.adaflo_b200_vec_multi_dot(self.h, _ptr(x), arr, k, n if n is not None else x.numel(), res))
        return [res[i] for i in range(k)]

    def multi_axpy(self, y, coeffs, xs, n=None):
        """y += sum_i coeffs[i] * xs[i] in one pass over y."""
        k = len(xs)
        if k == 0:
            return
        arr = (C.c_void_p * k)(*[x.data_ptr() for x in xs])
        a = (C.c_double * k)(*[float(c) for c in coeffs])
        check(self.lib.adaflo_b200_vec_multi_axpy(self.h, _ptr(y), a, arr, k, n if n is not None else y.numel()))

    def add_and_dot(self, y, a, x, z, n=None):
        r = C.c_double()
        check(self.lib.adaflo_b200_vec_add_and_dot(self.h, _ptr(y), a, _ptr(x), _ptr(z), n if n is not None else y.numel(), C.byref(r)))
        return r.value


class MatrixFree:
    """Device mirror of dealii::MatrixFree as adaflo sets it up: dof handlers
    {0: u, 1: p}, quadratures {0: QGauss(3), 1: QGauss(2)}
    (source/navier_stokes.cc:414-451)."""

    def __init__(self, ctx: Context, mesh, geometry="cartesian", deform=None):
        self.ctx = ctx
        self.mesh = mesh
        lib = ctx.lib
        keep = []  # keep numpy arrays alive during the call
        ls_space = getattr(mesh, "ls", None)
        space_list = [mesh.u, mesh.p] + ([ls_space] if ls_space is not None else [])
        spaces = (capi.SpaceDesc * len(space_list))()
        for i, sp in enumerate(space_list):
            cn = np.ascontiguousarray(sp.encoded_cell_nodes())
            con = np.ascontiguousarray(sp.constrained_dofs())
            n_ex = len(sp.send) + len(sp.recv)
            peers = sorted(set([p for p, _ in sp.send] + [p for p, _, _ in sp.recv]))
            ex = (capi.Exchange * max(1, len(peers)))()
            for j, peer in enumerate(peers):
                send = [idx for p, idx in sp.send if p == peer]
                recv = [(f, n) for p, f, n in sp.recv if p == peer]
                ex[j].peer = peer
                if send:
                    s = np.ascontiguousarray(send[0].astype(np.uint32))
                    keep.append(s)
                    ex[j].n_send = len(s)
                    ex[j].send = s.ctypes.data
                if recv:
                    ex[j].recv_first, ex[j].n_recv = recv[0]
            keep += [cn, con, ex]
            d = spaces[i]
            d.fe_kind = capi.FE_Q_DG0 if getattr(sp, "n_dg0", 0) else (capi.FE_Q_ISO_Q1 if i == 2 else capi.FE_Q)
            d.degree = 2 if i == 0 else (1 if i == 1 else sp.n_sub)
            d.n_components = sp.n_comp
            d.n_owned_nodes = sp.n_owned
            d.n_ghost_nodes = sp.n_ghost
            d.nodes_per_cell = sp.nodes_per_cell
            d.cell_nodes = cn.ctypes.data
            hm = getattr(sp, "hanging_mask", None)
            if hm is not None:
                hm = np.ascontiguousarray(hm.astype(np.uint32))
                keep.append(hm)
                d.hanging_masks = hm.ctypes.data
            else:
                d.hanging_masks = None
            d.n_constrained = len(con)
            d.constrained_dofs = con.ctypes.data if len(con) else None
            d.n_exchange = len(peers) if n_ex else 0
            d.exchange = ex
        quad_sizes = [3, 2] + ([2 * ls_space.n_sub] if ls_space is not None else [])
        quads = (capi.QuadDesc * len(quad_sizes))()
        for i, nq1 in enumerate(quad_sizes):
            q = quads[i]
            q.n_q_1d = nq1
            if geometry == "cartesian":
                table, row = mesh.cartesian_geometry()
                q.geometry_type = capi.GEOM_CARTESIAN
            elif geometry == "affine":
                K, jxw = mesh.general_geometry(1, deform)   # one point per cell: the affine Jacobian
                det = jxw[:, 0]                                # weight of the 1-point rule is 1
                table = np.concatenate([K[:, 0].reshape(-1, 9), det[:, None]], axis=1)
                row = np.arange(mesh.n_cells, dtype=np.uint32)
                q.geometry_type = capi.GEOM_AFFINE
            elif geometry == "general":
                if i == 2:   # level-set quadrature: QIterated(QGauss(2), n_sub)
                    K, jxw = mesh.iterated_geometry(ls_space.n_sub, deform)
                else:
                    K, jxw = mesh.general_geometry(nq1, deform)
                invj = np.ascontiguousarray(K.reshape(mesh.n_cells, nq1 ** 3, 9).transpose(0, 2, 1))
                jxw = np.ascontiguousarray(jxw)
                keep += [invj, jxw]
                q.geometry_type = capi.GEOM_GENERAL
                q.inv_jacobian = invj.ctypes.data
                q.jxw = jxw.ctypes.data
                continue
            else:
                raise ValueError(geometry)
            table = np.ascontiguousarray(table)
            row = np.ascontiguousarray(row)
            keep += [table, row]
            q.n_table_rows = table.shape[0]
            q.table = table.ctypes.data
            q.cell_row = row.ctypes.data
        desc = capi.MFDesc(3, mesh.n_cells, len(space_list), spaces, len(quad_sizes), quads)
        h = C.c_void_p()
        check(lib.adaflo_b200_mf_create(ctx.h, C.byref(desc), C.byref(h)))
        self.h = h
        del keep

    def initialize_dof_vector(self, space):
        n_owned, n_local = C.c_uint64(), C.c_uint64()
        check(self.ctx.lib.adaflo_b200_mf_vector_size(self.h, space, C.byref(n_owned), C.byref(n_local)))
        return torch.zeros(n_local.value, dtype=torch.float64, device=self.ctx.device)

    def n_owned(self, space):
        n_owned = C.c_uint64()
        check(self.ctx.lib.adaflo_b200_mf_vector_size(self.h, space, C.byref(n_owned), None))
        return n_owned.value

    def set_dof_numbering(self, space, host_position):
        """host_position[i]: position of library entry i (node-major, components interleaved) in the host application's vector,
        e.g. after DoFRenumbering::Cuthill_McKee on the velocity DoFHandler (navier_stokes.cc:208-209); None: default"""
        hp = None if host_position is None else np.ascontiguousarray(host_position, dtype=np.uint32)
        check(self.ctx.lib.adaflo_b200_mf_set_dof_numbering(self.h, space, _np(hp)))

    def upload_vector(self, space, vec_dev, vec_host):
        check(self.ctx.lib.adaflo_b200_mf_upload_vector(self.h, space, _ptr(vec_dev), _np(np.ascontiguousarray(vec_host))))

    def download_vector(self, space, vec_host, vec_dev):
        check(self.ctx.lib.adaflo_b200_mf_download_vector(self.h, space, _np(vec_host), _ptr(vec_dev)))

    def update_ghost_values(self, space, vec):
        check(self.ctx.lib.adaflo_b200_mf_update_ghost_values(self.h, space, _ptr(vec)))

    def compress_add(self, space, vec):
        check(self.ctx.lib.adaflo_b200_mf_compress_add(self.h, space, _ptr(vec)))

    def close(self):
        if self.h:
            self.ctx.lib.adaflo_b200_mf_destroy(self.h)
            self.h = None


class BlockVector:
    """LinearAlgebra::distributed::BlockVector<double> with blocks (u, p)."""

    def __init__(self, u, p):
        self.blocks = [u, p]

    def block(self, i):
        return self.blocks[i]

    @staticmethod
    def like(mf: MatrixFree):
        return BlockVector(mf.initialize_dof_vector(0), mf.initialize_dof_vector(1))


class NavierStokesMatrix:
    """adaflo::NavierStokesMatrix<3> (include/adaflo/navier_stokes_matrix.h:57-204)."""

    def __init__(self, parameters: FlowParams, dof_index_u, dof_index_p, quad_index_u, quad_index_p,
                 solution_old: BlockVector, solution_old_old: BlockVector):
        self.parameters = parameters
        self.dof_index_u, self.dof_index_p = dof_index_u, dof_index_p
        self.quad_index_u, self.quad_index_p = quad_index_u, quad_index_p
        self.solution_old, self.solution_old_old = solution_old, solution_old_old
        self.matrix_free = None
        self.h = None

    def initialize(self, matrix_free: MatrixFree, time_stepping: TimeWeights, pressure_average_fix: bool):
        self.clear()
        self.matrix_free = matrix_free
        self.ctx = matrix_free.ctx
        self.lib = self.ctx.lib
        h = C.c_void_p()
        check(self.lib.adaflo_b200_ns_create(self.ctx.h, matrix_free.h, C.byref(self.parameters), self.dof_index_u,
                                             self.dof_index_p, self.quad_index_u, self.quad_index_p,
                                             int(pressure_average_fix), C.byref(h)))
        self.h = h
        self.set_time_stepping(time_stepping)
        if self.solution_old is not None:
            check(self.lib.adaflo_b200_ns_set_old_solutions(self.h, _ptr(self.solution_old.block(0)),
                                                            _ptr(self.solution_old_old.block(0))))

    def set_time_stepping(self, time_stepping: TimeWeights):
        self.time_stepping = time_stepping
        check(self.lib.adaflo_b200_ns_set_time_weights(self.h, C.byref(time_stepping)))

    def set_parameters(self, parameters: FlowParams):
        self.parameters = parameters
        check(self.lib.adaflo_b200_ns_set_flow_params(self.h, C.byref(parameters)))

    def clear(self):
        if self.h:
            self.lib.adaflo_b200_ns_destroy(self.h)
        self.h = None
        self.matrix_free = None

    def _require(self):
        if not self.h:
            raise capi.AdafloB200Error("NavierStokesMatrix not initialized (ExcNotInitialized)")

    def vmult(self, dst: BlockVector, src: BlockVector):
        self._require()
        check(self.lib.adaflo_b200_ns_vmult(self.h, _ptr(dst.block(0)), _ptr(dst.block(1)), _ptr(src.block(0)), _ptr(src.block(1))))

    def Tvmult(self, dst: BlockVector, src: BlockVector):
        self._require()
        check(self.lib.adaflo_b200_ns_Tvmult(self.h, _ptr(dst.block(0)), _ptr(dst.block(1)), _ptr(src.block(0)), _ptr(src.block(1))))

    def residual(self, system_rhs: BlockVector, src: BlockVector, user_rhs: BlockVector):
        self._require()
        check(self.lib.adaflo_b200_ns_residual(self.h, _ptr(system_rhs.block(0)), _ptr(system_rhs.block(1)),
                                               _ptr(src.block(0)), _ptr(src.block(1)), _ptr(user_rhs.block(0)), _ptr(user_rhs.block(1))))

    def velocity_vmult(self, dst, src):
        self._require()
        check(self.lib.adaflo_b200_ns_velocity_vmult(self.h, _ptr(dst), _ptr(src)))

    def divergence_vmult_add(self, dst, src, weight_by_viscosity=False):
        self._require()
        check(self.lib.adaflo_b200_ns_divergence_vmult_add(self.h, _ptr(dst), _ptr(src), int(weight_by_viscosity)))

    def pressure_poisson_vmult(self, dst, src):
        self._require()
        check(self.lib.adaflo_b200_ns_pressure_poisson_vmult(self.h, _ptr(dst), _ptr(src)))

    def pressure_mass_vmult(self, dst, src):
        self._require()
        check(self.lib.adaflo_b200_ns_pressure_mass_vmult(self.h, _ptr(dst), _ptr(src)))

    def pressure_convdiff_vmult(self, dst, src):
        self._require()
        check(self.lib.adaflo_b200_ns_pressure_convdiff_vmult(self.h, _ptr(dst), _ptr(src)))

    def fix_linearization_point(self):
        self._require()
        check(self.lib.adaflo_b200_ns_fix_linearization_point(self.h))

    def compute_diagonal(self, diagonal: BlockVector):
        """exact diagonal of vmult (velocity block by unit-source passes, pressure block 0 / -1 on constrained rows)"""
        self._require()
        check(self.lib.adaflo_b200_ns_compute_diagonal(self.h, _ptr(diagonal.block(0)), _ptr(diagonal.block(1))))

    def apply_pressure_average_projection(self, vector):
        self._require()
        check(self.lib.adaflo_b200_ns_apply_pressure_average_projection(self.h, _ptr(vector)))

    def apply_pressure_shift(self, shift, pressure):
        self._require()
        check(self.lib.adaflo_b200_ns_apply_pressure_shift(self.h, float(shift), _ptr(pressure)))

    def _coef(self, which):
        self._require()
        p, n = C.c_void_p(), C.c_uint64()
        check(self.lib.adaflo_b200_ns_coefficient_ptr(self.h, which, C.byref(p), C.byref(n)))
        if not p.value:
            return None
        return device_view(p.value, n.value, self.ctx.device)

    # Per-quadrature-point coefficient arrays (navier_stokes_matrix.h:162-178,293-395).  On the device they are
    # decade blocked (csrc/common.cuh, coef_index / lin_index) -- a layout private to the library.  begin_*() return the
    # raw device views; set_coefficient / get_coefficient go through the C ABI's canonical copies
    # (adaflo_b200_ns_{set,get}_coefficient_canonical: the reference's [cell batch][q][row][lane] order).
    def begin_densities(self):
        return self._coef(capi.COEF_DENSITY)

    def begin_viscosities(self):
        return self._coef(capi.COEF_VISCOSITY)

    def begin_damping_coeff(self):
        return self._coef(capi.COEF_DAMPING)

    def begin_linearized_velocities(self):
        """rows 0..2 velocity, 3+3d+e = d u_d / d x_e"""
        return self._coef(capi.COEF_LINEARIZED_VELOCITY)

    def set_coefficient(self, which, canonical):
        """which: capi.COEF_*; canonical: [n_cells, 27] (or [n_cells, 12, 27]) device tensor, q = qx + 3 qy + 9 qz"""
        self._require()
        n_cells = self.matrix_free.mesh.n_cells
        if which == capi.COEF_LINEARIZED_VELOCITY:
            c = canonical.reshape(n_cells, 12, 27).permute(0, 2, 1).contiguous()   # ABI order: [cell][q][row]
        else:
            c = canonical.reshape(n_cells, 27).contiguous()
        check(self.lib.adaflo_b200_ns_set_coefficient_canonical(self.h, which, _ptr(c), 0, n_cells, 1, None))

    def get_coefficient(self, which):
        self._require()
        n_cells = self.matrix_free.mesh.n_cells
        n_rows = 12 if which == capi.COEF_LINEARIZED_VELOCITY else 1
        c = torch.zeros(n_cells, 27, n_rows, dtype=torch.float64, device=self.ctx.device)
        check(self.lib.adaflo_b200_ns_get_coefficient_canonical(self.h, which, _ptr(c), 0, n_cells, 1, None))
        return c.permute(0, 2, 1).contiguous() if n_rows == 12 else c.reshape(n_cells, 27)

    def use_variable_coefficients(self):
        return self.parameters.density_diff != 0 or self.parameters.viscosity_diff != 0

    def get_matvec_statistics(self):
        self._require()
        n, s = C.c_uint64(), C.c_double()
        check(self.lib.adaflo_b200_ns_get_matvec_statistics(self.h, C.byref(n), C.byref(s)))
        return s.value, n.value

    def set_kernel_timing(self, enable=True):
        self._require()
        check(self.lib.adaflo_b200_ns_set_kernel_timing(self.h, int(enable)))

    def get_kernel_timing(self):
        """(launches, seconds) of the vmult cell kernel measured with CUDA events; resets."""
        self._require()
        n, s = C.c_uint64(), C.c_double()
        check(self.lib.adaflo_b200_ns_get_kernel_timing(self.h, C.byref(n), C.byref(s)))
        return n.value, s.value

    def vmult_host(self, dst_u, dst_p, src_u, src_p):
        """End-to-end call with HOST (pinned) torch tensors or numpy arrays."""
        self._require()

        def hp(t):
            return C.c_void_p(t.data_ptr()) if isinstance(t, torch.Tensor) else _np(t)
        check(self.lib.adaflo_b200_ns_vmult_host(self.h, hp(dst_u), hp(dst_p), hp(src_u), hp(src_p)))

    # host-pointer entries of the sub-block operators (the Epetra-side wrappers of navier_stokes_preconditioner.cc)
    def velocity_vmult_host(self, dst_u, src_u):
        self._require()
        check(self.lib.adaflo_b200_ns_velocity_vmult_host(self.h, _host_ptr(dst_u), _host_ptr(src_u)))

    def pressure_poisson_vmult_host(self, dst_p, src_p):
        self._require()
        check(self.lib.adaflo_b200_ns_pressure_poisson_vmult_host(self.h, _host_ptr(dst_p), _host_ptr(src_p)))

    def pressure_mass_vmult_host(self, dst_p, src_p):
        self._require()
        check(self.lib.adaflo_b200_ns_pressure_mass_vmult_host(self.h, _host_ptr(dst_p), _host_ptr(src_p)))

    def pressure_convdiff_vmult_host(self, dst_p, src_p):
        self._require()
        check(self.lib.adaflo_b200_ns_pressure_convdiff_vmult_host(self.h, _host_ptr(dst_p), _host_ptr(src_p)))

    def divergence_vmult_add_host(self, dst_p, src_u, weight_by_viscosity=False):
        self._require()
        check(self.lib.adaflo_b200_ns_divergence_vmult_add_host(self.h, _host_ptr(dst_p), _host_ptr(src_u), int(weight_by_viscosity)))


def _host_ptr(t):
    return C.c_void_p(t.data_ptr()) if isinstance(t, torch.Tensor) else _np(t)


class NavierStokesMatrix2D:
    """NavierStokesMatrix<2> for Q2/Q1 on Cartesian box meshes (adaflo_b200.mesh.BoxMesh2D).

    The CUDA kernels are three-dimensional.  A 2-D operator is applied exactly as A_2d = R A_3d E on the
    one-cell-thick extrusion of the mesh: E copies a 2-D field to the node planes (third velocity component
    constrained to zero), R sums the planes; the z shape functions sum to one, so every bilinear form of
    source/navier_stokes_matrix.cc reduces to its 2-D counterpart for z-constant fields (verified against an
    independent dense 2-D oracle, tests/test_2d_cpu.py / tests/test_2d_gpu.py).  Vectors are 2-D: velocity entry
    (node, c) = 2 node + c.  Meant for the small 2-D configurations; it costs 3-D work per 2-D cell."""

    def __init__(self, parameters, ctx, mesh2d, time_stepping, pressure_average_fix=True):
        self.mesh2d, self.ctx = mesh2d, ctx
        self.mf = MatrixFree(ctx, mesh2d.mesh3d, "cartesian")
        self.old, self.old_old = BlockVector.like(self.mf), BlockVector.like(self.mf)
        self.A = NavierStokesMatrix(parameters, 0, 1, 0, 1, self.old, self.old_old)
        self.A.initialize(self.mf, time_stepping, pressure_average_fix)
        dev = ctx.device
        self.ext_u = torch.from_numpy(mesh2d.ext_u).to(dev)
        self.ext_p = torch.from_numpy(mesh2d.ext_p).to(dev)
        self.constrained_u = torch.from_numpy(np.nonzero(mesh2d.u_dirichlet.reshape(-1))[0]).to(dev)
        self._src, self._dst, self._aux = BlockVector.like(self.mf), BlockVector.like(self.mf), BlockVector.like(self.mf)

    # ---- extension / restriction on the device
    def _extend(self, dst3: BlockVector, u2, p2):
        dst3.block(0).zero_()
        dst3.block(0)[self.ext_u] = u2.unsqueeze(1)
        if p2 is not None:
            dst3.block(1).zero_()
            dst3.block(1)[self.ext_p] = p2.unsqueeze(1)

    def _restrict(self, src3: BlockVector, u2, p2):
        torch.sum(src3.block(0)[self.ext_u], dim=1, out=u2)
        if p2 is not None:
            torch.sum(src3.block(1)[self.ext_p], dim=1, out=p2)

    def initialize_dof_vectors(self):
        dev = self.ctx.device
        return (torch.zeros(self.mesh2d.n_u_nodes * 2, dtype=torch.float64, device=dev),
                torch.zeros(self.mesh2d.n_p_nodes, dtype=torch.float64, device=dev))

    def set_old_solutions(self, old_u2, old_old_u2):
        self._extend(self._aux, old_u2, None)
        self.old.block(0).copy_(self._aux.block(0))
        self._extend(self._aux, old_old_u2, None)
        self.old_old.block(0).copy_(self._aux.block(0))

    def _apply(self, op, dst2, src2):
        self._extend(self._src, src2[0], src2[1])
        op(self._dst, self._src)
        self._restrict(self._dst, dst2[0], dst2[1])
        dst2[0][self.constrained_u] = src2[0][self.constrained_u]   # one copy of a constrained row, not one per plane

    def vmult(self, dst2, src2):
        self._apply(self.A.vmult, dst2, src2)

    def Tvmult(self, dst2, src2):
        self._apply(self.A.Tvmult, dst2, src2)

    def velocity_vmult(self, dst_u2, src_u2):
        self._extend(self._src, src_u2, None)
        self.A.velocity_vmult(self._dst.block(0), self._src.block(0))
        self._restrict(self._dst, dst_u2, None)
        dst_u2[self.constrained_u] = src_u2[self.constrained_u]

    def residual(self, rhs2, src2, user2):
        """rhs = user - (cell integrals + rhs preload), navier_stokes_matrix.cc:266-293; also stores the linearisation"""
        self._extend(self._src, src2[0], src2[1])
        self._dst.block(0).zero_()
        self._dst.block(1).zero_()
        self._aux.block(0).zero_()
        self._aux.block(1).zero_()
        self.A.residual(self._dst, self._src, self._aux)           # = -(cell integrals) on the extrusion
        pre_u, pre_p = rhs2[0].clone(), rhs2[1].clone()
        self._restrict(self._dst, rhs2[0], rhs2[1])
        rhs2[0].add_(user2[0] - pre_u)
        rhs2[1].add_(user2[1] - pre_p)

    def clear(self):
        self.A.clear()
        self.mf.close()


class DiagonalPreconditioner:
    """adaflo::DiagonalPreconditioner<double> (source/diagonal_preconditioner.cc:38-124)."""

    def __init__(self, ctx: Context, diagonal_vector=None):
        self.ctx = ctx
        self.inverse_diagonal_vector = None
        if diagonal_vector is not None:
            self.reinit(diagonal_vector)

    def reinit(self, diagonal_vector, n_owned=None):
        self.n = n_owned if n_owned is not None else diagonal_vector.numel()
        self.inverse_diagonal_vector = torch.zeros_like(diagonal_vector)
        check(self.ctx.lib.adaflo_b200_jacobi_reinit(self.ctx.h, _ptr(self.inverse_diagonal_vector), _ptr(diagonal_vector), self.n))

    def vmult(self, dst, src, n_blocks=1):
        check(self.ctx.lib.adaflo_b200_jacobi_vmult(self.ctx.h, _ptr(dst), _ptr(src), _ptr(self.inverse_diagonal_vector), self.n, n_blocks))


class LevelSetOperators:
    """The matrix-free operators of adaflo's LevelSetOKZSolver{AdvanceConcentration, ComputeNormal,
    ComputeCurvature, Reinitialization} (dof handler 2 = level set, quadrature 2 =
    QIterated(QGauss(2), n_sub): source/two_phase_base.cc:242-271).  Normal fields are lists of three
    device vectors (the blocks of the reference's BlockVector)."""

    def __init__(self, matrix_free: MatrixFree, parameters: capi.LSParams, cell_diameters, dof_index_ls=2, dof_index_vel=0, quad_index=2):
        self.matrix_free = matrix_free
        self.ctx = matrix_free.ctx
        self.lib = self.ctx.lib
        self.parameters = parameters
        cd = np.ascontiguousarray(cell_diameters, dtype=np.float64)
        h = C.c_void_p()
        check(self.lib.adaflo_b200_ls_create(self.ctx.h, matrix_free.h, C.byref(parameters), dof_index_ls, dof_index_vel, quad_index,
                                             _np(cd), C.byref(h)))
        self.h = h
        self.dof_index_ls = dof_index_ls

    def close(self):
        if self.h:
            self.lib.adaflo_b200_ls_destroy(self.h)
            self.h = None

    @staticmethod
    def _blocks(vs):
        return (C.c_void_p * 3)(*[v.data_ptr() for v in vs])

    def set_time(self, weight, weight_old, weight_old_old, step_size):
        check(self.lib.adaflo_b200_ls_set_time(self.h, weight, weight_old, weight_old_old, step_size))

    def set_diagonal(self, diag):
        self._diag = diag
        check(self.lib.adaflo_b200_ls_set_diagonal(self.h, _ptr(diag)))

    def initialize_mass_matrix_diagonal(self):
        d = self.matrix_free.initialize_dof_vector(self.dof_index_ls)
        check(self.lib.adaflo_b200_ls_mass_diagonal(self.h, _ptr(d)))
        return d

    def advance_concentration_vmult(self, dst, src):
        check(self.lib.adaflo_b200_ls_advance_concentration_vmult(self.h, _ptr(dst), _ptr(src)))

    def advance_concentration_rhs(self, rhs, vel, c, c_old, c_old_old):
        check(self.lib.adaflo_b200_ls_advance_concentration_rhs(self.h, _ptr(rhs), _ptr(vel), _ptr(c), _ptr(c_old), _ptr(c_old_old)))

    def compute_normal_vmult(self, dst, src):
        check(self.lib.adaflo_b200_ls_compute_normal_vmult(self.h, self._blocks(dst), self._blocks(src)))

    def compute_normal_rhs(self, rhs, level_set):
        check(self.lib.adaflo_b200_ls_compute_normal_rhs(self.h, self._blocks(rhs), _ptr(level_set)))

    def compute_curvature_vmult(self, dst, src, apply_diffusion=True):
        check(self.lib.adaflo_b200_ls_compute_curvature_vmult(self.h, _ptr(dst), _ptr(src), int(apply_diffusion)))

    def compute_curvature_rhs(self, rhs, normal):
        check(self.lib.adaflo_b200_ls_compute_curvature_rhs(self.h, _ptr(rhs), self._blocks(normal)))

    def curvature_correction(self, curvature, level_set):
        check(self.lib.adaflo_b200_ls_curvature_correction(self.h, _ptr(curvature), _ptr(level_set)))

    def reinitialization_vmult(self, dst, src, diffuse_only=False):
        check(self.lib.adaflo_b200_ls_reinitialization_vmult(self.h, _ptr(dst), _ptr(src), int(diffuse_only)))

    def reinitialization_rhs(self, rhs, solution, normal, diffuse_only=False, first_reinit_step=True):
        blocks = self._blocks(normal) if normal is not None else None
        check(self.lib.adaflo_b200_ls_reinitialization_rhs(self.h, _ptr(rhs), _ptr(solution), blocks, int(diffuse_only), int(first_reinit_step)))

    def compute_heaviside(self, heaviside, level_set):
        check(self.lib.adaflo_b200_ls_compute_heaviside(self.h, _ptr(heaviside), _ptr(level_set)))

    # ---- solver-level routines (the bodies of compute_normal / compute_curvature / reinitialize / advance_concentration
    # of the reference's level-set classes, matrix-free; csrc/ls_api.cu)
    def projection_vmult(self, dst, src):
        """mass + damping_scale_factor delta^2 Laplace on one field: the reference's assembled projection_matrix"""
        check(self.lib.adaflo_b200_ls_projection_vmult(self.h, _ptr(dst), _ptr(src)))

    def compute_normal(self, normal, level_set, fast_computation=False):
        """returns the CG iteration count"""
        its = C.c_int()
        check(self.lib.adaflo_b200_ls_compute_normal(self.h, self._blocks(normal), _ptr(level_set), int(fast_computation), C.byref(its)))
        return its.value

    def compute_curvature(self, curvature, level_set, normal, curvature_correction=True, projection_operator=1):
        its = C.c_int()
        check(self.lib.adaflo_b200_ls_compute_curvature(self.h, _ptr(curvature), _ptr(level_set), self._blocks(normal), int(curvature_correction),
                                                        int(projection_operator), C.byref(its)))
        return its.value

    def reinitialize(self, solution, normal, dt, stab_steps, diff_steps=0, last_concentration_range=None):
        """returns the CG iteration counts of the pseudo time steps taken (the numbers of the log's 'reinitialize (a + b)')"""
        n_max = stab_steps + diff_steps + 3
        its, done = (C.c_int * n_max)(), C.c_int()
        rng = None if last_concentration_range is None else (C.c_double * 2)(*[float(v) for v in last_concentration_range])
        check(self.lib.adaflo_b200_ls_reinitialize(self.h, _ptr(solution), self._blocks(normal), float(dt), int(stab_steps), int(diff_steps), rng,
                                                   its, C.byref(done)))
        return [its[i] for i in range(done.value)]

    def advance_concentration(self, solution, solution_old, solution_old_old, vel_solution, tol_nl_iteration=1e-8):
        """returns (initial residual, iterations): the numbers of the log line 'advect [r0/its]'"""
        r0, its = C.c_double(), C.c_int()
        check(self.lib.adaflo_b200_ls_advance_concentration(self.h, _ptr(solution), _ptr(solution_old), _ptr(solution_old_old), _ptr(vel_solution),
                                                            float(tol_nl_iteration), C.byref(r0), C.byref(its)))
        return r0.value, its.value

    def compute_force(self, navier_stokes_matrix, dst_u, heaviside, curvature, surface_tension, gravity, interpolation=None):
        ip = None if interpolation is None else np.ascontiguousarray(interpolation, dtype=np.float64)
        check(self.lib.adaflo_b200_ls_compute_force(self.h, navier_stokes_matrix.h, _ptr(dst_u), _ptr(heaviside), _ptr(curvature),
                                                    float(surface_tension), float(gravity), _np(ip)))

    def pointdata(self, which):
        """0: evaluated_convection, 1: evaluated_normal; device view [n_cells, 3, 8, n_sub^3] (component, point of the
        sub-cell, sub-cell: the layout of the cell kernel in csrc/ls_ops.cu)"""
        p, n = C.c_void_p(), C.c_uint64()
        check(self.lib.adaflo_b200_ls_pointdata_ptr(self.h, which, C.byref(p), C.byref(n)))
        s = self.parameters.concentration_subdivisions
        return device_view(p.value, n.value, self.ctx.device).view(-1, 3, 8, s ** 3)

    def pointdata_canonical(self, which):
        """[n_cells, NQ, 3] with q = qx + nq qy + nq^2 qz, the order the oracle uses"""
        s = self.parameters.concentration_subdivisions
        t = self.pointdata(which)                               # cell, comp, (pc,pb,pa), (sz,sy,sx)
        n = t.shape[0]
        t = t.view(n, 3, 2, 2, 2, s, s, s).permute(0, 5, 2, 6, 3, 7, 4, 1)   # cell, sz,pc, sy,pb, sx,pa, comp
        return t.reshape(n, (2 * s) ** 3, 3)


class LevelSetOperators2D:
    """The level-set operators of LevelSetOperators for dim = 2 (FE_Q_iso_Q1(n_sub) on BoxMesh2D), applied through the
    3-D kernels on the one-cell-thick extrusion like NavierStokesMatrix2D: fields are copied to the n_sub + 1 node
    planes (third normal / velocity component zero), results are summed over the planes.  The hat functions sum to
    one in z, so all weak forms reduce exactly; two fix-ups: the mass diagonal is 3/2 of the plane sum (sum_a int
    N_a^2 dz = 2/3), and constrained rows get diag * src once.  Point-wise operations (Heaviside, curvature
    correction) are read from one plane.  Checked against an independent dense 2-D restatement (tests/test_2d_gpu.py)."""

    def __init__(self, ctx, mesh2d, parameters: capi.LSParams, cell_diameters=None, matrix_free=None):
        """matrix_free: share the MatrixFree of a NavierStokesMatrix2D built after mesh2d.build_ls_space (compute_force)"""
        assert parameters.dim == 2, "LSParams.dim must be 2 (pseudo time step and curvature correction depend on it)"
        self.ctx, self.mesh2d = ctx, mesh2d
        self._own_mf = matrix_free is None
        self.mf = MatrixFree(ctx, mesh2d.mesh3d, "cartesian") if matrix_free is None else matrix_free
        cd = mesh2d.cell_diameters() if cell_diameters is None else cell_diameters
        self.L = LevelSetOperators(self.mf, parameters, cd)
        dev = ctx.device
        self.ext = torch.from_numpy(mesh2d.ext_ls).to(dev)
        self.ext_u = torch.from_numpy(mesh2d.ext_u).to(dev)
        self.con = torch.from_numpy(np.nonzero(mesh2d.ls_dirichlet)[0]).to(dev)
        self.n, self.n3 = mesh2d.n_ls_nodes, self.mf.initialize_dof_vector(2).numel()
        d3 = self.L.initialize_mass_matrix_diagonal()
        self.L.set_diagonal(d3)
        self.diag = 1.5 * self._R(d3)

    def new_vector(self):
        return torch.zeros(self.n, dtype=torch.float64, device=self.ctx.device)

    def _E(self, v2):
        v3 = torch.zeros(self.n3, dtype=torch.float64, device=self.ctx.device)
        if v2 is not None:
            v3[self.ext] = v2.unsqueeze(1)
        return v3

    def _R(self, v3):
        return v3[self.ext].sum(dim=1)

    def _rows(self, dst2, src2):
        dst2[self.con] = self.diag[self.con] * src2[self.con]

    def set_time(self, weight, weight_old, weight_old_old, step_size):
        self.L.set_time(weight, weight_old, weight_old_old, step_size)

    def initialize_mass_matrix_diagonal(self):
        return self.diag.clone()

    def advance_concentration_vmult(self, dst, src):
        d3 = self._E(None)
        self.L.advance_concentration_vmult(d3, self._E(src))
        dst.copy_(self._R(d3))
        self._rows(dst, src)

    def advance_concentration_rhs(self, rhs, vel, c, c_old, c_old_old):
        u3 = torch.zeros(self.mf.initialize_dof_vector(0).numel(), dtype=torch.float64, device=self.ctx.device)
        u3[self.ext_u] = vel.unsqueeze(1)
        r3 = self._E(None)
        self.L.advance_concentration_rhs(r3, u3, self._E(c), self._E(c_old), self._E(c_old_old))
        rhs.add_(self._R(r3))

    def compute_normal_vmult(self, dst, src):
        d3 = [self._E(None) for _ in range(3)]
        self.L.compute_normal_vmult(d3, [self._E(src[0]), self._E(src[1]), self._E(None)])
        for d in range(2):
            dst[d].copy_(self._R(d3[d]))
            self._rows(dst[d], src[d])

    def compute_normal_rhs(self, rhs, level_set):
        r3 = [self._E(None) for _ in range(3)]
        self.L.compute_normal_rhs(r3, self._E(level_set))
        for d in range(2):
            rhs[d].add_(self._R(r3[d]))

    def compute_curvature_vmult(self, dst, src, apply_diffusion=True):
        d3 = self._E(None)
        self.L.compute_curvature_vmult(d3, self._E(src), apply_diffusion)
        dst.copy_(self._R(d3))
        self._rows(dst, src)

    def compute_curvature_rhs(self, rhs, normal):
        r3 = self._E(None)
        self.L.compute_curvature_rhs(r3, [self._E(normal[0]), self._E(normal[1]), self._E(None)])
        rhs.add_(self._R(r3))

    def curvature_correction(self, curvature, level_set):
        k3 = self._E(curvature)
        self.L.curvature_correction(k3, self._E(level_set))
        curvature.copy_(k3[self.ext[:, 0]])

    def reinitialization_vmult(self, dst, src, diffuse_only=False):
        d3 = self._E(None)
        self.L.reinitialization_vmult(d3, self._E(src), diffuse_only)
        dst.copy_(self._R(d3))
        self._rows(dst, src)

    def reinitialization_rhs(self, rhs, solution, normal, diffuse_only=False, first_reinit_step=True):
        r3 = self._E(None)
        n3 = None if normal is None else [self._E(normal[0]), self._E(normal[1]), self._E(None)]
        self.L.reinitialization_rhs(r3, self._E(solution), n3, diffuse_only, first_reinit_step)
        rhs.add_(self._R(r3))

    def compute_heaviside(self, heaviside, level_set):
        h3 = self._E(None)
        self.L.compute_heaviside(h3, self._E(level_set))
        heaviside.copy_(h3[self.ext[:, 0]])

    def compute_force(self, ns2d, dst_u2, heaviside, curvature, surface_tension, gravity, interpolate=False):
        """local_compute_force for dim = 2 (gravity on the y component); ns2d: NavierStokesMatrix2D sharing self.mf.
        rho / mu at the quadrature points are written into ns2d's (extruded) coefficient arrays.
        interpolate: take the gradient of the Heaviside function from its nodal interpolant on the pressure space
        (interpolate_grad_onto_pressure, level_set_okz.cc:367-378); the extruded field is constant in z, so the
        corner-picking matrix of the 3-D element gives the 2-D interpolant on every node plane."""
        f3 = torch.zeros(self.mf.initialize_dof_vector(0).numel(), dtype=torch.float64, device=self.ctx.device)
        interp = None
        if interpolate:
            s = self.L.parameters.concentration_subdivisions
            n1 = s + 1
            interp = np.zeros((8, n1 ** 3))
            for i in range(8):
                interp[i, s * (i & 1) + n1 * (s * ((i >> 1) & 1) + n1 * s * (i >> 2))] = 1.0
        self.L.compute_force(ns2d.A, f3, self._E(heaviside), self._E(curvature), surface_tension, gravity, interp)
        dst_u2.add_(f3[self.ext_u].sum(dim=1))

    def close(self):
        self.L.close()
        if self._own_mf:
            self.mf.close()


class MatrixFreeGeneric:
    """Device mirror of MatrixFree for the degree / dimension dispatch: Taylor-Hood Q_k / Q_(k-1) on an
    adaflo_b200.mesh.TensorMesh (dim 2 or 3), dof handlers {0: u, 1: p}, quadratures {0: QGauss(k+1), 1: QGauss(k)}."""

    def __init__(self, ctx: Context, mesh):
        self.ctx, self.mesh = ctx, mesh
        k, dim = mesh.degree, mesh.dim
        cu = np.ascontiguousarray(mesh.encoded_cell_u())
        cp = np.ascontiguousarray(mesh.cell_p.astype(np.uint32))
        con_u = np.ascontiguousarray(mesh.constrained_u())
        con_p = np.ascontiguousarray(mesh.constrained_p()) if hasattr(mesh, "constrained_p") else np.zeros(0, dtype=np.uint32)
        # hanging nodes: general constraint rows (adaflo_b200_space_desc::hanging_*), e.g. mesh.RefinedTensorMesh2D
        rows_u = tuple(np.ascontiguousarray(a) for a in mesh.encoded_rows_u()) if hasattr(mesh, "encoded_rows_u") else None
        rows_p = tuple(np.ascontiguousarray(a) for a in mesh.p_rows) if hasattr(mesh, "p_rows") else None
        spaces = (capi.SpaceDesc * 2)()
        for i, (cells, n_nodes, n_comp, deg, con, rows) in enumerate([(cu, mesh.n_u_nodes, dim, k, con_u, rows_u),
                                                                      (cp, mesh.n_p_nodes, 1, k - 1, con_p, rows_p)]):
            d = spaces[i]
            d.fe_kind, d.degree, d.n_components = capi.FE_Q, deg, n_comp
            d.n_owned_nodes, d.n_ghost_nodes, d.nodes_per_cell = n_nodes, 0, cells.shape[1]
            d.cell_nodes = cells.ctypes.data
            d.hanging_masks = None
            d.n_constrained = len(con)
            d.constrained_dofs = con.ctypes.data if len(con) else None
            d.n_exchange, d.exchange = 0, None
            if rows is not None and len(rows[0]) > 1:
                d.n_hanging_nodes = len(rows[0]) - 1
                d.hanging_row_ptr, d.hanging_cols, d.hanging_weights = rows[0].ctypes.data, rows[1].ctypes.data, rows[2].ctypes.data
        table, row = mesh.cartesian_geometry()
        table, row = np.ascontiguousarray(table), np.ascontiguousarray(row)
        quads = (capi.QuadDesc * 2)()
        for i, nq1 in enumerate((k + 1, k)):
            q = quads[i]
            q.n_q_1d, q.geometry_type, q.n_table_rows = nq1, capi.GEOM_CARTESIAN, table.shape[0]
            q.table, q.cell_row = table.ctypes.data, row.ctypes.data
        desc = capi.MFDesc(dim, mesh.n_cells, 2, spaces, 2, quads)
        h = C.c_void_p()
        check(ctx.lib.adaflo_b200_mf_create(ctx.h, C.byref(desc), C.byref(h)))
        self.h = h

    def initialize_dof_vector(self, space):
        n_owned, n_local = C.c_uint64(), C.c_uint64()
        check(self.ctx.lib.adaflo_b200_mf_vector_size(self.h, space, C.byref(n_owned), C.byref(n_local)))
        return torch.zeros(n_local.value, dtype=torch.float64, device=self.ctx.device)

    def n_owned(self, space):
        n_owned = C.c_uint64()
        check(self.ctx.lib.adaflo_b200_mf_vector_size(self.h, space, C.byref(n_owned), None))
        return n_owned.value

    def distribute_constraints(self, space, vec):
        """AffineConstraints::distribute for the hanging-node rows (navier_stokes.cc:643-644, 1252)"""
        check(self.ctx.lib.adaflo_b200_mf_distribute_constraints(self.h, space, _ptr(vec)))

    def close(self):
        if self.h:
            self.ctx.lib.adaflo_b200_mf_destroy(self.h)
            self.h = C.c_void_p()
