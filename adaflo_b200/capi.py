"""ctypes binding of the C ABI (include/adaflo_b200.h).

The shared library is built in-tree by ``__graft_entry__.build()`` /
``make -C adaflo_b200/csrc``.  There is no fallback: if the library is missing
or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ADAFLO_B200_LIB") or os.path.join(_HERE, "lib", "libadaflo_b200.so")  # override: kernel A/B experiments


class AdafloB200Error(RuntimeError):
    """a non-zero return code of the C ABI; .code is the ADAFLO_B200_ERR_* value"""

    def __init__(self, message, code=None):
        super().__init__(message)
        self.code = code


class Exchange(C.Structure):
    _fields_ = [("peer", C.c_int32), ("n_send", C.c_uint32), ("send", C.c_void_p),
                ("recv_first", C.c_uint32), ("n_recv", C.c_uint32)]


class SpaceDesc(C.Structure):
    _fields_ = [("fe_kind", C.c_int32), ("degree", C.c_int32), ("n_components", C.c_int32),
                ("n_owned_nodes", C.c_uint32), ("n_ghost_nodes", C.c_uint32), ("nodes_per_cell", C.c_uint32),
                ("cell_nodes", C.c_void_p), ("hanging_masks", C.c_void_p),
                ("n_constrained", C.c_uint32), ("constrained_dofs", C.c_void_p),
                ("n_exchange", C.c_int32), ("exchange", C.POINTER(Exchange)),
                ("n_hanging_nodes", C.c_uint32), ("hanging_row_ptr", C.c_void_p), ("hanging_cols", C.c_void_p),
                ("hanging_weights", C.c_void_p)]


class QuadDesc(C.Structure):
    _fields_ = [("n_q_1d", C.c_int32), ("geometry_type", C.c_int32), ("n_table_rows", C.c_uint32),
                ("table", C.c_void_p), ("cell_row", C.c_void_p), ("inv_jacobian", C.c_void_p), ("jxw", C.c_void_p)]


class MFDesc(C.Structure):
    _fields_ = [("dim", C.c_int32), ("n_cells", C.c_uint32), ("n_spaces", C.c_int32),
                ("spaces", C.POINTER(SpaceDesc)), ("n_quads", C.c_int32), ("quads", C.POINTER(QuadDesc))]


class FlowParams(C.Structure):
    """adaflo_b200_flow_params -- mirrors adaflo::FlowParameters (include/adaflo/parameters.h:54-86)."""
    _fields_ = [("physical_type", C.c_int32), ("linearization", C.c_int32), ("augmented_taylor_hood", C.c_int32),
                ("reserved", C.c_int32), ("density", C.c_double), ("viscosity", C.c_double),
                ("density_diff", C.c_double), ("viscosity_diff", C.c_double),
                ("damping", C.c_double), ("tau_grad_div", C.c_double),
                ("beta_convective_term_momentum_balance", C.c_double)]


class LSParams(C.Structure):
    """adaflo_b200_ls_params"""
    _fields_ = [("dim", C.c_int32), ("concentration_subdivisions", C.c_int32), ("convection_stabilization", C.c_int32),
                ("reserved", C.c_int32), ("epsilon", C.c_double), ("epsilon_used", C.c_double),
                ("damping_scale_factor", C.c_double), ("minimal_edge_length", C.c_double)]


class TimeWeights(C.Structure):
    """adaflo_b200_time_weights -- the six scalars of adaflo::TimeStepping the kernels read."""
    _fields_ = [("weight", C.c_double), ("weight_old", C.c_double), ("weight_old_old", C.c_double),
                ("tau1", C.c_double), ("extrapolate_old", C.c_double), ("extrapolate_old_old", C.c_double)]


class SolverControlABI(C.Structure):
    """adaflo_b200_solver_control"""
    _fields_ = [("max_iter", C.c_int), ("abs_tol", C.c_double), ("rel_tol", C.c_double), ("last_step", C.c_int),
                ("initial_value", C.c_double), ("last_value", C.c_double), ("history", C.POINTER(C.c_double)), ("history_capacity", C.c_int),
                ("history_size", C.c_int)]


class OperatorRef(C.Structure):
    """adaflo_b200_operator_ref: a library operator as the user data of adaflo_b200_apply_operator"""
    _fields_ = [("kind", C.c_int), ("handle", C.c_void_p), ("flag", C.c_int), ("inv_diag", C.c_void_p), ("n", C.c_uint64),
                ("n_blocks", C.c_int)]


# int (*adaflo_b200_apply_fn)(void *user, double *const *dst, const double *const *src)
APPLY_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p))
OP_NS_VMULT, OP_NS_VELOCITY, OP_NS_PRESSURE_POISSON, OP_NS_PRESSURE_MASS, OP_LS_ADVECT, OP_LS_NORMAL, OP_LS_CURVATURE, OP_LS_REINIT, OP_JACOBI, \
    OP_LS_NORMAL_SCALAR = range(10)
ERR_ARG, ERR_CUDA, ERR_UNSUPPORTED, ERR_STATE, ERR_COMM, ERR_NO_CONVERGENCE = 1, 2, 3, 4, 5, 6

INCOMPRESSIBLE, INCOMPRESSIBLE_STATIONARY, STOKES = 0, 1, 2
NEWTON, PICARD, SEMI_IMPLICIT, EXPLICIT, PROJECTION = 0, 1, 2, 3, 4
FE_Q, FE_Q_DG0, FE_Q_ISO_Q1 = 0, 1, 2
GEOM_CARTESIAN, GEOM_AFFINE, GEOM_GENERAL = 0, 1, 2
COEF_DENSITY, COEF_VISCOSITY, COEF_DAMPING, COEF_LINEARIZED_VELOCITY = 0, 1, 2, 3

# every symbol include/adaflo_b200.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "adaflo_b200_last_error", "adaflo_b200_version",
    "adaflo_b200_ctx_create", "adaflo_b200_ctx_destroy", "adaflo_b200_ctx_set_stream", "adaflo_b200_ctx_synchronize",
    "adaflo_b200_comm_unique_id", "adaflo_b200_ctx_init_comm",
    "adaflo_b200_mf_create", "adaflo_b200_mf_destroy", "adaflo_b200_mf_vector_size", "adaflo_b200_mf_space_info", "adaflo_b200_mf_n_cells",
    "adaflo_b200_mf_set_dof_numbering", "adaflo_b200_mf_upload_vector", "adaflo_b200_mf_download_vector",
    "adaflo_b200_mf_update_ghost_values", "adaflo_b200_mf_compress_add", "adaflo_b200_mf_distribute_constraints",
    "adaflo_b200_ns_create", "adaflo_b200_ns_destroy", "adaflo_b200_ns_set_flow_params",
    "adaflo_b200_ns_set_time_weights", "adaflo_b200_ns_set_old_solutions",
    "adaflo_b200_ns_vmult", "adaflo_b200_ns_Tvmult", "adaflo_b200_ns_residual", "adaflo_b200_ns_velocity_vmult",
    "adaflo_b200_ns_divergence_vmult_add", "adaflo_b200_ns_pressure_poisson_vmult",
    "adaflo_b200_ns_pressure_mass_vmult", "adaflo_b200_ns_pressure_convdiff_vmult",
    "adaflo_b200_ns_compute_diagonal", "adaflo_b200_ns_fix_linearization_point",
    "adaflo_b200_ns_apply_pressure_average_projection", "adaflo_b200_ns_apply_pressure_shift",
    "adaflo_b200_ns_coefficient_ptr", "adaflo_b200_ns_set_coefficient_canonical", "adaflo_b200_ns_get_coefficient_canonical",
    "adaflo_b200_ns_nullspace", "adaflo_b200_ns_get_matvec_statistics",
    "adaflo_b200_ns_vmult_host", "adaflo_b200_ns_velocity_vmult_host", "adaflo_b200_ns_pressure_poisson_vmult_host",
    "adaflo_b200_ns_pressure_mass_vmult_host", "adaflo_b200_ns_pressure_convdiff_vmult_host", "adaflo_b200_ns_divergence_vmult_add_host",
    "adaflo_b200_ns_set_kernel_timing", "adaflo_b200_ns_get_kernel_timing",
    "adaflo_b200_kernel_launch_count",
    "adaflo_b200_ls_create", "adaflo_b200_ls_destroy", "adaflo_b200_ls_set_time", "adaflo_b200_ls_set_diagonal",
    "adaflo_b200_ls_mass_diagonal", "adaflo_b200_ls_advance_concentration_vmult", "adaflo_b200_ls_advance_concentration_rhs",
    "adaflo_b200_ls_compute_normal_vmult", "adaflo_b200_ls_compute_normal_rhs", "adaflo_b200_ls_compute_curvature_vmult",
    "adaflo_b200_ls_compute_curvature_rhs", "adaflo_b200_ls_curvature_correction", "adaflo_b200_ls_reinitialization_vmult",
    "adaflo_b200_ls_reinitialization_rhs", "adaflo_b200_ls_pointdata_ptr", "adaflo_b200_ls_compute_heaviside",
    "adaflo_b200_ls_compute_force",
    "adaflo_b200_vec_alloc", "adaflo_b200_vec_free", "adaflo_b200_vec_upload", "adaflo_b200_vec_download",
    "adaflo_b200_vec_set", "adaflo_b200_vec_axpy", "adaflo_b200_vec_sadd", "adaflo_b200_vec_scale",
    "adaflo_b200_vec_dot", "adaflo_b200_vec_multi_dot", "adaflo_b200_vec_multi_axpy", "adaflo_b200_vec_add_and_dot",
    "adaflo_b200_jacobi_reinit", "adaflo_b200_jacobi_vmult",
    "adaflo_b200_solve_cg", "adaflo_b200_solve_bicgstab", "adaflo_b200_solve_gmres", "adaflo_b200_apply_operator",
    "adaflo_b200_ns_preconditioner_create", "adaflo_b200_ns_preconditioner_destroy", "adaflo_b200_ns_preconditioner_vmult",
    "adaflo_b200_ns_preconditioner_apply", "adaflo_b200_ns_preconditioner_statistics",
    "adaflo_b200_ls_projection_vmult", "adaflo_b200_ls_compute_normal", "adaflo_b200_ls_compute_curvature", "adaflo_b200_ls_reinitialize",
    "adaflo_b200_ls_advance_concentration",
]

_lib = None


def load():
    """Load libadaflo_b200.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AdafloB200Error(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C adaflo_b200/csrc`; there is no CPU fallback")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    lib.adaflo_b200_last_error.restype = C.c_char_p
    lib.adaflo_b200_version.restype = C.c_char_p
    vp, dbl, i32, u64 = C.c_void_p, C.c_double, C.c_int, C.c_uint64
    sig = {
        "adaflo_b200_ctx_create": [C.POINTER(vp), i32, vp],
        "adaflo_b200_ctx_destroy": [vp],
        "adaflo_b200_ctx_set_stream": [vp, vp],
        "adaflo_b200_ctx_synchronize": [vp],
        "adaflo_b200_comm_unique_id": [vp],
        "adaflo_b200_ctx_init_comm": [vp, vp, i32, i32],
        "adaflo_b200_mf_create": [vp, C.POINTER(MFDesc), C.POINTER(vp)],
        "adaflo_b200_mf_destroy": [vp],
        "adaflo_b200_mf_vector_size": [vp, i32, C.POINTER(u64), C.POINTER(u64)],
        "adaflo_b200_mf_space_info": [vp, i32, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)],
        "adaflo_b200_mf_n_cells": [vp, C.POINTER(C.c_uint32)],
        "adaflo_b200_mf_set_dof_numbering": [vp, i32, vp],
        "adaflo_b200_mf_upload_vector": [vp, i32, vp, vp],
        "adaflo_b200_mf_download_vector": [vp, i32, vp, vp],
        "adaflo_b200_mf_update_ghost_values": [vp, i32, vp],
        "adaflo_b200_mf_compress_add": [vp, i32, vp],
        "adaflo_b200_mf_distribute_constraints": [vp, i32, vp],
        "adaflo_b200_ns_create": [vp, vp, C.POINTER(FlowParams), i32, i32, i32, i32, i32, C.POINTER(vp)],
        "adaflo_b200_ns_destroy": [vp],
        "adaflo_b200_ns_set_flow_params": [vp, C.POINTER(FlowParams)],
        "adaflo_b200_ns_set_time_weights": [vp, C.POINTER(TimeWeights)],
        "adaflo_b200_ns_set_old_solutions": [vp, vp, vp],
        "adaflo_b200_ns_vmult": [vp, vp, vp, vp, vp],
        "adaflo_b200_ns_Tvmult": [vp, vp, vp, vp, vp],
        "adaflo_b200_ns_residual": [vp, vp, vp, vp, vp, vp, vp],
        "adaflo_b200_ns_velocity_vmult": [vp, vp, vp],
        "adaflo_b200_ns_divergence_vmult_add": [vp, vp, vp, i32],
        "adaflo_b200_ns_pressure_poisson_vmult": [vp, vp, vp],
        "adaflo_b200_ns_pressure_mass_vmult": [vp, vp, vp],
        "adaflo_b200_ns_pressure_convdiff_vmult": [vp, vp, vp],
        "adaflo_b200_ns_compute_diagonal": [vp, vp, vp],
        "adaflo_b200_ns_fix_linearization_point": [vp],
        "adaflo_b200_ns_apply_pressure_average_projection": [vp, vp],
        "adaflo_b200_ns_apply_pressure_shift": [vp, dbl, vp],
        "adaflo_b200_ns_coefficient_ptr": [vp, i32, C.POINTER(vp), C.POINTER(u64)],
        "adaflo_b200_ns_set_coefficient_canonical": [vp, i32, vp, i32, C.c_uint32, C.c_uint32, vp],
        "adaflo_b200_ns_get_coefficient_canonical": [vp, i32, vp, i32, C.c_uint32, C.c_uint32, vp],
        "adaflo_b200_ns_nullspace": [vp, i32, C.POINTER(vp), C.POINTER(vp), C.POINTER(dbl)],
        "adaflo_b200_ns_get_matvec_statistics": [vp, C.POINTER(u64), C.POINTER(dbl)],
        "adaflo_b200_ns_vmult_host": [vp, vp, vp, vp, vp],
        "adaflo_b200_ns_velocity_vmult_host": [vp, vp, vp],
        "adaflo_b200_ns_pressure_poisson_vmult_host": [vp, vp, vp],
        "adaflo_b200_ns_pressure_mass_vmult_host": [vp, vp, vp],
        "adaflo_b200_ns_pressure_convdiff_vmult_host": [vp, vp, vp],
        "adaflo_b200_ns_divergence_vmult_add_host": [vp, vp, vp, i32],
        "adaflo_b200_ns_set_kernel_timing": [vp, i32],
        "adaflo_b200_ns_get_kernel_timing": [vp, C.POINTER(u64), C.POINTER(dbl)],
        "adaflo_b200_ls_create": [vp, vp, C.POINTER(LSParams), i32, i32, i32, vp, C.POINTER(vp)],
        "adaflo_b200_ls_destroy": [vp],
        "adaflo_b200_ls_set_time": [vp, dbl, dbl, dbl, dbl],
        "adaflo_b200_ls_set_diagonal": [vp, vp],
        "adaflo_b200_ls_mass_diagonal": [vp, vp],
        "adaflo_b200_ls_advance_concentration_vmult": [vp, vp, vp],
        "adaflo_b200_ls_advance_concentration_rhs": [vp, vp, vp, vp, vp, vp],
        "adaflo_b200_ls_compute_normal_vmult": [vp, C.POINTER(vp), C.POINTER(vp)],
        "adaflo_b200_ls_compute_normal_rhs": [vp, C.POINTER(vp), vp],
        "adaflo_b200_ls_compute_curvature_vmult": [vp, vp, vp, i32],
        "adaflo_b200_ls_compute_curvature_rhs": [vp, vp, C.POINTER(vp)],
        "adaflo_b200_ls_curvature_correction": [vp, vp, vp],
        "adaflo_b200_ls_reinitialization_vmult": [vp, vp, vp, i32],
        "adaflo_b200_ls_reinitialization_rhs": [vp, vp, vp, C.POINTER(vp), i32, i32],
        "adaflo_b200_ls_pointdata_ptr": [vp, i32, C.POINTER(vp), C.POINTER(u64)],
        "adaflo_b200_ls_compute_heaviside": [vp, vp, vp],
        "adaflo_b200_ls_compute_force": [vp, vp, vp, vp, vp, dbl, dbl, vp],
        "adaflo_b200_vec_alloc": [vp, u64, C.POINTER(vp)],
        "adaflo_b200_vec_free": [vp, vp],
        "adaflo_b200_vec_upload": [vp, vp, vp, u64],
        "adaflo_b200_vec_download": [vp, vp, vp, u64],
        "adaflo_b200_vec_set": [vp, vp, dbl, u64],
        "adaflo_b200_vec_axpy": [vp, vp, dbl, vp, u64],
        "adaflo_b200_vec_sadd": [vp, vp, dbl, dbl, vp, u64],
        "adaflo_b200_vec_scale": [vp, vp, vp, u64],
        "adaflo_b200_vec_dot": [vp, vp, vp, u64, C.POINTER(dbl)],
        "adaflo_b200_vec_multi_dot": [vp, vp, C.POINTER(vp), i32, u64, C.POINTER(dbl)],
        "adaflo_b200_vec_multi_axpy": [vp, vp, C.POINTER(dbl), C.POINTER(vp), i32, u64],
        "adaflo_b200_vec_add_and_dot": [vp, vp, dbl, vp, vp, u64, C.POINTER(dbl)],
        "adaflo_b200_jacobi_reinit": [vp, vp, vp, u64],
        "adaflo_b200_jacobi_vmult": [vp, vp, vp, vp, u64, i32],
        "adaflo_b200_solve_cg": [vp, i32, C.POINTER(u64), C.POINTER(u64), APPLY_FN, vp, APPLY_FN, vp, C.POINTER(vp), C.POINTER(vp),
                                 C.POINTER(SolverControlABI)],
        "adaflo_b200_solve_bicgstab": [vp, i32, C.POINTER(u64), C.POINTER(u64), APPLY_FN, vp, APPLY_FN, vp, C.POINTER(vp), C.POINTER(vp),
                                       C.POINTER(SolverControlABI)],
        "adaflo_b200_solve_gmres": [vp, i32, C.POINTER(u64), C.POINTER(u64), APPLY_FN, vp, APPLY_FN, vp, C.POINTER(vp), C.POINTER(vp), i32,
                                    C.POINTER(SolverControlABI)],
        "adaflo_b200_apply_operator": [vp, C.POINTER(vp), C.POINTER(vp)],
        "adaflo_b200_ns_preconditioner_create": [vp, C.POINTER(vp)],
        "adaflo_b200_ns_preconditioner_destroy": [vp],
        "adaflo_b200_ns_preconditioner_vmult": [vp, vp, vp, vp, vp],
        "adaflo_b200_ns_preconditioner_apply": [vp, C.POINTER(vp), C.POINTER(vp)],
        "adaflo_b200_ns_preconditioner_statistics": [vp, C.POINTER(i32), C.POINTER(i32)],
        "adaflo_b200_ls_projection_vmult": [vp, vp, vp],
        "adaflo_b200_ls_compute_normal": [vp, C.POINTER(vp), vp, i32, C.POINTER(i32)],
        "adaflo_b200_ls_compute_curvature": [vp, vp, vp, C.POINTER(vp), i32, i32, C.POINTER(i32)],
        "adaflo_b200_ls_reinitialize": [vp, vp, C.POINTER(vp), dbl, C.c_uint, C.c_uint, C.POINTER(dbl), C.POINTER(i32), C.POINTER(i32)],
        "adaflo_b200_ls_advance_concentration": [vp, vp, vp, vp, vp, dbl, C.POINTER(dbl), C.POINTER(i32)],
    }
    for name, args in sig.items():
        if os.environ.get("ADAFLO_B200_LIB") and not hasattr(lib, name):
            continue   # developer A/B runs against older builds of the library (tools/ab.sh); the product path has every symbol
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int
    lib.adaflo_b200_kernel_launch_count.argtypes = []
    lib.adaflo_b200_kernel_launch_count.restype = C.c_uint64
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise AdafloB200Error(f"adaflo_b200 error {rc}: {load().adaflo_b200_last_error().decode()}", rc)
