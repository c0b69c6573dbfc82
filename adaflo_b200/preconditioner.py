"""Device-resident block preconditioner for the linearised Navier-Stokes system.

Thin wrapper over the library's adaflo_b200_ns_preconditioner_* (csrc/precond.cu): NavierStokesPreconditioner::vmult with
inner solves (source/navier_stokes_preconditioner.cc:593-737) -- lower block-triangular,

    z_u = A_u^-1 r_u,   t = -r_p + B z_u,   z_p = M_p^-1 t + L_p^-1 t       (Cahouet-Chabard; the Poisson part only if density > 0)

with matrix-free inner solves on the device (Jacobi-preconditioned BiCGStab on velocity_vmult, CG on pressure_mass_vmult and
pressure_poisson_vmult) instead of the reference's Trilinos ILU / AMG on the host.  The inner solves are inexact, so the outer
solver must be flexible (adaflo_b200.solvers.SolverGMRES is).  Iteration counts therefore differ from the reference's golden logs,
which belong to its ILU / AMG.
"""
import ctypes as C

from . import capi
from .capi import check


class NavierStokesPreconditioner:
    def __init__(self, ctx, matrix, matrix_free=None):
        """matrix: an initialised NavierStokesMatrix whose linearisation point is set (residual() has been called)"""
        self.ctx, self.A = ctx, matrix
        self.h = C.c_void_p()
        check(ctx.lib.adaflo_b200_ns_preconditioner_create(matrix.h, C.byref(self.h)))

    def vmult(self, dst, src):
        """dst, src: [u, p] lists of device vectors (the convention of adaflo_b200.solvers)"""
        check(self.ctx.lib.adaflo_b200_ns_preconditioner_vmult(self.h, C.c_void_p(dst[0].data_ptr()), C.c_void_p(dst[1].data_ptr()),
                                                               C.c_void_p(src[0].data_ptr()), C.c_void_p(src[1].data_ptr())))

    def as_library_callback(self):
        """(APPLY_FN, user pointer) for solves that stay inside the library"""
        return C.cast(self.ctx.lib.adaflo_b200_ns_preconditioner_apply, capi.APPLY_FN), self.h

    @property
    def inner_iterations(self):
        n, its = C.c_int(), (C.c_int * 3)()
        check(self.ctx.lib.adaflo_b200_ns_preconditioner_statistics(self.h, C.byref(n), its))
        return [its[0], its[1], its[2]]

    def close(self):
        if self.h:
            self.ctx.lib.adaflo_b200_ns_preconditioner_destroy(self.h)
            self.h = C.c_void_p()

    # no __del__: the handle refers to the operator; at interpreter shutdown the destruction order is arbitrary.  Call close().
