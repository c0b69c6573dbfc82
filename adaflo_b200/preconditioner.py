"""Device-resident block preconditioner for the linearised Navier-Stokes system.

Same structure as adaflo's NavierStokesPreconditioner (source/navier_stokes_preconditioner.cc:593-737): block-
triangular with a Cahouet-Chabard Schur complement approximation built from the two operators the reference
calls for exactly this purpose,

    S^-1  ~  -( pressure_mass_vmult^-1 + pressure_poisson_vmult^-1 ),        (M_p / (mu + tau_gd))^-1 + (L_p / (gamma rho))^-1
    z_p = S^-1 r_p,      z_u = A_u^-1 (r_u - G z_p),

but with matrix-free inner solves on the device instead of Trilinos ILU/AMG on the host: Jacobi-preconditioned CG
for the two pressure operators, Jacobi-preconditioned BiCGStab on velocity_vmult for the velocity block.  The inner
solves are inexact, so the outer solver must be flexible (adaflo_b200.solvers.SolverGMRES stores the preconditioned
basis).  Iteration counts therefore differ from the reference's golden logs, which belong to its ILU/AMG.
"""
import torch

from .operators import BlockVector, DiagonalPreconditioner
from .solvers import NoConvergence, SolverBicgstab, SolverCG, SolverControl


class NavierStokesPreconditioner:
    def __init__(self, ctx, matrix, matrix_free, velocity_tol=1e-2, pressure_tol=1e-2, max_inner=100):
        self.ctx, self.A, self.mf = ctx, matrix, matrix_free
        self.velocity_tol, self.pressure_tol, self.max_inner = velocity_tol, pressure_tol, max_inner
        self.nu, self.np = matrix_free.n_owned(0), matrix_free.n_owned(1)
        diag = BlockVector.like(matrix_free)
        matrix.compute_diagonal(diag)
        self.jacobi_u = DiagonalPreconditioner(ctx)
        self.jacobi_u.reinit(diag.block(0), self.nu)
        # diagonals of the two pressure operators by probing would cost n_p operator calls; their Jacobi
        # preconditioners use the lumped mass (row sums), which is spectrally equivalent on uniform meshes
        ones = torch.ones_like(diag.block(1))
        lump = torch.zeros_like(ones)
        matrix.pressure_mass_vmult(lump, ones)
        self.jacobi_m = DiagonalPreconditioner(ctx)
        self.jacobi_m.reinit(lump, self.np)
        self._tmp = BlockVector.like(matrix_free)
        self._zero_u = torch.zeros_like(diag.block(0))
        self._a, self._b = torch.zeros_like(ones), torch.zeros_like(ones)
        self._rhs_u = torch.zeros_like(diag.block(0))
        self.inner_iterations = [0, 0, 0]      # velocity, pressure mass, pressure Poisson

    def _solve(self, solver_cls, op, x, b, P, n, tol, slot):
        x.zero_()
        ctl = SolverControl(self.max_inner, 1e-300, tol)
        try:
            solver_cls(self.ctx, ctl).solve(op, x, b, P, n)
        except NoConvergence:
            pass
        self.inner_iterations[slot] += ctl.last_step

    def vmult(self, dst, src):
        """dst, src: [u, p] lists of device vectors (the convention of adaflo_b200.solvers)"""
        A, ru, rp = self.A, src[0], src[1]
        rp0 = rp - rp[: self.np].mean()        # consistent right-hand side for the singular Neumann problem
        self._solve(SolverCG, lambda d, s: A.pressure_mass_vmult(d, s), self._a, rp0, lambda d, s: self.jacobi_m.vmult(d, s), self.np,
                    self.pressure_tol, 1)
        self._solve(SolverCG, lambda d, s: A.pressure_poisson_vmult(d, s), self._b, rp0, None, self.np, self.pressure_tol, 2)
        dst[1].copy_(self._a).add_(self._b).mul_(-1.0)
        # r_u - G z_p: the velocity rows of vmult applied to (0, z_p)
        A.vmult(self._tmp, BlockVector(self._zero_u, dst[1]))
        torch.sub(ru, self._tmp.block(0), out=self._rhs_u)
        self._solve(SolverBicgstab, lambda d, s: A.velocity_vmult(d, s), dst[0], self._rhs_u, lambda d, s: self.jacobi_u.vmult(d, s),
                    self.nu, self.velocity_tol, 0)
