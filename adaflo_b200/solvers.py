"""Krylov drivers on device vectors: the loops deal.II's SolverCG / SolverBicgstab / SolverGMRES run
around the operators (callers: source/level_set_okz_reinitialization.cc:334-336,
source/level_set_okz_advance_concentration.cc:626-644, source/navier_stokes.cc:593-609).

The iteration itself lives in the library (csrc/solvers.cu, C ABI adaflo_b200_solve_{cg,bicgstab,gmres}: callable from a
C++ host as well); these classes keep deal.II's calling convention on top of it.  An operator / preconditioner is either
  * a capi.OperatorRef (a library operator: the whole solve runs without touching Python), built by library_operator(), or
  * a Python callable A(dst, src) on torch tensors / lists of tensors, called back through a ctypes trampoline.
Vectors are torch tensors used as device buffers only.

Stopping rule = deal.II ReductionControl: stop when ||r|| <= max(abs_tol, rel_tol * ||r_0||).
"""
from __future__ import annotations

import ctypes as C

from . import capi
from .capi import check


class SolverControl:
    def __init__(self, max_iter=1000, abs_tol=1e-12, rel_tol=0.0):
        self.max_iter, self.abs_tol, self.rel_tol = max_iter, abs_tol, rel_tol
        self.last_step = 0
        self.last_value = 0.0
        self.initial_value = 0.0
        self.history = []

    def target(self, r0):
        return max(self.abs_tol, self.rel_tol * r0)


class NoConvergence(RuntimeError):
    """SolverControl::NoConvergence"""


def library_operator(kind, handle, flag=0, inv_diag=None, n=0, n_blocks=1):
    """capi.OperatorRef for adaflo_b200_apply_operator (kind: capi.OP_*; handle: the .h of the owning object)"""
    return capi.OperatorRef(kind, handle, int(flag), inv_diag.data_ptr() if inv_diag is not None else None, int(n), int(n_blocks))


class _Solver:
    """common driver: marshals block vectors and callbacks, maps ERR_NO_CONVERGENCE to NoConvergence"""
    entry = None
    name = ""

    def __init__(self, ctx, control: SolverControl):
        self.ctx, self.control = ctx, control

    def _callback(self, fn, blocks, as_list):
        """(APPLY_FN, user pointer, keep-alive) for a library operator, a Python callable or None"""
        lib = self.ctx.lib
        if fn is None:
            return C.cast(None, capi.APPLY_FN), None, None
        if isinstance(fn, tuple):          # (APPLY_FN, user pointer): a callback that lives in the library
            return fn[0], fn[1], None
        if isinstance(fn, capi.OperatorRef):
            return C.cast(lib.adaflo_b200_apply_operator, capi.APPLY_FN), C.cast(C.pointer(fn), C.c_void_p), fn
        from .operators import device_view
        device = blocks[0].device
        sizes = [b.numel() for b in blocks]
        err = []

        def trampoline(_user, dst, src):
            try:
                d = [device_view(dst[i], sizes[i], device) for i in range(len(sizes))]
                s = [device_view(src[i], sizes[i], device) for i in range(len(sizes))]
                fn(d, s) if as_list else fn(d[0], s[0])
                return 0
            except Exception as e:   # an exception must not unwind through the C frames
                err.append(e)
                return capi.ERR_NO_CONVERGENCE + 100
        cb = capi.APPLY_FN(trampoline)
        return cb, None, (cb, err)

    def _run(self, A, x, b, P, ns, as_list, extra=()):
        ctl = self.control
        nb = len(x)
        u64 = C.c_uint64 * nb
        n_owned = u64(*[int(n) for n in ns])
        n_local = u64(*[v.numel() for v in x])
        xs = (C.c_void_p * nb)(*[v.data_ptr() for v in x])
        bs = (C.c_void_p * nb)(*[v.data_ptr() for v in b])
        cA, uA, keepA = self._callback(A, x, as_list)
        cP, uP, keepP = self._callback(P, x, as_list)
        cap = int(ctl.max_iter) + 2
        hist = (C.c_double * cap)()
        c = capi.SolverControlABI(int(ctl.max_iter), float(ctl.abs_tol), float(ctl.rel_tol), 0, 0.0, 0.0, hist, cap, 0)
        rc = getattr(self.ctx.lib, self.entry)(self.ctx.h, nb, n_owned, n_local, cA, uA, cP, uP, xs, bs, *extra, C.byref(c))
        ctl.last_step, ctl.last_value, ctl.initial_value = c.last_step, c.last_value, c.initial_value
        ctl.history = [hist[i] for i in range(c.history_size)]
        for keep in (keepA, keepP):
            if isinstance(keep, tuple) and keep[1]:
                raise keep[1][0]
        if rc == capi.ERR_NO_CONVERGENCE:
            raise NoConvergence(f"{self.name}: {c.last_step} iterations, residual {c.last_value}")
        check(rc)
        return c.last_step


class SolverCG(_Solver):
    """Preconditioned conjugate gradients."""
    entry, name = "adaflo_b200_solve_cg", "CG"

    def solve(self, A, x, b, P=None, n=None):
        if isinstance(x, (list, tuple)):
            return self._run(A, list(x), list(b), P, n if n is not None else [v.numel() for v in x], True)
        return self._run(A, [x], [b], P, [n if n is not None else x.numel()], False)


class SolverBicgstab(_Solver):
    """BiCGStab (van der Vorst) with preconditioned directions as in deal.II."""
    entry, name = "adaflo_b200_solve_bicgstab", "BiCGStab"

    def solve(self, A, x, b, P=None, n=None):
        if isinstance(x, (list, tuple)):
            return self._run(A, list(x), list(b), P, n if n is not None else [v.numel() for v in x], True)
        return self._run(A, [x], [b], P, [n if n is not None else x.numel()], False)


class SolverGMRES(_Solver):
    """Restarted GMRES(m), right preconditioned and flexible, classical Gram-Schmidt with one fused multi-dot per step (and a
    second pass on severe cancellation).  x and b are lists of device vectors (blocks)."""
    entry, name = "adaflo_b200_solve_gmres", "GMRES"

    def __init__(self, ctx, control: SolverControl, max_basis=50):
        super().__init__(ctx, control)
        self.m = max_basis

    def solve(self, A, x, b, P=None, ns=None):
        return self._run(A, list(x), list(b), P, ns if ns is not None else [v.numel() for v in x], True, (int(self.m),))
