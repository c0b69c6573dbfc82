"""Krylov drivers on device vectors: the loops deal.II's SolverCG / SolverBicgstab /
SolverGMRES run around the operators (callers: source/level_set_okz_reinitialization.cc:334-336,
source/level_set_okz_advance_concentration.cc:626-644, source/navier_stokes.cc:593-609).  The host drives
the iteration; every vector operation is a fused kernel of libadaflo_b200 (axpy / sadd / add_and_dot /
multi_dot), every reduction result is one small D2H copy.  Vectors are torch tensors used as device
buffers only.

Stopping rule = deal.II ReductionControl: stop when ||r|| <= max(abs_tol, rel_tol * ||r_0||).
"""
from __future__ import annotations

import math

import torch


class SolverControl:
    def __init__(self, max_iter=1000, abs_tol=1e-12, rel_tol=0.0):
        self.max_iter, self.abs_tol, self.rel_tol = max_iter, abs_tol, rel_tol
        self.last_step = 0
        self.last_value = 0.0
        self.history = []

    def target(self, r0):
        return max(self.abs_tol, self.rel_tol * r0)


class NoConvergence(RuntimeError):
    """SolverControl::NoConvergence"""


class SolverCG:
    """Preconditioned conjugate gradients."""

    def __init__(self, ctx, control: SolverControl):
        self.ctx, self.control = ctx, control

    def solve(self, A, x, b, P=None, n=None):
        ctx, ctl = self.ctx, self.control
        n = n if n is not None else x.numel()
        r, z, p, Ap = (torch.zeros_like(x) for _ in range(4))
        A(Ap, x)
        r.copy_(b)
        ctx.axpy(r, -1.0, Ap)                       # r = b - A x
        res = math.sqrt(ctx.dot(r, r, n))
        ctl.history = [res]
        tol = ctl.target(res)
        it = 0
        if res > tol:
            (P(z, r) if P else z.copy_(r))
            p.copy_(z)
            rz = ctx.dot(r, z, n)
            while True:
                it += 1
                A(Ap, p)
                alpha = rz / ctx.dot(p, Ap, n)
                ctx.axpy(x, alpha, p)
                ctx.axpy(r, -alpha, Ap)
                res = math.sqrt(ctx.dot(r, r, n))
                ctl.history.append(res)
                if res <= tol:
                    break
                if it >= ctl.max_iter:
                    ctl.last_step, ctl.last_value = it, res
                    raise NoConvergence(f"CG: {it} iterations, residual {res}")
                (P(z, r) if P else z.copy_(r))
                rz_new = ctx.dot(r, z, n)
                ctx.sadd(p, rz_new / rz, 1.0, z)    # p = z + beta p
                rz = rz_new
        ctl.last_step, ctl.last_value = it, res
        return it


class SolverBicgstab:
    """BiCGStab (van der Vorst) with right... left preconditioning as in deal.II (preconditioned directions)."""

    def __init__(self, ctx, control: SolverControl):
        self.ctx, self.control = ctx, control

    def solve(self, A, x, b, P=None, n=None):
        ctx, ctl = self.ctx, self.control
        n = n if n is not None else x.numel()
        r, rbar, p, v, y, s, z, t = (torch.zeros_like(x) for _ in range(8))
        A(v, x)
        r.copy_(b)
        ctx.axpy(r, -1.0, v)
        rbar.copy_(r)
        res = math.sqrt(ctx.dot(r, r, n))
        ctl.history = [res]
        tol = ctl.target(res)
        rho = alpha = omega = 1.0
        v.zero_()
        it = 0
        while res > tol:
            it += 1
            rho_new = ctx.dot(rbar, r, n)
            beta = (rho_new / rho) * (alpha / omega)
            ctx.axpy(p, -omega, v)                  # p = r + beta (p - omega v)
            ctx.sadd(p, beta, 1.0, r)
            (P(y, p) if P else y.copy_(p))
            A(v, y)
            alpha = rho_new / ctx.dot(rbar, v, n)
            s.copy_(r)
            ctx.axpy(s, -alpha, v)
            (P(z, s) if P else z.copy_(s))
            A(t, z)
            tt, ts = ctx.multi_dot(t, [t, s], n)
            omega = ts / tt
            ctx.axpy(x, alpha, y)
            ctx.axpy(x, omega, z)
            r.copy_(s)
            ctx.axpy(r, -omega, t)
            rho = rho_new
            res = math.sqrt(ctx.dot(r, r, n))
            ctl.history.append(res)
            if it >= ctl.max_iter and res > tol:
                ctl.last_step, ctl.last_value = it, res
                raise NoConvergence(f"BiCGStab: {it} iterations, residual {res}")
        ctl.last_step, ctl.last_value = it, res
        return it


class SolverGMRES:
    """Restarted GMRES(m), right preconditioned, classical Gram-Schmidt with one fused multi-dot per step
    (and a second pass for stability).  x and b are lists of device vectors (blocks)."""

    def __init__(self, ctx, control: SolverControl, max_basis=50):
        self.ctx, self.control, self.m = ctx, control, max_basis

    def _dots(self, w, V, ns):
        out = [0.0] * len(V)
        for blk, nb in enumerate(ns):
            for j0 in range(0, len(V), 64):          # the library reduces at most 64 dots per call
                part = self.ctx.multi_dot(w[blk], [v[blk] for v in V[j0:j0 + 64]], nb)
                for j, val in enumerate(part):
                    out[j0 + j] += val
        return out

    def _maxpy(self, y, coeffs, X):
        """y += sum_i coeffs[i] X[i], block by block, one pass over y per 8 vectors"""
        for blk in range(len(y)):
            self.ctx.multi_axpy(y[blk], coeffs, [v[blk] for v in X])

    def solve(self, A, x, b, P=None, ns=None):
        ctx, ctl, m = self.ctx, self.control, self.m
        ns = ns if ns is not None else [v.numel() for v in x]
        nb = len(x)
        new = lambda: [torch.zeros_like(v) for v in x]
        w, z = new(), new()
        it_total = 0
        tol = None
        while True:
            A(w, x)
            r = [bi.clone() for bi in b]
            for k in range(nb):
                ctx.axpy(r[k], -1.0, w[k])
            beta = math.sqrt(self._dots(r, [r], ns)[0])
            if tol is None:
                tol = ctl.target(beta)
                ctl.history = [beta]
            if beta <= tol:
                break
            V = [r]
            for k in range(nb):
                V[0][k].mul_(1.0 / beta)
            H = []
            g = [beta]
            cs, sn = [], []
            Z = []
            res = beta
            for j in range(m):
                it_total += 1
                if P:
                    P(z, V[j])
                    Z.append([t.clone() for t in z])
                    A(w, z)
                else:
                    A(w, V[j])
                # classical Gram-Schmidt in two vector passes: V^T w and |w|^2 fused into one multi-dot, w -= V h
                # fused into one multi-axpy.  The second (re-orthogonalisation) pass is taken only on severe
                # cancellation |w_new| < |w_old| / 10, which bounds the loss of orthogonality by ~10 eps per step.
                hw = self._dots(w, V + [w], ns)
                h, ww = hw[:-1], hw[-1]
                self._maxpy(w, [-a for a in h], V)
                hn2 = self._dots(w, [w], ns)[0]
                if hn2 < 1e-2 * ww:
                    h2 = self._dots(w, V, ns)
                    self._maxpy(w, [-a for a in h2], V)
                    h = [a + c for a, c in zip(h, h2)]
                    hn2 = self._dots(w, [w], ns)[0]
                hn = math.sqrt(max(hn2, 0.0))
                col = h + [hn]
                for i in range(j):                   # apply previous Givens rotations
                    t0 = cs[i] * col[i] + sn[i] * col[i + 1]
                    col[i + 1] = -sn[i] * col[i] + cs[i] * col[i + 1]
                    col[i] = t0
                d = math.hypot(col[j], col[j + 1])
                cs.append(col[j] / d)
                sn.append(col[j + 1] / d)
                col[j] = d
                col[j + 1] = 0.0
                g.append(-sn[j] * g[j])
                g[j] = cs[j] * g[j]
                H.append(col)
                res = abs(g[j + 1])
                ctl.history.append(res)
                if hn > 0:
                    V.append([t * (1.0 / hn) for t in w])
                if res <= tol or it_total >= ctl.max_iter or hn == 0:
                    break
            k_dim = len(H)
            y = [0.0] * k_dim
            for i in range(k_dim - 1, -1, -1):
                s = g[i] - sum(H[c][i] * y[c] for c in range(i + 1, k_dim))
                y[i] = s / H[i][i]
            self._maxpy(x, y, Z if P else V[:k_dim])
            if res <= tol:
                break
            if it_total >= ctl.max_iter:
                ctl.last_step, ctl.last_value = it_total, res
                raise NoConvergence(f"GMRES: {it_total} iterations, residual {res}")
        ctl.last_step, ctl.last_value = it_total, ctl.history[-1]
        return it_total
