"""ctypes front end of oracle/ns_cpu_port.c -- the timed CPU baseline ("port").
TEST/BENCH INFRASTRUCTURE ONLY (see the header of ns_cpu_port.c)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


class PortArgs(C.Structure):
    _fields_ = [("n_cells", C.c_int32), ("nx", C.c_int32), ("ny", C.c_int32), ("cell_u", C.c_void_p), ("cell_p", C.c_void_p),
                ("inv_h", C.c_double * 3), ("det", C.c_double), ("mode", C.c_int32),
                ("gamma", C.c_double), ("tau1", C.c_double), ("beta", C.c_double), ("rho", C.c_double),
                ("mu", C.c_double), ("damping", C.c_double), ("tau_gd", C.c_double), ("lin", C.c_void_p), ("lin_b", C.c_void_p)]


def build(force=False):
    so = os.path.join(_HERE, "libcpu_port.so")
    src = os.path.join(_HERE, "ns_cpu_port.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O3", "-march=x86-64-v3", "-ffp-contract=fast", "-fopenmp", "-fPIC", "-std=c11", "-shared", "-o", so, src, "-lm"])
    return so


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


class CpuPortVmult:
    """Cell loop of NavierStokesMatrix::vmult on a Cartesian BoxMesh (constant coefficients)."""

    def __init__(self, mesh, mode, gamma, tau1, beta, rho, mu, damping, tau_gd, lin=None):
        self.mesh = mesh
        self.cu = np.ascontiguousarray(mesh.u.encoded_cell_nodes())
        self.cp = np.ascontiguousarray(mesh.p.encoded_cell_nodes())
        self.lin = None if lin is None else np.ascontiguousarray(lin)
        a = PortArgs()
        a.n_cells = mesh.n_cells
        a.nx, a.ny = mesh.n[0], mesh.n[1]
        a.cell_u = self.cu.ctypes.data
        a.cell_p = self.cp.ctypes.data
        for d in range(3):
            a.inv_h[d] = 1.0 / mesh.h[d]
        a.det = float(np.prod(mesh.h))
        a.mode = mode
        a.gamma, a.tau1, a.beta, a.rho, a.mu, a.damping, a.tau_gd = gamma, tau1, beta, rho, mu, damping, tau_gd
        a.lin = None if self.lin is None else self.lin.ctypes.data
        # the linearisation point in MatrixFree's layout: batches of 4 x-adjacent cells, [row][nx // 4][12 rows][27 points][4 lanes]
        self.lin_b = None
        nb = mesh.n[0] // 4
        if self.lin is not None and nb > 0:
            rows = mesh.n_cells // mesh.n[0]
            l4 = self.lin.reshape(rows, mesh.n[0], 12, 27)[:, :4 * nb].reshape(rows, nb, 4, 12, 27)
            self.lin_b = np.ascontiguousarray(l4.transpose(0, 1, 3, 4, 2))
        a.lin_b = None if self.lin_b is None else self.lin_b.ctypes.data
        self.args = a

    def cells(self, src_u, src_p, dst_u, dst_p):
        return lib().port_ns_vmult_cells(C.byref(self.args), src_u.ctypes.data_as(C.c_void_p), src_p.ctypes.data_as(C.c_void_p),
                                         dst_u.ctypes.data_as(C.c_void_p), dst_p.ctypes.data_as(C.c_void_p))


def max_threads():
    return lib().port_max_threads()


def use_all_cores():
    """Run the port on every core of this process's affinity mask, whatever OMP_NUM_THREADS says (torchrun exports
    OMP_NUM_THREADS=1 to its workers); returns the thread count now in effect."""
    lib().port_set_threads(len(os.sched_getaffinity(0)))
    return max_threads()
