"""Shared set-up for the parity tests: the same mesh, parameters and seeded
vectors are handed to the CPU oracle and to the CUDA operator (through the C ABI)."""
import numpy as np

from adaflo_b200 import capi
from adaflo_b200.mesh import BoxMesh, beltrami_velocity, beltrami_pressure, seeded_uniform, sine_deformation, linear_deformation
from oracle import oracle as orc


def flow_params(physical_type=capi.INCOMPRESSIBLE, linearization=capi.NEWTON, augmented=False, density=1.0,
                viscosity=1.0, density_diff=0.0, viscosity_diff=0.0, damping=0.0, tau_grad_div=0.0, beta=0.5):
    dev = capi.FlowParams(physical_type, linearization, int(augmented), 0, density, viscosity, density_diff,
                          viscosity_diff, damping, tau_grad_div, beta)
    var = int(density_diff != 0 or viscosity_diff != 0)
    ora = orc.FlowParams(physical_type, linearization, int(augmented), var, density, viscosity, damping, tau_grad_div, beta)
    return dev, ora


def time_weights(dt=0.05, bdf2=True, step=3):
    """BDF-2 weights for uniform dt (source/time_stepping.cc:158-198)."""
    if bdf2 and step > 1:
        w, wo, woo = 1.5 / dt, -2.0 / dt, 0.5 / dt
    else:
        w, wo, woo = 1.0 / dt, -1.0 / dt, 0.0
    eo, eoo = (2.0, -1.0) if step > 2 else (1.0, 0.0)
    return capi.TimeWeights(w, wo, woo, 1.0, eo, eoo), orc.TimeWeights(w, wo, woo, 1.0, eo, eoo)


def deformation(geometry):
    if geometry == "general":
        return sine_deformation(0.04)
    if geometry == "affine":
        return linear_deformation(0.15)
    return None


def variable_coefficients(mesh, deform=None):
    """Smooth two-phase-like rho, mu and a damping field at the quadrature points."""
    xq = mesh.quadrature_points(3, deform)
    r = np.linalg.norm(xq, axis=-1)
    H = 0.5 * (1 + np.tanh((0.6 - r) / 0.2))
    rho = 1.0 + 9.0 * H
    mu = 1.0 + 0.5 * H
    damp = -0.3 * H
    return np.ascontiguousarray(rho), np.ascontiguousarray(mu), np.ascontiguousarray(damp)


def rel_err(a, b):
    a = np.asarray(a)
    b = np.asarray(b)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
