"""kernel time and roofline fraction of the vmult variants of SURVEY 8(d): S (Stokes), N (Newton), P (Picard),
V (Newton + variable coefficients), G (Newton + general geometry), T (Newton transpose), U (velocity block)"""
import os, sys, json
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adaflo_b200 import capi
from adaflo_b200.mesh import BoxMesh, beltrami_velocity, seeded_uniform, sine_deformation
from adaflo_b200.operators import BlockVector, Context, MatrixFree, NavierStokesMatrix

PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))).get("hbm_gbs", 6547.5)
ctx = Context(0)
dt = 0.05
tw = capi.TimeWeights(1.5 / dt, -2.0 / dt, 0.5 / dt, 1.0, 2.0, -1.0)


def run(name, n, bytes_per_dof, phys=capi.INCOMPRESSIBLE, lin=capi.NEWTON, var=False, geometry="cartesian", op="vmult"):
    mesh = BoxMesh((n, n, n))
    deform = sine_deformation(0.04) if geometry == "general" else None
    mf = MatrixFree(ctx, mesh, geometry, deform)
    prm = capi.FlowParams(phys, lin, 0, 0, 1.0, 1.0, 9.0 if var else 0.0, 0.5 if var else 0.0, 0.0, 0.0, 0.5)
    sol, old, old_old = BlockVector.like(mf), BlockVector.like(mf), BlockVector.like(mf)
    A = NavierStokesMatrix(prm, 0, 1, 0, 1, old, old_old)
    A.initialize(mf, tw, True)
    sol.block(0).copy_(torch.from_numpy(beltrami_velocity(mesh.u.coords, 0.0).reshape(-1)).to(ctx.device))
    old.block(0).copy_(sol.block(0)); old_old.block(0).copy_(sol.block(0))
    if var:
        for which, base in ((capi.COEF_DENSITY, 1.0), (capi.COEF_VISCOSITY, 1.0), (capi.COEF_DAMPING, 0.0)):
            A.set_coefficient(which, torch.full((mesh.n_cells, 27), base + 0.1, dtype=torch.float64, device=ctx.device))
    mesh.u.coords = None
    rhs, zero = BlockVector.like(mf), BlockVector.like(mf)
    if phys != capi.STOKES:
        A.residual(rhs, sol, zero)
    src = BlockVector(torch.from_numpy(seeded_uniform(mesh.u.global_ids, 3, 1)).to(ctx.device),
                      torch.from_numpy(seeded_uniform(mesh.p.global_ids, 1, 2)).to(ctx.device))
    dst = BlockVector.like(mf)
    f = {"vmult": lambda: A.vmult(dst, src), "Tvmult": lambda: A.Tvmult(dst, src),
         "velocity": lambda: A.velocity_vmult(dst.block(0), src.block(0))}[op]
    for _ in range(5):
        f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        f()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    n_dofs = mf.n_owned(0) + (mf.n_owned(1) if op != "velocity" else 0)
    gbs = bytes_per_dof * n_dofs / (ms * 1e-3) / 1e9
    print(f"{name:28s} n={n:3d} {n_dofs/1e6:6.2f} MDoF  {ms:6.3f} ms/call  {n_dofs/ms/1e6:6.2f} GDoF/s  {gbs:7.1f} GB/s algorithmic  {100*gbs/PEAK:5.1f} % of {PEAK:.0f}", flush=True)
    A.clear(); mf.close()
    del A, mf, mesh, src, dst, rhs, zero, sol, old, old_old
    torch.cuda.empty_cache()


run("S  Stokes", 128, 21.6, phys=capi.STOKES)
run("N  Newton", 128, 125.3)
run("P  Picard", 128, 21.6 + 34.56, lin=capi.PICARD)
run("V  Newton + variable rho/mu", 128, 151.2, var=True)
run("T  Newton Tvmult", 128, 125.3, op="Tvmult")
run("U  velocity block (Newton)", 128, 125.3 - 0.32 - 8 / 25 * 0, op="velocity")
run("G  Newton + general geometry", 96, 211.7, geometry="general")
