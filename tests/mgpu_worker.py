"""Worker for the multi-GPU parity test (launched by torchrun, one rank per GPU).

Every rank builds its z-slab of the same global box, runs residual + vmult through the
C ABI (ghost import / compress(add) / mean projection over NCCL inside the library) and
compares its owned entries with the single-rank CPU oracle evaluated on the global mesh.
Exit code 0 = parity within 1e-12.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from adaflo_b200 import capi  # noqa: E402
from adaflo_b200.mesh import BoxMesh, beltrami_velocity, seeded_uniform  # noqa: E402
from adaflo_b200.operators import BlockVector, Context, MatrixFree, NavierStokesMatrix  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from helpers import flow_params, time_weights  # noqa: E402


def level_set_checks(ctx, rank, world):
    """Level-set operators on the partitioned mesh (ghost import of the sources, compress(add) of the results inside
    the library, MatrixFree::cell_loop semantics) against the dense oracle on the global mesh."""
    from adaflo_b200.operators import LevelSetOperators
    n_sub = 3
    n = (4, 3, 2 * world + 1)
    mesh, gmesh = BoxMesh(n, rank=rank, nranks=world), BoxMesh(n)
    mesh.build_ls_space(n_sub, (0, 3))
    gmesh.build_ls_space(n_sub, (0, 3))
    mf = MatrixFree(ctx, mesh, "cartesian")
    h_min = float(mesh.h.min())
    eps, eps_used = 1.5, 1.5 * h_min / n_sub
    A = LevelSetOperators(mf, capi.LSParams(3, n_sub, 0, 0, eps, eps_used, 4.0, h_min), mesh.cell_diameters())
    O = orc.OracleLevelSet(gmesh, orc.LSParams(eps, eps_used, 4.0, h_min, 0.02, 3, 0))
    gamma, w_old, w_oo, dt = 75.0, -100.0, 25.0, 0.02
    A.set_time(gamma, w_old, w_oo, dt)
    diag = O.mass_diagonal()
    O.diag = diag
    own = mesh.ls.n_owned
    gid = mesh.ls.global_ids
    ok = True

    def dev(a):
        return torch.from_numpy(np.ascontiguousarray(a)).to(ctx.device)

    def loc(gvec, garbage=True):
        """local copy of a global LS vector; ghost entries start with garbage (the library must import them)"""
        v = np.asarray(gvec)[gid].copy()
        if garbage:
            v[own:] = 321.0
        return dev(v)

    def check(name, got, ref):
        err = np.linalg.norm(got.cpu().numpy()[:own] - ref[gid[:own]]) / max(np.linalg.norm(ref), 1e-300)
        good = err < 1e-12
        print(f"[rank {rank}] ls {name}: rel err {err:.2e} {'ok' if good else 'FAIL'}", flush=True)
        return good

    def gvec(seed):
        return seeded_uniform(gmesh.ls.global_ids, 1, seed)

    d = A.initialize_mass_matrix_diagonal()
    ok &= check("mass_diagonal", d, diag)
    A.set_diagonal(loc(diag, garbage=False))
    x = gmesh.ls.coords
    r = np.linalg.norm(x - np.array([0.1, -0.05, -0.2]), axis=1)
    c = -np.tanh((r - 0.45) / (2 * 0.12)) + 0.01 * gvec(5)
    gvel = beltrami_velocity(gmesh.u.coords, 0.0).reshape(-1) + 0.1 * seeded_uniform(gmesh.u.global_ids, 3, 6)
    vel = gvel.reshape(-1, 3)[mesh.u.global_ids].copy()
    vel[mesh.u.n_owned:] = 7.0
    c_old, c_oo = c + 0.05 * gvec(7), c + 0.05 * gvec(8)
    o = np.zeros(len(c))
    O.advance_concentration_rhs(o, gvel, c, c_old, c_oo, gamma, w_old, w_oo)
    g = mf.initialize_dof_vector(2)
    A.advance_concentration_rhs(g, dev(vel.reshape(-1)), loc(c), loc(c_old), loc(c_oo))
    ok &= check("advance_concentration_rhs", g, o)
    src = gvec(10)
    g = mf.initialize_dof_vector(2)
    A.advance_concentration_vmult(g, loc(src))
    ok &= check("advance_concentration_vmult", g, O.advance_concentration_vmult(src, gamma))
    rhs_o = [np.zeros(len(c)) for _ in range(3)]
    O.compute_normal_rhs(rhs_o, c)
    rhs_g = [mf.initialize_dof_vector(2) for _ in range(3)]
    A.compute_normal_rhs(rhs_g, loc(c))
    for k in range(3):
        ok &= check(f"compute_normal_rhs[{k}]", rhs_g[k], rhs_o[k])
    nsrc = [gvec(11 + k) for k in range(3)]
    ndst = [mf.initialize_dof_vector(2) for _ in range(3)]
    A.compute_normal_vmult(ndst, [loc(v) for v in nsrc])
    for k, ref in enumerate(O.compute_normal_vmult(nsrc)):
        ok &= check(f"compute_normal_vmult[{k}]", ndst[k], ref)
    normal = [x[:, k] - [0.1, -0.05, -0.2][k] + 0.02 * gvec(20 + k) for k in range(3)]
    o = np.zeros(len(c))
    O.compute_curvature_rhs(o, normal)
    g = mf.initialize_dof_vector(2)
    A.compute_curvature_rhs(g, [loc(v) for v in normal])
    ok &= check("compute_curvature_rhs", g, o)
    g = mf.initialize_dof_vector(2)
    A.compute_curvature_vmult(g, loc(src), True)
    ok &= check("compute_curvature_vmult", g, O.compute_curvature_vmult(src, True))
    o = np.zeros(len(c))
    O.reinitialization_rhs(o, c, normal, False, True)
    g = mf.initialize_dof_vector(2)
    A.reinitialization_rhs(g, loc(c), [loc(v) for v in normal], False, True)
    ok &= check("reinitialization_rhs", g, o)
    g = mf.initialize_dof_vector(2)
    A.reinitialization_vmult(g, loc(src), False)
    ok &= check("reinitialization_vmult", g, O.reinitialization_vmult(src, False))
    A.close()
    mf.close()
    return ok


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = Context(local)

    def bcast(b):
        obj = [b]
        dist.broadcast_object_list(obj, src=0)
        return obj[0]
    ctx.init_comm(rank, world, bcast)

    n = (6, 5, 6 * world + 1)  # >= 18 decades per rank: the halo import overlaps the interior decades
    mesh = BoxMesh(n, rank=rank, nranks=world)
    gmesh = BoxMesh(n)
    dprm, oprm = flow_params(damping=-0.2, tau_grad_div=0.3)
    dtw, otw = time_weights()
    mf = MatrixFree(ctx, mesh, "cartesian")
    old, old_old = BlockVector.like(mf), BlockVector.like(mf)
    A = NavierStokesMatrix(dprm, 0, 1, 0, 1, old, old_old)
    A.initialize(mf, dtw, True)
    O = orc.OracleNavierStokesMatrix(gmesh, oprm, otw)

    def dev(a):
        return torch.from_numpy(np.ascontiguousarray(a)).to(ctx.device)

    def fields(m, seed_shift=0):
        u = beltrami_velocity(m.u.coords, 0.1).reshape(-1) + 0.1 * seeded_uniform(m.u.global_ids, 3, 7 + seed_shift)
        p = seeded_uniform(m.p.global_ids, 1, 8 + seed_shift)
        return u, p

    def owned(m, sp, vec, ncomp):
        return vec.reshape(-1, ncomp)[: sp.n_owned]

    def check(name, got_u, got_p, ref_u, ref_p):
        # compare owned entries against the global oracle vector via global ids
        gu = ref_u.reshape(-1, 3)[mesh.u.global_ids[: mesh.u.n_owned]]
        gp = ref_p[mesh.p.global_ids[: mesh.p.n_owned]]
        eu = np.linalg.norm(owned(mesh, mesh.u, got_u, 3) - gu) / np.linalg.norm(ref_u)
        ep = np.linalg.norm(got_p[: mesh.p.n_owned] - gp) / np.linalg.norm(ref_p)
        ok = eu < 1e-12 and ep < 1e-12
        print(f"[rank {rank}] {name}: rel err u {eu:.2e} p {ep:.2e} {'ok' if ok else 'FAIL'}", flush=True)
        return ok

    # residual (writes the linearisation point); ghosts of the inputs are imported inside
    u, p = fields(mesh)
    Gu, Gp = fields(gmesh)
    uo, _ = fields(mesh, 10)
    Guo, _ = fields(gmesh, 10)
    old.block(0).copy_(dev(uo))
    old_old.block(0).copy_(dev(uo))
    O.old_u, O.old_old_u = Guo.copy(), Guo.copy()
    ru, rp = np.zeros(O.n_u), np.zeros(O.n_p)
    O.residual(ru, rp, Gu, Gp, np.zeros(O.n_u), np.zeros(O.n_p))
    rhs = BlockVector.like(mf)
    A.residual(rhs, BlockVector(dev(u), dev(p)), BlockVector.like(mf))
    ok = check("residual", rhs.block(0).cpu().numpy(), rhs.block(1).cpu().numpy(), ru, rp)

    su, sp = seeded_uniform(mesh.u.global_ids, 3, 31), seeded_uniform(mesh.p.global_ids, 1, 32)
    Gsu, Gsp = seeded_uniform(gmesh.u.global_ids, 3, 31), seeded_uniform(gmesh.p.global_ids, 1, 32)
    # ghost entries of the source start with garbage: vmult must import them itself
    su[mesh.u.n_owned * 3:] = 123.0
    ou, op = O.vmult(Gsu, Gsp)
    dst = BlockVector.like(mf)
    A.vmult(dst, BlockVector(dev(su), dev(sp)))
    ok &= check("vmult", dst.block(0).cpu().numpy(), dst.block(1).cpu().numpy(), ou, op)

    # host entry point: owned entries in, owned entries out; export range uploaded first, ghost import and
    # compress(add) inside the three-stream pipeline
    nu_own, np_own = mesh.u.n_owned * 3, mesh.p.n_owned
    for rep in range(2):
        hu, hp = torch.from_numpy(su[:nu_own].copy()).pin_memory(), torch.from_numpy(sp[:np_own].copy()).pin_memory()
        du, dp = torch.full_like(hu, 5.0).pin_memory(), torch.full_like(hp, 5.0).pin_memory()
        A.vmult_host(du, dp, hu, hp)
        gu_ = np.zeros(mesh.u.n_local * 3) if hasattr(mesh.u, "n_local") else np.zeros_like(su)
        gu_[:nu_own] = du.numpy()
        gp_ = np.zeros_like(sp)
        gp_[:np_own] = dp.numpy()
        ok &= check(f"vmult_host #{rep}", gu_, gp_, ou, op)

    # global dot through the library (all-reduce inside)
    x = dev(seeded_uniform(mesh.u.global_ids, 3, 41))
    d = ctx.dot(x, x, mesh.u.n_owned * 3)
    gx = seeded_uniform(gmesh.u.global_ids, 3, 41)
    okd = abs(d - float(gx @ gx)) < 1e-11 * float(gx @ gx)
    print(f"[rank {rank}] dot: {d:.15e} vs {float(gx @ gx):.15e} {'ok' if okd else 'FAIL'}", flush=True)
    ok &= okd

    ok &= level_set_checks(ctx, rank, world)

    flag = torch.tensor([0 if ok else 1], device=ctx.device)
    dist.all_reduce(flag)
    A.clear()
    mf.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(int(flag.item() != 0))


if __name__ == "__main__":
    main()
