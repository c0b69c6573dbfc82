"""CPU worker (gloo, world_size 2) for the host-side partition logic: z-slab ownership,
ghost import and compress(add) plans produced by adaflo_b200.mesh.BoxMesh -- the same plans
the C ABI receives (adaflo_b200_exchange).  The cell loop itself is the CPU oracle here."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from adaflo_b200.mesh import BoxMesh, seeded_uniform  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from helpers import flow_params, time_weights  # noqa: E402


def update_ghost_values(sp, vec):
    v = vec.reshape(-1, sp.n_comp)
    reqs = []
    for peer, idx in sp.send:
        reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(v[idx])), peer))
    bufs = []
    for peer, first, cnt in sp.recv:
        b = torch.empty((cnt, sp.n_comp), dtype=torch.float64)
        bufs.append((first, cnt, b))
        reqs.append(dist.irecv(b, peer))
    for r in reqs:
        r.wait()
    for first, cnt, b in bufs:
        v[first:first + cnt] = b.numpy()


def compress_add(sp, vec):
    v = vec.reshape(-1, sp.n_comp)
    reqs = []
    for peer, first, cnt in sp.recv:
        reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(v[first:first + cnt])), peer))
    bufs = []
    for peer, idx in sp.send:
        b = torch.empty((len(idx), sp.n_comp), dtype=torch.float64)
        bufs.append((idx, b))
        reqs.append(dist.irecv(b, peer))
    for r in reqs:
        r.wait()
    for idx, b in bufs:
        v[idx] += b.numpy()
    v[sp.n_owned:] = 0.0


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    n = (4, 3, 5)
    mesh, gmesh = BoxMesh(n, rank=rank, nranks=world), BoxMesh(n)
    _, oprm = flow_params(physical_type=2, viscosity=0.7)   # Stokes: no linearisation state needed
    _, otw = time_weights()
    L = orc.OracleNavierStokesMatrix(mesh, oprm, otw, pressure_average_fix=False)
    G = orc.OracleNavierStokesMatrix(gmesh, oprm, otw, pressure_average_fix=False)
    # ownership: every global node is owned by exactly one rank
    owned = torch.zeros(gmesh.u.n_owned, dtype=torch.int32)
    owned[torch.from_numpy(mesh.u.global_ids[:mesh.u.n_owned])] += 1
    dist.all_reduce(owned)
    assert int(owned.min()) == 1 and int(owned.max()) == 1
    su, sp = seeded_uniform(mesh.u.global_ids, 3, 31), seeded_uniform(mesh.p.global_ids, 1, 32)
    su[mesh.u.n_owned * 3:] = 77.0      # ghosts must come from the owner
    sp[mesh.p.n_owned:] = 77.0
    update_ghost_values(mesh.u, su)
    update_ghost_values(mesh.p, sp)
    du, dp = np.zeros(L.n_u), np.zeros(L.n_p)
    L._loop(orc.OP_VMULT, su, sp, du, dp)
    compress_add(mesh.u, du)
    compress_add(mesh.p, dp)
    gu, gp = np.zeros(G.n_u), np.zeros(G.n_p)
    G._loop(orc.OP_VMULT, seeded_uniform(gmesh.u.global_ids, 3, 31), seeded_uniform(gmesh.p.global_ids, 1, 32), gu, gp)
    ref_u = gu.reshape(-1, 3)[mesh.u.global_ids[:mesh.u.n_owned]]
    ref_p = gp[mesh.p.global_ids[:mesh.p.n_owned]]
    eu = np.abs(du.reshape(-1, 3)[:mesh.u.n_owned] - ref_u).max() / np.abs(gu).max()
    ep = np.abs(dp[:mesh.p.n_owned] - ref_p).max() / np.abs(gp).max()
    print(f"[rank {rank}] partitioned oracle vs global oracle: {eu:.2e} {ep:.2e}", flush=True)
    ok = eu < 1e-13 and ep < 1e-13
    flag = torch.tensor([0 if ok else 1])
    dist.all_reduce(flag)
    dist.destroy_process_group()
    sys.exit(int(flag.item() != 0))


if __name__ == "__main__":
    main()
