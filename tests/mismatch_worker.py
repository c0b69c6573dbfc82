"""Worker of the rank-mismatch test (torchrun, 2 ranks): rank 1 issues one collective-bearing vmult fewer than rank 0.
MatrixFree::cell_loop is collective (navier_stokes_matrix.cc:236-242), so this is a caller bug; the library's stall watchdog
(csrc/comm.cu stream_wait) must turn the resulting hang into ADAFLO_B200_ERR_COMM on the waiting rank.  Exit code 0 = rank 0 saw
the error within the time-out and rank 1 finished normally."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
os.environ["ADAFLO_B200_COMM_TIMEOUT"] = "3"

from adaflo_b200 import capi  # noqa: E402
from adaflo_b200.mesh import BoxMesh  # noqa: E402
from adaflo_b200.operators import BlockVector, Context, MatrixFree, NavierStokesMatrix  # noqa: E402
from helpers import flow_params, time_weights  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
    dist.init_process_group("gloo")            # out-of-band plumbing only: the library owns its NCCL communicator
    ctx = Context(torch.cuda.current_device())

    def bcast(b):
        obj = [b]
        dist.broadcast_object_list(obj, src=0)
        return obj[0]
    ctx.init_comm(rank, world, bcast)
    mesh = BoxMesh((4, 4, 2 * world), rank=rank, nranks=world)
    mf = MatrixFree(ctx, mesh, "cartesian")
    dprm, _ = flow_params()
    dtw, _ = time_weights()
    old, old_old = BlockVector.like(mf), BlockVector.like(mf)
    A = NavierStokesMatrix(dprm, 0, 1, 0, 1, old, old_old)
    A.initialize(mf, dtw, True)
    src, dst = BlockVector.like(mf), BlockVector.like(mf)
    src.block(0).fill_(1.0)
    A.vmult(dst, src)                          # matched on both ranks
    ctx.synchronize()
    dist.barrier()
    ok = True
    if rank == 0:
        A.vmult(dst, src)                      # rank 1 never issues this one
        t0 = time.time()
        try:
            ctx.synchronize()
            ok = False
            print("rank 0: synchronize returned although the peer never arrived", flush=True)
        except capi.AdafloB200Error as e:
            waited = time.time() - t0
            ok = e.code == capi.ERR_COMM and 2.0 < waited < 30.0
            print(f"rank 0: {e} (after {waited:.1f} s)", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    # the aborted communicator must not be used again; leave without the regular teardown of the library objects
    os._exit(0 if ok else 1)


if __name__ == "__main__":
    main()
