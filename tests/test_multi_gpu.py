"""Multi-GPU parity (needs >= 2 GPUs): z-slab partition, NCCL ghost exchange inside the library."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("overlap", ["1", "0", "serial"])
@pytest.mark.parametrize("world", [2, 4])
def test_partitioned_vmult_matches_global_oracle(world, overlap):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + world), os.path.join(HERE, "mgpu_worker.py")]
    # "1": import overlapped with interior decades (two launches); "0": the default, import next to the zero-fill of dst
    # on the communication stream; "serial": everything on the compute stream
    env = dict(os.environ, ADAFLO_B200_HALO_OVERLAP="1" if overlap == "1" else "0",
               ADAFLO_B200_FILL_OVERLAP="0" if overlap == "serial" else "1")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    print(r.stdout[-3000:])
    print(r.stderr[-3000:])
    assert r.returncode == 0


def test_rank_mismatch_fails_loudly_instead_of_hanging():
    """one rank issues a collective-bearing vmult its peer never issues: the stall watchdog of the library (csrc/comm.cu
    stream_wait) aborts the communicator and returns ADAFLO_B200_ERR_COMM within the time-out"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
           "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.join(HERE, "mismatch_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    print(r.stdout[-3000:])
    print(r.stderr[-3000:])
    assert r.returncode == 0
