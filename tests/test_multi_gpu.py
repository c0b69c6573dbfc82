"""Multi-GPU parity (needs >= 2 GPUs): z-slab partition, NCCL ghost exchange inside the library."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("overlap", ["1", "0"])
@pytest.mark.parametrize("world", [2, 4])
def test_partitioned_vmult_matches_global_oracle(world, overlap):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + world), os.path.join(HERE, "mgpu_worker.py")]
    env = dict(os.environ, ADAFLO_B200_HALO_OVERLAP=overlap)  # import overlapped with interior decades / serial
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    print(r.stdout[-3000:])
    print(r.stderr[-3000:])
    assert r.returncode == 0
