"""Host-side multi-rank logic on CPU (gloo backend, world_size 2) and mesh bookkeeping."""
import os
import subprocess
import sys

import numpy as np

from adaflo_b200.mesh import BoxMesh, FLAG_SHIFT, NODE_MASK

HERE = os.path.dirname(os.path.abspath(__file__))


def test_gloo_two_ranks_partitioned_cell_loop():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(HERE, "gloo_worker.py")]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    print(r.stdout[-2000:], r.stderr[-2000:])
    assert r.returncode == 0


def test_dof_counts_of_the_benchmark_sizes():
    # n = 32 -> 0.86 M dofs, (u = 3 (2n+1)^3, p = (n+1)^3)
    m = BoxMesh((8, 8, 8), build_coords=False)
    assert m.u.n_dofs_owned == 3 * 17 ** 3 and m.p.n_dofs_owned == 9 ** 3
    assert m.n_cells == 512


def test_slab_partition_covers_all_cells_and_nodes():
    n, world = (3, 4, 7), 3
    meshes = [BoxMesh(n, rank=r, nranks=world, build_coords=False) for r in range(world)]
    assert sum(m.n_cells for m in meshes) == 3 * 4 * 7
    assert sum(m.u.n_owned for m in meshes) == 7 * 9 * 15
    assert sum(m.p.n_owned for m in meshes) == 4 * 5 * 8
    for r, m in enumerate(meshes):
        # ghosts: one node plane below, none above (interface owned by the lower rank)
        assert m.u.n_ghost == (7 * 9 if r > 0 else 0)
        assert (len(m.u.send) == 1) == (r < world - 1)
        if r > 0:
            peer, first, cnt = m.u.recv[0]
            assert peer == r - 1 and first == m.u.n_owned and cnt == m.u.n_ghost
            # what the lower rank sends are exactly my ghosts, in the same order
            send_ids = meshes[r - 1].u.global_ids[meshes[r - 1].u.send[0][1]]
            assert np.array_equal(send_ids, m.u.global_ids[first:first + cnt])


def test_flag_encoding_roundtrip():
    m = BoxMesh((2, 2, 2), build_coords=False)
    enc = m.u.encoded_cell_nodes()
    assert np.array_equal(enc & NODE_MASK, m.u.cell_nodes)
    flags = enc >> FLAG_SHIFT
    # all boundary nodes are Dirichlet in all 3 components, the single interior-of-domain nodes are free
    assert set(np.unique(flags)) == {0, 7}
    assert (flags == 0).sum() > 0
    # pressure has no Dirichlet constraint
    assert (m.p.encoded_cell_nodes() >> FLAG_SHIFT).max() == 0


def test_augmented_pressure_space_layout():
    m = BoxMesh((2, 3, 2), augmented=True, build_coords=False)
    assert m.p.nodes_per_cell == 9 and m.p.n_owned == 3 * 4 * 3 + m.n_cells
    assert np.array_equal(m.p.cell_nodes[:, 8], m.p.n_q1_owned + np.arange(m.n_cells))
