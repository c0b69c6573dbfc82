"""The C++ mirror class (include/adaflo_b200/navier_stokes_matrix.h) driven like the reference
drives NavierStokesMatrix, checked against the CPU oracle (GPU test; the driver binary is built
by __graft_entry__.build())."""
import os
import struct
import subprocess

import numpy as np
import pytest

from adaflo_b200 import capi
from adaflo_b200.mesh import BoxMesh, beltrami_velocity, seeded_uniform
from oracle import oracle as orc

from helpers import flow_params, rel_err, time_weights

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "tests", "cpp", "ns_matrix_driver")


def write_dump(path, arrays):
    with open(path, "wb") as f:
        for name, a in arrays.items():
            a = np.ascontiguousarray(a)
            f.write(struct.pack("<I", len(name)))
            f.write(name.encode())
            f.write(struct.pack("<Q", a.nbytes))
            f.write(a.tobytes())


def read_dump(path):
    out = {}
    with open(path, "rb") as f:
        while True:
            h = f.read(4)
            if not h:
                break
            n = struct.unpack("<I", h)[0]
            name = f.read(n).decode()
            nbytes = struct.unpack("<Q", f.read(8))[0]
            out[name] = np.frombuffer(f.read(nbytes), dtype=np.float64)
    return out


def test_cpp_navier_stokes_matrix_matches_oracle(tmp_path):
    assert os.path.exists(DRIVER), "build the driver with __graft_entry__.build()"
    mesh = BoxMesh((5, 4, 3))
    _, oprm = flow_params(damping=-0.2, tau_grad_div=0.3, viscosity=0.7, density=1.3)
    _, otw = time_weights()
    O = orc.OracleNavierStokesMatrix(mesh, oprm, otw)
    u = beltrami_velocity(mesh.u.coords, 0.1).reshape(-1) + 0.1 * seeded_uniform(mesh.u.global_ids, 3, 7)
    p = seeded_uniform(mesh.p.global_ids, 1, 8)
    uo, uoo = seeded_uniform(mesh.u.global_ids, 3, 9), seeded_uniform(mesh.u.global_ids, 3, 10)
    su, sp = seeded_uniform(mesh.u.global_ids, 3, 11), seeded_uniform(mesh.p.global_ids, 1, 12)
    table, row = mesh.cartesian_geometry()
    con = mesh.u.constrained_dofs()
    write_dump(tmp_path / "in.bin", {
        "meta": np.array([mesh.n_cells, mesh.u.n_owned, mesh.p.n_owned, len(con)], dtype=np.uint32),
        "params": np.array([oprm.viscosity, oprm.density, oprm.damping, oprm.tau_grad_div, oprm.beta, otw.weight, otw.weight_old,
                            otw.weight_old_old, otw.tau1, otw.extrapolate_old, otw.extrapolate_old_old]),
        "cell_u": mesh.u.encoded_cell_nodes(), "cell_p": mesh.p.encoded_cell_nodes(), "constrained_u": con,
        "geom_table": table, "geom_row": row, "solution_u": u, "solution_p": p, "old_u": uo, "old_old_u": uoo,
        "src_u": su, "src_p": sp})
    r = subprocess.run([DRIVER, str(tmp_path / "in.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr)
    assert r.returncode == 0
    got = read_dump(tmp_path / "out.bin")
    O.old_u, O.old_old_u = uo, uoo
    ru, rp = np.zeros(O.n_u), np.zeros(O.n_p)
    O.residual(ru, rp, u, p, np.zeros(O.n_u), np.zeros(O.n_p))
    O.apply_pressure_average_projection(rp)
    ou, op = O.vmult(su, sp)
    assert rel_err(got["residual_u"], ru) < 1e-12 and rel_err(got["residual_p"], rp) < 1e-12
    assert rel_err(got["vmult_u"], ou) < 1e-12 and rel_err(got["vmult_p"], op) < 1e-12
    assert abs(got["scalars"][0] - (ou @ su + op @ sp)) < 1e-10 * abs(ou @ su + op @ sp)
    assert got["scalars"][1] == 1
    # begin_linearized_velocities(macro_cell): [batch][q][row][lane] with 4 lanes vs the oracle's [cell][row][q]
    nb = (mesh.n_cells + 3) // 4
    lin = got["lin_canonical"].reshape(nb, 27, 12, 4)
    ref = np.zeros((nb * 4, 12, 27))
    ref[:mesh.n_cells] = O.lin
    assert rel_err(lin, ref.reshape(nb, 4, 12, 27).transpose(0, 3, 2, 1)) < 1e-12


def test_cpp_level_set_classes_match_oracle(tmp_path):
    """include/adaflo_b200/level_set_okz.h: LevelSetOKZSolver{AdvanceConcentration, ComputeNormal, ComputeCurvature,
    Reinitialization} driven from C++ (tests/cpp/level_set_driver.cpp)"""
    driver = os.path.join(ROOT, "tests", "cpp", "level_set_driver")
    assert os.path.exists(driver), "build the driver with __graft_entry__.build()"
    n_sub = 3
    mesh = BoxMesh((4, 3, 3), lower=(-1, -1, -1), upper=(1.0, 0.7, 0.5))
    mesh.build_ls_space(n_sub)
    h_min = float(mesh.h.min())
    eps, eps_used, dt = 1.5, 1.5 * h_min / n_sub, 0.02
    O = orc.OracleLevelSet(mesh, orc.LSParams(eps, eps_used, 4.0, h_min, dt, 3, 0))
    O.diag = O.mass_diagonal()
    x = mesh.ls.coords
    c = -np.tanh((np.linalg.norm(x - np.array([0.1, -0.05, -0.2]), axis=1) - 0.45) / 0.24)
    vec = lambda s: seeded_uniform(mesh.ls.global_ids, 1, s)
    c_old, c_oo, src = c + 0.05 * vec(7), c + 0.05 * vec(8), vec(10)
    vel = beltrami_velocity(mesh.u.coords, 0.0).reshape(-1) + 0.1 * seeded_uniform(mesh.u.global_ids, 3, 6)
    table, row = mesh.cartesian_geometry()
    write_dump(tmp_path / "in.bin", {
        "meta": np.array([mesh.n_cells, mesh.u.n_owned, mesh.p.n_owned, mesh.ls.n_owned, n_sub], dtype=np.uint32),
        "params": np.array([eps, eps_used, 4.0, h_min, 75.0, -100.0, 25.0, dt]),
        "cell_u": mesh.u.encoded_cell_nodes(), "cell_p": mesh.p.encoded_cell_nodes(), "cell_ls": mesh.ls.encoded_cell_nodes(),
        "geom_table": table, "geom_row": row, "cell_diameters": mesh.cell_diameters(),
        "c": c, "c_old": c_old, "c_old_old": c_oo, "src": src, "vel": vel})
    r = subprocess.run([driver, str(tmp_path / "in.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr)
    assert r.returncode == 0
    got = read_dump(tmp_path / "out.bin")
    assert rel_err(got["mass_diagonal"], O.diag) < 1e-12
    o = np.zeros(O.n)
    O.advance_concentration_rhs(o, vel, c, c_old, c_oo, 75.0, -100.0, 25.0)
    assert rel_err(got["advance_rhs"], o) < 1e-12
    assert rel_err(got["advance_vmult"], O.advance_concentration_vmult(src, 75.0)) < 1e-12
    nr = [np.zeros(O.n) for _ in range(3)]
    O.compute_normal_rhs(nr, c)
    assert rel_err(got["normal_rhs_0"], nr[0]) < 1e-12 and rel_err(got["normal_rhs_2"], nr[2]) < 1e-12
    assert rel_err(got["curvature_vmult"], O.compute_curvature_vmult(src, True)) < 1e-12
    assert rel_err(got["reinit_vmult_diffuse"], O.reinitialization_vmult(src, True)) < 1e-12

    # ---- the reference's constructor contracts + solver-level routines, checked against dense solves with the oracle's operators
    def dense(apply):
        M = np.zeros((O.n, O.n))
        e = np.zeros(O.n)
        for j in range(O.n):
            e[j] = 1.0
            M[:, j] = apply(e)
            e[j] = 0.0
        return M
    counts = got["counts"].astype(int)
    assert len(counts) == 6 and all(1 <= k < 200 for k in counts), counts
    N = dense(lambda v: O._scalar(0, 4.0, v))                       # one block of compute_normal_vmult = the projection matrix
    normal = [np.linalg.solve(N, nr[d]) for d in range(3)]
    for d in range(3):
        assert rel_err(got[f"solved_normal_{d}"], normal[d]) < 1e-5          # CG stops at 1e-7 relative residual
    rk = np.zeros(O.n)
    O.compute_curvature_rhs(rk, [got[f"solved_normal_{d}"].copy() for d in range(3)])
    kappa = np.linalg.solve(N, rk)
    assert rel_err(got["solved_curvature"], kappa) < 1e-5
    # advection, first BDF step: (gamma M + C(u)) dc = rhs with gamma = 1 / dt, c += dc
    A_adv = dense(lambda v: O.advance_concentration_vmult(v, 1.0 / dt))
    ra = np.zeros(O.n)
    O.advance_concentration_rhs(ra, vel, c, c_old, c_oo, 1.0 / dt, -1.0 / dt, 0.0)
    assert rel_err(got["advected"], c + np.linalg.solve(A_adv, ra)) < 1e-6
    # reinitialisation: 1 diffusive + 2 compressive pseudo time steps, normal recomputed (fast) at the first compressive one
    cr = c.copy()
    O2 = orc.OracleLevelSet(mesh, orc.LSParams(eps, eps_used, 4.0, h_min, dt, 3, 0))
    O2.diag = O.diag
    for tau in range(3):
        diffuse, first = tau < 1, tau == 1
        if first:
            nrr = [np.zeros(O.n) for _ in range(3)]
            O2.compute_normal_rhs(nrr, cr)
            nfast = [np.linalg.solve(N, nrr[d]) for d in range(3)]
        rr = np.zeros(O.n)
        O2.reinitialization_rhs(rr, cr, None if diffuse else nfast, diffuse, first)
        R = dense(lambda v: O2.reinitialization_vmult(v, diffuse))
        cr = cr + np.linalg.solve(R, rr)
    assert rel_err(got["reinitialized"], cr) < 1e-3                        # fast normal: 1e-5, CG: 1e-6 per step, three nonlinear steps


def test_cpp_matrix_free_adaptor_with_renumbered_velocity(tmp_path):
    """include/adaflo_b200/matrix_free_adaptor.h fed by a mock of dealii::MatrixFree's batch / lane
    storage with a velocity numbering scrambled like DoFRenumbering::Cuthill_McKee leaves it (source/navier_stokes.cc:208-209):
    host vectors in that numbering go through vmult_host / velocity_vmult_host and come back in it"""
    driver = os.path.join(ROOT, "tests", "cpp", "adaptor_driver")
    assert os.path.exists(driver), "build the driver with __graft_entry__.build()"
    mesh = BoxMesh((5, 4, 3))
    _, oprm = flow_params(damping=-0.2, tau_grad_div=0.3, viscosity=0.7, density=1.3)
    _, otw = time_weights()
    O = orc.OracleNavierStokesMatrix(mesh, oprm, otw)
    u = beltrami_velocity(mesh.u.coords, 0.1).reshape(-1) + 0.1 * seeded_uniform(mesh.u.global_ids, 3, 7)
    p = seeded_uniform(mesh.p.global_ids, 1, 8)
    uo, uoo = seeded_uniform(mesh.u.global_ids, 3, 9), seeded_uniform(mesh.u.global_ids, 3, 10)
    su, sp = seeded_uniform(mesh.u.global_ids, 3, 11), seeded_uniform(mesh.p.global_ids, 1, 12)
    rng = np.random.default_rng(5)
    nu, np_ = mesh.u.n_owned * 3, mesh.p.n_owned
    perm_u = rng.permutation(nu)           # host position of the node-major entry 3 node + c
    perm_p = rng.permutation(np_)
    cell_u = mesh.u.cell_nodes.astype(np.int64)                     # [cell][27] node ids
    dofs_u = np.stack([perm_u[3 * cell_u + c] for c in range(3)], axis=1).reshape(mesh.n_cells, 81)   # component-blocked per cell
    dofs_p = perm_p[mesh.p.cell_nodes.astype(np.int64)]
    con = np.sort(perm_u[mesh.u.constrained_dofs().astype(np.int64)])

    def to_host(v, perm):
        h = np.empty_like(v)
        h[perm] = v
        return h
    inv_h = np.tile(1.0 / mesh.h, (mesh.n_cells, 1)) if np.ndim(mesh.h) == 1 else 1.0 / mesh.h
    write_dump(tmp_path / "in.bin", {
        "meta": np.array([mesh.n_cells, nu, np_, len(con), 4], dtype=np.uint32),
        "params": np.array([oprm.viscosity, oprm.density, oprm.damping, oprm.tau_grad_div, oprm.beta, otw.weight, otw.weight_old,
                            otw.weight_old_old, otw.tau1, otw.extrapolate_old, otw.extrapolate_old_old]),
        "plain_dof_indices_u": dofs_u.astype(np.uint32), "plain_dof_indices_p": dofs_p.astype(np.uint32), "constrained_u": con.astype(np.uint32),
        "inverse_jacobian_diagonal": np.ascontiguousarray(inv_h, dtype=np.float64),
        "solution_u": to_host(u, perm_u), "solution_p": to_host(p, perm_p), "old_u": to_host(uo, perm_u), "old_old_u": to_host(uoo, perm_u),
        "src_u": to_host(su, perm_u), "src_p": to_host(sp, perm_p)})
    r = subprocess.run([driver, str(tmp_path / "in.bin"), str(tmp_path / "out.bin")], capture_output=True, text=True, timeout=300)
    print(r.stdout, r.stderr)
    assert r.returncode == 0
    got = read_dump(tmp_path / "out.bin")
    O.old_u, O.old_old_u = uo, uoo
    ru, rp = np.zeros(O.n_u), np.zeros(O.n_p)
    O.residual(ru, rp, u, p, np.zeros(O.n_u), np.zeros(O.n_p))
    ou, op = O.vmult(su, sp)
    assert rel_err(got["residual_u"], to_host(ru, perm_u)) < 1e-12
    assert rel_err(got["vmult_u"], to_host(ou, perm_u)) < 1e-12 and rel_err(got["vmult_p"], to_host(op, perm_p)) < 1e-12
    O.fix_linearization_point()
    assert rel_err(got["velocity_vmult"], to_host(O.velocity_vmult(su), perm_u)) < 1e-12
