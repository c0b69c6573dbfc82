#!/usr/bin/env python
"""bench.py -- NS matrix-free vmult throughput (3-D Q2/Q1, fp64) on N B200s.

Metric (BASELINE.json): DoF/s of NavierStokesMatrix::vmult, plus % of the HBM roofline.
Workload (configs[1], "beltrami_3d scaled to a ~50M-DoF synthetic box mesh"): 128^3 cells
per GPU on [-1,1]^3, Q2/Q1, Newton linearisation stored at the quadrature points (variant
N of SURVEY 8(d)), Cartesian cells, constant coefficients, nu = rho = 1, BDF-2, dt = 0.05,
Dirichlet velocity on all faces, pressure mean fixed.  Multi-GPU is weak scaling: the box
grows in z by 128 cell layers per rank (z-slab partition, ghost exchange over NCCL).

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun)
  python bench.py --impl reference ...                     (CPU arm, see below)

One "step" = one vmult over the whole mesh.  `value` times device-resident vectors with
CUDA events; `e2e` times adaflo_b200_ns_vmult_host (pinned HOST vectors, H2D + vmult +
D2H inside the timed region).  The reference arm cannot run adaflo itself (deal.II,
p4est, Trilinos and MPI are absent): it times the CPU port of the same
operator (oracle/ns_cpu_port.c, validated against the dense oracle) with all host threads.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

B_PER_DOF_N = 125.3   # algorithmic bytes per DoF, variant N (16 vec + 5.6 idx + 103.68 lin)
METRIC = "ns_vmult_dofs_per_s"
UNIT = "DoF/s"


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of the cell kernel at the bench workload, from the committed
    summary of the latest `ncu --set full` capture (profiles/); None if the summary is missing."""
    path = os.path.join(ROOT, "profiles", "r1_g_ns_vmult_r1_key_metrics.txt")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    total, seen = 0.0, 0
    try:
        for line in open(path):
            f = line.split()
            if len(f) == 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                total += float(f[2]) * scale[f[1]]
                seen += 1
    except OSError:
        return None
    return total if seen == 2 else None


def bind_to_gpu_numa_node(device_index):
    """Pin this process to the CPUs of the NUMA node the GPU hangs off, so that the pinned host vectors of the
    end-to-end path are allocated next to its PCIe root port (matters with several ranks per box).  Best effort."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(device_index)
        bus = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_port_sample(n_xy, layers, steps, warmup):
    """Time the CPU port on a bounded sample: an n_xy x n_xy x layers slab of the same workload."""
    from adaflo_b200.mesh import BoxMesh, seeded_uniform
    from oracle import cpu_port
    mesh = BoxMesh((n_xy, n_xy, layers), upper=(1.0, 1.0, -1.0 + 2.0 * layers / n_xy), build_coords=False)
    rng = np.random.default_rng(999)
    lin = rng.uniform(-1, 1, size=(mesh.n_cells, 12, 27))
    port = cpu_port.CpuPortVmult(mesh, 1, 30.0, 1.0, 0.5, 1.0, 1.0, 0.0, 0.0, lin)
    su = seeded_uniform(mesh.u.global_ids, 3, 12345)
    sp = seeded_uniform(mesh.p.global_ids, 1, 12345)
    du, dp = np.zeros_like(su), np.zeros_like(sp)
    n_dofs = mesh.u.n_dofs_owned + mesh.p.n_dofs_owned
    threads = cpu_port.max_threads()
    times = []
    for i in range(warmup + steps):
        du[:] = 0
        dp[:] = 0
        t0 = time.perf_counter()
        port.cells(su, sp, du, dp)
        t1 = time.perf_counter()
        if i >= warmup:
            times.append(t1 - t0)
    t = statistics.mean(times)
    return {"value": n_dofs / t, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{n_xy}x{n_xy}x{layers} cells slab of the 128^3 workload ({n_dofs} DoFs), Newton/Cartesian/constant coefficients, "
                      f"{steps} vmult cell loops, mean {t * 1e3:.1f} ms each; oracle/ns_cpu_port.c (restatement, not adaflo+deal.II)"}, t


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.n
    layers = max(2, min(n, args.cpu_layers))
    cb, t = cpu_port_sample(n, layers, args.steps, min(args.warmup, 2))
    line = {"metric": METRIC, "value": cb["value"], "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"ns_vmult Q2/Q1 3D box {n}^3 cells/GPU, Newton stored linearisation, Cartesian (CPU port timed on a bounded slab sample)"},
            "cpu_baseline": cb, "gpu_launches": 0,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)  # 10 warm-up + >= 50 timed vmults
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=128, help="cells per direction per GPU")
    ap.add_argument("--cpu-layers", type=int, default=32, help="z layers of the CPU baseline sample (about 10-30 s of CPU work)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--preroll", type=float, default=0.7, help="seconds of untimed load before the timed steps (clock sampling)")
    ap.add_argument("--no-nonlinear-step", action="store_true")
    ap.add_argument("--kernel-only", action="store_true", help="developer A/B mode: kernel timings + checksum only (not a bench line)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from adaflo_b200 import capi
    from adaflo_b200.mesh import BoxMesh, beltrami_velocity, seeded_uniform
    from adaflo_b200.operators import BlockVector, Context, MatrixFree, NavierStokesMatrix

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    numa_node = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = Context(local_rank)
    if world > 1:
        def bcast(b):
            obj = [b]
            dist.broadcast_object_list(obj, src=0)
            return obj[0]
        ctx.init_comm(rank, world, bcast)

    n = args.n
    t_setup = time.perf_counter()
    # weak scaling: n x n x (n * world) cells on [-1,1]^2 x [-1, -1 + 2 world]
    mesh = BoxMesh((n, n, n * world), upper=(1.0, 1.0, -1.0 + 2.0 * world), rank=rank, nranks=world)
    mf = MatrixFree(ctx, mesh, "cartesian")
    prm = capi.FlowParams(capi.INCOMPRESSIBLE, capi.NEWTON, 0, 0, 1.0, 1.0, 0.0, 0.0, 0.0, 0.0, 0.5)
    dt = 0.05
    tw = capi.TimeWeights(1.5 / dt, -2.0 / dt, 0.5 / dt, 1.0, 2.0, -1.0)   # BDF-2, tests/beltrami_3d.prm
    sol, old, old_old = BlockVector.like(mf), BlockVector.like(mf), BlockVector.like(mf)
    A = NavierStokesMatrix(prm, 0, 1, 0, 1, old, old_old)
    A.initialize(mf, tw, True)

    def dev(a):
        return torch.from_numpy(np.ascontiguousarray(a)).to(ctx.device)

    # linearisation point: Beltrami field at t = 0 (tests/beltrami.cc:95-107), produced by the residual kernel
    u0 = beltrami_velocity(mesh.u.coords, 0.0).reshape(-1)
    sol.block(0).copy_(dev(u0))
    old.block(0).copy_(sol.block(0))
    old_old.block(0).copy_(sol.block(0))
    del u0
    mesh.u.coords = None
    rhs, zero = BlockVector.like(mf), BlockVector.like(mf)
    A.residual(rhs, sol, zero)
    del rhs, zero
    src = BlockVector(dev(seeded_uniform(mesh.u.global_ids, 3, 12345)), dev(seeded_uniform(mesh.p.global_ids, 1, 12345)))
    dst = BlockVector.like(mf)
    n_dofs_rank = mf.n_owned(0) + mf.n_owned(1)
    n_dofs = n_dofs_rank
    if world > 1:
        t = torch.tensor([n_dofs_rank], dtype=torch.int64, device=ctx.device)
        dist.all_reduce(t)
        n_dofs = int(t.item())
    ctx.synchronize()
    t_setup = time.perf_counter() - t_setup

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_steps():
        A.get_kernel_timing()
        A.set_kernel_timing(True)
        l0 = ctx.lib.adaflo_b200_kernel_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(args.steps):
            A.vmult(dst, src)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        nl_, ks_ = A.get_kernel_timing()
        A.set_kernel_timing(False)
        return ms, ks_ / max(nl_, 1), ctx.lib.adaflo_b200_kernel_launch_count() - l0

    for _ in range(args.warmup):
        A.vmult(dst, src)
    # (1) burst: K steps on a GPU that has not reached its power limit yet (reported as "burst" only)
    b_ms, b_k, _ = timed_steps()
    # (2) sustained: the K timed steps last only tens of milliseconds, so the GPU is first kept under the same
    # load for --preroll seconds (untimed, additional warm-up); nvidia-smi samples clocks and throttle reasons
    # over pre-roll + timed steps.  This is the headline number: a Krylov solver runs vmult back to back.
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    torch.cuda.synchronize()
    t_pre = time.perf_counter()
    while time.perf_counter() - t_pre < args.preroll:
        for _ in range(10):
            A.vmult(dst, src)
        torch.cuda.synchronize()
    ms_total, k_avg_local, launches = timed_steps()
    clocks = sampler.stop() if rank == 0 else None
    k_launches, k_seconds = 1, k_avg_local
    tt = torch.tensor([ms_total, k_seconds / max(k_launches, 1)], dtype=torch.float64, device=ctx.device)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_total, k_avg = float(tt[0]), float(tt[1])
    ms_per_step = ms_total / args.steps
    value = n_dofs / (ms_per_step * 1e-3)

    if args.kernel_only and os.environ.get("ADAFLO_B200_HOST_CHUNKS"):
        nu_own, np_own = mf.n_owned(0), mf.n_owned(1)
        hu = src.block(0)[:nu_own].cpu().pin_memory()
        hp = src.block(1)[:np_own].cpu().pin_memory()
        ou, op = torch.empty_like(hu).pin_memory(), torch.empty_like(hp).pin_memory()
        A.vmult_host(ou, op, hu, hp)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5):
            A.vmult_host(ou, op, hu, hp)
        torch.cuda.synchronize()
        print("host chunks", os.environ["ADAFLO_B200_HOST_CHUNKS"], "e2e ms", (time.perf_counter() - t0) / 5 * 1e3)
    if args.kernel_only:
        if rank == 0:
            print(json.dumps({"lib": os.environ.get("ADAFLO_B200_LIB", "default"), "burst_kernel_ms": b_k * 1e3, "burst_ms_per_step": b_ms / args.steps,
                              "kernel_ms": k_avg * 1e3, "ms_per_step": ms_per_step, "clocks": clocks,
                              "checksum": [float(dst.block(0).norm()), float(dst.block(1).norm())]}))
        return
    # ---- e2e: HOST (pinned) vectors through adaflo_b200_ns_vmult_host
    nu_own, np_own = mf.n_owned(0), mf.n_owned(1)
    hu = src.block(0)[:nu_own].cpu().pin_memory()      # host vectors hold the locally owned range
    hp = src.block(1)[:np_own].cpu().pin_memory()
    ou, op = torch.empty_like(hu).pin_memory(), torch.empty_like(hp).pin_memory()
    A.vmult_host(ou, op, hu, hp)   # warm-up (allocates staging buffers)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        A.vmult_host(ou, op, hu, hp)
    barrier()
    e2e_t = (time.perf_counter() - t0) / args.e2e_steps
    te = torch.tensor([e2e_t], dtype=torch.float64, device=ctx.device)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_t = float(te[0])
    # parity of the two paths (cheap sanity, not timed)
    err = float((ou.to(ctx.device) - dst.block(0)[:nu_own]).norm() / dst.block(0)[:nu_own].norm())

    # ---- time per nonlinear step (north_star / SURVEY 8(d)): residual + k GMRES iterations per Newton step with
    # the golden iteration counts of tests/beltrami_3d.output:13-16, no preconditioner (the reference's Trilinos
    # ILU/AMG is out of scope), everything device resident
    nl = None
    if not args.no_nonlinear_step:
        from adaflo_b200.solvers import SolverControl, SolverGMRES, NoConvergence
        counts = [30, 30, 30, 23]
        rhs, zero = BlockVector.like(mf), BlockVector.like(mf)
        upd = [mf.initialize_dof_vector(0), mf.initialize_dof_vector(1)]
        def newton_steps(its):
            for k in its:
                A.residual(rhs, sol, zero)
                for v in upd:
                    v.zero_()
                ctl = SolverControl(k, 1e-300, 0.0)
                try:
                    SolverGMRES(ctx, ctl, 50).solve(lambda d, s_: A.vmult(BlockVector(*d), BlockVector(*s_)), upd, [rhs.block(0), rhs.block(1)],
                                                    None, [nu_own, np_own])
                except NoConvergence:
                    pass

        newton_steps([30])     # untimed warm-up: the 31 basis vectors (13 GB) come from cudaMalloc the first time
        barrier()
        t0 = time.perf_counter()
        newton_steps(counts)
        barrier()
        nl_t = time.perf_counter() - t0
        tn = torch.tensor([nl_t], dtype=torch.float64, device=ctx.device)
        if world > 1:
            dist.all_reduce(tn, op=dist.ReduceOp.MAX)
        nl = {"ms": float(tn[0]) * 1e3, "newton_iterations": len(counts), "gmres_iterations": counts,
              "note": "residual + unpreconditioned GMRES(50) with the golden iteration counts of tests/beltrami_3d.output; wall clock"}
        del rhs, zero, upd

    if rank == 0:
        peak, which = measured_peak()
        alg_bytes = B_PER_DOF_N * n_dofs_rank
        achieved = alg_bytes / k_avg / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"ns_vmult Q2/Q1 3D box {n}^3 cells/GPU ({n_dofs_rank} DoFs/GPU, {n_dofs} total), Newton stored linearisation, "
                                   "Cartesian, constant coefficients, Dirichlet u on all faces, pressure mean fixed",
                       "l2": "inputs larger than L2 (vectors 0.42 GB, linearisation data 5.4 GB per GPU)",
                       "partition": f"z-slabs x{world}", "setup_s": round(t_setup, 1), "numa_node_rank0": numa_node},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic() if n == 128 and world == 1 else None,
                         "traffic_source": "profiles/r1_g_ns_vmult_r1_key_metrics.txt (ncu --set full, same workload, 1 GPU)", "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({which})",
                         "kernel": "ns_vmult_r1_kernel<with_p, newton, cartesian> (persistent, 148 blocks x 11 warps)", "kernel_ms": k_avg * 1e3,
                         "algorithmic_bytes_per_launch": alg_bytes, "bytes_per_dof": B_PER_DOF_N},
            "e2e": {"value": n_dofs / e2e_t, "unit": UNIT, "h2d_bytes_per_step": 8 * n_dofs, "d2h_bytes_per_step": 8 * n_dofs,   # all ranks
                    "ms_per_step": e2e_t * 1e3, "rel_diff_vs_device_path": err},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "nonlinear_step": nl,
            "burst": {"ms_per_step": b_ms / args.steps, "kernel_ms": b_k * 1e3, "value": n_dofs / (b_ms / args.steps * 1e-3),
                      "roofline_frac": B_PER_DOF_N * n_dofs_rank / b_k / 1e9 / peak,
                      "note": "same K steps right after warm-up, before the GPU reaches its power limit (rank 0 timing)"},
        }
        if not args.no_cpu_baseline and world >= 1:
            cb, _ = cpu_port_sample(n, max(2, min(n, args.cpu_layers)), 10, 2)
            line["cpu_baseline"] = cb
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
