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


TRAFFIC_FILES = ("r2_j_ns_key_metrics.txt", "r2_i_ns_key_metrics.txt", "r2_f_ns_key_metrics.txt", "r1_g_ns_vmult_r1_key_metrics.txt")


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum of the cell kernel at the bench workload, from the committed
    summary of the latest `ncu --set full` capture (profiles/); None if the summary is missing."""
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    for name in TRAFFIC_FILES:  # newest capture first
        total, seen = 0.0, 0
        try:
            for line in open(os.path.join(ROOT, "profiles", name)):
                f = line.split()
                if line.startswith("==") and seen:
                    break  # first launch of the file only
                if len(f) == 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    total += float(f[2]) * scale[f[1]]
                    seen += 1
        except OSError:
            continue
        if seen == 2:
            return total, name
    return None, None


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
    try:  # sysfs has no NUMA node for the device (virtualised boxes): ask NVML for the CPUs next to the GPU
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus and len(cpus) < len(os.sched_getaffinity(0)):
            os.sched_setaffinity(0, cpus)
            return f"nvml:{len(cpus)} cpus"
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
    cpu_port.use_all_cores()   # explicit: torchrun exports OMP_NUM_THREADS=1 to its workers
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
                      f"{steps} vmult cell loops, mean {t * 1e3:.1f} ms each; oracle/ns_cpu_port.c (restatement, not adaflo+deal.II: sum factorisation, "
                      f"SIMD over batches of 4 cells like VectorizedArray, AVX2 + FMA, OpenMP over 4-coloured cell rows)"}, t


def parity_vs_single_gpu(ctx, rank, world, local_rank):
    """Cross-N parity inside the bench (the driver's GPU-test box has one GPU): residual + vmult of one small box on the
    N-rank partition against the SAME global problem applied by a single-rank operator on rank 0's GPU; inputs are seeded
    by global DoF id.  Returns the largest relative l2 difference over the blocks (all ranks)."""
    import torch
    import torch.distributed as dist
    from adaflo_b200 import capi
    from adaflo_b200.mesh import BoxMesh, beltrami_velocity, seeded_uniform
    from adaflo_b200.operators import BlockVector, Context, MatrixFree, NavierStokesMatrix
    n = (12, 10, 6 * world + 1)
    prm = capi.FlowParams(capi.INCOMPRESSIBLE, capi.NEWTON, 0, 0, 1.0, 1.0, 0.0, 0.0, 0.0, 0.0, 0.5)
    tw = capi.TimeWeights(30.0, -40.0, 10.0, 1.0, 2.0, -1.0)

    def dev(a, c):
        return torch.from_numpy(np.ascontiguousarray(a)).to(c.device)

    def apply(c, mesh):
        mf = MatrixFree(c, mesh, "cartesian")
        sol, old, old_old = BlockVector.like(mf), BlockVector.like(mf), BlockVector.like(mf)
        A = NavierStokesMatrix(prm, 0, 1, 0, 1, old, old_old)
        A.initialize(mf, tw, True)
        u0 = beltrami_velocity(mesh.u.coords, 0.0).reshape(-1)
        for v in (sol, old, old_old):
            v.block(0).copy_(dev(u0, c))
        rhs, zero, dst = BlockVector.like(mf), BlockVector.like(mf), BlockVector.like(mf)
        A.residual(rhs, sol, zero)
        src = BlockVector(dev(seeded_uniform(mesh.u.global_ids, 3, 12345), c), dev(seeded_uniform(mesh.p.global_ids, 1, 12345), c))
        A.vmult(dst, src)
        out = [rhs.block(0).clone(), rhs.block(1).clone(), dst.block(0).clone(), dst.block(1).clone()]
        c.synchronize()
        A.clear()
        mf.close()
        return out

    mesh = BoxMesh(n, rank=rank, nranks=world)
    mine = apply(ctx, mesh)
    gmesh = BoxMesh(n)
    sizes = [gmesh.u.n_local * 3, gmesh.p.n_local, gmesh.u.n_local * 3, gmesh.p.n_local]
    if rank == 0:
        serial = Context(local_rank)
        ref = apply(serial, gmesh)
        serial.close()
    else:
        ref = [torch.empty(k, dtype=torch.float64, device=ctx.device) for k in sizes]
    worst = 0.0
    for k, (got, full) in enumerate(zip(mine, ref)):
        full = full.contiguous()
        dist.broadcast(full, src=0)
        sp, nc = (mesh.u, 3) if k % 2 == 0 else (mesh.p, 1)
        ids = torch.from_numpy(np.asarray(sp.global_ids[: sp.n_owned]).astype(np.int64)).to(ctx.device)
        want = full.reshape(-1, nc)[ids]
        have = got.reshape(-1, nc)[: sp.n_owned]
        num = torch.stack([((have - want) ** 2).sum(), (want ** 2).sum()])
        dist.all_reduce(num)
        worst = max(worst, float(torch.sqrt(num[0] / num[1])))
    return {"max_rel_l2_diff": worst, "mesh": "x".join(map(str, n)) + " cells", "checked": ["residual_u", "residual_p", "vmult_u", "vmult_p"],
            "note": "N-rank partition vs single-rank operator on the same global problem (global-id seeded inputs)"}


def level_set_timings(ctx, n, steps, warmup):
    """Level-set operator applications at n^3 cells, FE_Q_iso_Q1(3), one GPU, each with its
    algorithmic-bytes model; returns (dict per operator, workload description, DoFs)."""
    import torch
    from adaflo_b200 import capi
    from adaflo_b200.mesh import BoxMesh, seeded_uniform
    from adaflo_b200.operators import LevelSetOperators, MatrixFree
    n_sub = 3
    mesh = BoxMesh((n, n, n), build_coords=False)
    mesh.build_ls_space(n_sub, build_coords=False)
    mf = MatrixFree(ctx, mesh, "cartesian")
    h_min = float(mesh.h.min())
    eps = 1.5
    A = LevelSetOperators(mf, capi.LSParams(3, n_sub, 0, 0, eps, eps * h_min / n_sub, 4.0, h_min), mesh.cell_diameters())
    A.set_time(75.0, -100.0, 25.0, 0.02)
    n_dofs = mesh.ls.n_owned
    pts_per_dof, idx_per_dof = (2 * n_sub) ** 3 / n_sub ** 3, 4.0 * (n_sub + 1) ** 3 / n_sub ** 3
    for which in (0, 1):   # per-point streams: a smooth unit field (values do not matter for the timing)
        A.pointdata(which).fill_(0.57735)

    def dev(a):
        return torch.from_numpy(np.ascontiguousarray(a)).to(ctx.device)
    src = [dev(seeded_uniform(mesh.ls.global_ids, 1, 100 + k)) for k in range(3)]
    dst = [mf.initialize_dof_vector(2) for _ in range(3)]
    peak, _ = measured_peak()
    # fp64 work of the cell kernels per DoF (csrc/ls_ops.cu): the projection operator is not HBM-bound on its byte model
    ops = {
        "advance_concentration_vmult": (lambda: A.advance_concentration_vmult(dst[0], src[0]), 16 + idx_per_dof + 24 * pts_per_dof),
        "reinitialization_vmult": (lambda: A.reinitialization_vmult(dst[0], src[0], False), 16 + idx_per_dof + 24 * pts_per_dof),
        "compute_normal_vmult": (lambda: A.compute_normal_vmult(dst, src), 3 * 16 + idx_per_dof),
        "compute_curvature_vmult": (lambda: A.compute_curvature_vmult(dst[0], src[0], True), 16 + idx_per_dof),
    }
    out = {}
    for name, (fn, bpd) in ops.items():
        for _ in range(max(warmup, 3)):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        gbs = bpd * n_dofs / (ms * 1e-3) / 1e9
        out[name] = {"ms": ms, "dofs_per_s": n_dofs / (ms * 1e-3), "bytes_per_dof": bpd, "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / peak}
    workload = f"level-set operators on {n}^3 cells, FE_Q_iso_Q1({n_sub}), {n_dofs} DoFs, QIterated(QGauss(2),{n_sub}), Cartesian"
    del A, mf, src, dst
    return out, workload, n_dofs


LS_BYTE_MODEL = ("per level-set DoF: 8 src + 8 dst per component, 4 B x (n_sub+1)^3 indices per cell, 24 B per quadrature point for the "
                 "stored convection velocity / frozen normal (advection, reinitialisation); the projection operators (normal, curvature) "
                 "carry no per-point stream and are bound by fp64 issue and the scatter, not by HBM")


def run_level_set(args):
    """--workload ls: the level-set operator timings alone (the default run carries them under the key level_set)."""
    import torch
    from adaflo_b200.operators import Context
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    ctx = Context(0)
    l0 = ctx.lib.adaflo_b200_kernel_launch_count()
    out, workload, _ = level_set_timings(ctx, args.n, args.steps, args.warmup)
    peak, which_peak = measured_peak()
    line = {"metric": "ls_operator_dofs_per_s", "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "dtype": "f64",
            "data": "synthetic", "higher_is_better": True, "config": {"workload": workload, "byte_model": LS_BYTE_MODEL},
            "peak_gbs": peak, "peak_source": which_peak, "operators": out,
            "gpu_launches": int(ctx.lib.adaflo_b200_kernel_launch_count() - l0)}
    print(json.dumps(line))


def run_generic(args):
    """--workload generic: the degree / dimension dispatch path (csrc/ns_generic.cu, navier_stokes_matrix.cc:64-82) -- Newton vmult
    of Q_k/Q_(k-1) on uniform meshes with the pairs of the reference's 2-D configurations (tests/beltrami_2d.prm: Q4/Q3).  Extra
    line, not the headline: the path is the correctness path for the small configurations."""
    import torch
    from adaflo_b200 import capi
    from adaflo_b200.mesh import TensorMesh, seeded_uniform
    from adaflo_b200.operators import BlockVector, Context, MatrixFreeGeneric, NavierStokesMatrix
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    ctx = Context(0)
    l0 = ctx.lib.adaflo_b200_kernel_launch_count()
    peak, which_peak = measured_peak()
    out = {}
    for dim, k, n in ((2, 2, (512, 512)), (2, 4, (256, 256)), (3, 3, (32, 32, 32))):
        mesh = TensorMesh(dim, n, k)
        mf = MatrixFreeGeneric(ctx, mesh)
        old, old_old = BlockVector.like(mf), BlockVector.like(mf)
        A = NavierStokesMatrix(capi.FlowParams(capi.INCOMPRESSIBLE, capi.NEWTON, 0, 0, 1.0, 0.01, 0.0, 0.0, 0.0, 0.0, 0.5), 0, 1, 0, 1, old, old_old)
        A.initialize(mf, capi.TimeWeights(30.0, -40.0, 10.0, 1.0, 2.0, -1.0), True)

        def dev(a):
            return torch.from_numpy(np.ascontiguousarray(a)).to(ctx.device)
        sol = BlockVector(dev(seeded_uniform(mesh.u_ids, dim, 1)), dev(seeded_uniform(mesh.p_ids, 1, 2)))
        rhs, zero, dst = BlockVector.like(mf), BlockVector.like(mf), BlockVector.like(mf)
        A.residual(rhs, sol, zero)   # stores the linearisation point
        src = BlockVector(dev(seeded_uniform(mesh.u_ids, dim, 11)), dev(seeded_uniform(mesh.p_ids, 1, 12)))
        for _ in range(max(args.warmup, 3)):
            A.vmult(dst, src)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(args.steps):
            A.vmult(dst, src)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        n_dofs = int(sol.block(0).numel() + sol.block(1).numel())
        n_cells = int(np.prod(n))
        # algorithmic bytes: src + dst, 4-byte cell->dof lists, dim + dim^2 doubles per Gauss point (Newton linearisation point)
        alg = 16.0 * n_dofs + n_cells * (4.0 * ((k + 1) ** dim + k ** dim) + 8.0 * (dim + dim * dim) * (k + 1) ** dim)
        gbs = alg / (ms * 1e-3) / 1e9
        out[f"{dim}d_Q{k}Q{k - 1}_{'x'.join(map(str, n))}"] = {"ms": ms, "dofs": n_dofs, "dofs_per_s": n_dofs / (ms * 1e-3), "bytes_per_dof": alg / n_dofs,
                                                             "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / peak}
        A.clear()
        mf.close()
    line = {"metric": "ns_generic_vmult_dofs_per_s", "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": max(args.warmup, 3), "dtype": "f64",
            "data": "synthetic", "higher_is_better": True,
            "config": {"workload": "Newton vmult of Taylor-Hood Q_k/Q_(k-1) on uniform Cartesian meshes, generic sum-factorisation kernel (one block per cell)"},
            "peak_gbs": peak, "peak_source": which_peak, "operators": out, "gpu_launches": int(ctx.lib.adaflo_b200_kernel_launch_count() - l0)}
    print(json.dumps(line))


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
    ap.add_argument("--strong", action="store_true", help="strong scaling: n^3 cells in total, split over the GPUs (default: weak, n^3 per GPU)")
    ap.add_argument("--no-level-set", action="store_true", help="skip the level-set operator timings of the default line")
    ap.add_argument("--workload", default="ns", choices=["ns", "ls", "generic"],
                    help="ns: the headline Navier-Stokes vmult line; ls: level-set operator timings (extra); generic: the degree-dispatch path (extra)")
    ap.add_argument("--kernel-only", action="store_true", help="developer A/B mode: kernel timings + checksum only (not a bench line)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "ls":
        return run_level_set(args)
    if args.workload == "generic":
        return run_generic(args)

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

    cross_n = parity_vs_single_gpu(ctx, rank, world, local_rank) if world > 1 else None
    if cross_n is not None and cross_n["max_rel_l2_diff"] > 1e-12:
        raise SystemExit(f"bench.py: {world}-rank result differs from the single-rank operator: {cross_n}")

    n = args.n
    t_setup = time.perf_counter()
    # weak scaling: n x n x (n * world) cells on [-1,1]^2 x [-1, -1 + 2 world]
    if args.strong:   # strong scaling (extra measurement): the n^3 box is split into N z-slabs
        mesh = BoxMesh((n, n, n), rank=rank, nranks=world)
    else:
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
    # The pre-roll count is fixed BEFORE the loop and agreed between the ranks (every vmult carries NCCL send/recv and an
    # all-reduce, so all ranks must issue the same number of them; a wall-clock bound per rank deadlocked round 1's N=2 run).
    pre = torch.tensor([max(10, int(args.preroll * 1e3 / max(b_ms / args.steps, 1e-3)))], dtype=torch.int64, device=ctx.device)
    if world > 1:
        dist.all_reduce(pre, op=dist.ReduceOp.MAX)
    n_preroll = int(pre.item())
    barrier()
    for i in range(n_preroll):
        A.vmult(dst, src)
        if i % 10 == 9:
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

    # ---- the residual kernel (producer of the linearisation point; once per Newton iteration)
    res_vecs = (BlockVector.like(mf), BlockVector.like(mf))
    for _ in range(2):
        A.residual(res_vecs[0], sol, res_vecs[1])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(5):
        A.residual(res_vecs[0], sol, res_vecs[1])
    e1.record()
    barrier()
    residual_ms = e0.elapsed_time(e1) / 5
    del res_vecs
    # algorithmic bytes per DoF of residual: 3 velocity vectors + pressure read (3 x 8 + ...), rhs read-modify-write 16,
    # indices 5.6, linearisation point written 103.68
    res_bytes = (24.0 + 16.0 + 5.6 + 103.68) * n_dofs_rank

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
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if args.strong else "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"ns_vmult Q2/Q1 3D box {n}^3 cells{' in total' if args.strong else '/GPU'} ({n_dofs_rank} DoFs/GPU, {n_dofs} total), Newton stored linearisation, "
                                   "Cartesian, constant coefficients, Dirichlet u on all faces, pressure mean fixed",
                       "l2": "inputs larger than L2 (vectors 0.42 GB, linearisation data 5.4 GB per GPU)",
                       "partition": f"z-slabs x{world}", "setup_s": round(t_setup, 1), "numa_node_rank0": numa_node},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": ncu_traffic()[0] if n == 128 and world == 1 else None,
                         "traffic_source": f"profiles/{ncu_traffic()[1]} (ncu --set full, same workload, 1 GPU)", "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({which})",
                         "kernel": "ns_vmult_r1_kernel<with_p, newton, cartesian> (persistent, 148 blocks x 11 warps)", "kernel_ms": k_avg * 1e3,
                         "algorithmic_bytes_per_launch": alg_bytes, "bytes_per_dof": B_PER_DOF_N},
            "e2e": {"value": n_dofs / e2e_t, "unit": UNIT, "h2d_bytes_per_step": 8 * n_dofs, "d2h_bytes_per_step": 8 * n_dofs,   # all ranks
                    "ms_per_step": e2e_t * 1e3, "rel_diff_vs_device_path": err},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "residual": {"ms": residual_ms, "bytes_per_dof": 149.28, "achieved_gbs": res_bytes / (residual_ms * 1e-3) / 1e9,
                         "frac_of_hbm_peak": res_bytes / (residual_ms * 1e-3) / 1e9 / peak,
                         "note": "NavierStokesMatrix::residual incl. ghost import, sadd and the store of the linearisation point (rank 0 timing)"},
            "nonlinear_step": nl,
            "parity_vs_1gpu": cross_n,
            "burst": {"ms_per_step": b_ms / args.steps, "kernel_ms": b_k * 1e3, "value": n_dofs / (b_ms / args.steps * 1e-3),
                      "roofline_frac": B_PER_DOF_N * n_dofs_rank / b_k / 1e9 / peak,
                      "note": "same K steps right after warm-up, before the GPU reaches its power limit (rank 0 timing)"},
        }
        if world == 1 and not args.no_level_set:   # extra keys: the level-set operators on the same box, after the timed region
            del A
            torch.cuda.empty_cache()
            ls_ops, ls_workload, _ = level_set_timings(ctx, n, 10, 5)
            line["level_set"] = {"workload": ls_workload, "byte_model": LS_BYTE_MODEL, "operators": ls_ops}
        if not args.no_cpu_baseline and world == 1:   # cpu_baseline: rank 0 at N = 1 only
            cb, _ = cpu_port_sample(n, max(2, min(n, args.cpu_layers)), 10, 2)
            line["cpu_baseline"] = cb
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
