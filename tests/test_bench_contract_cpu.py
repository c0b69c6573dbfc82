"""bench.py pieces that run without a GPU: the reference arm prints one JSON line with the contract's keys (it times the CPU
port of the path on the host cores, oracle/ns_cpu_port.c), and the roofline's `traffic` comes from a committed ncu summary."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS="1")  # what torchrun sets: the arm must take the cores of its affinity mask anyway
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--n", "16", "--cpu-layers", "4", "--steps", "2",
                        "--warmup", "1"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "ns_vmult_dofs_per_s" and line["unit"] == "DoF/s"
    assert line["higher_is_better"] is True and line["gpu_launches"] == 0 and line["value"] > 0
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["value"] == line["value"] and cb["cores"] == len(os.sched_getaffinity(0))
    assert line["e2e"] == {"value": line["value"], "unit": "DoF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_is_silent_on_the_other_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--n", "16"], capture_output=True,
                       text=True, timeout=120, env=env, cwd=ROOT)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_traffic_comes_from_a_committed_capture():
    sys.path.insert(0, ROOT)
    import bench
    traffic, name = bench.ncu_traffic()
    assert name in bench.TRAFFIC_FILES and os.path.exists(os.path.join(ROOT, "profiles", name))
    alg = bench.B_PER_DOF_N * 53070468
    assert alg <= traffic < 1.15 * alg  # no wasted re-reads: DRAM traffic within 15 % of the algorithmic bytes
