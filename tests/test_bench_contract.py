"""bench.py contract checks that run without a GPU: the reference arm (CPU oracle) prints exactly ONE JSON line on
stdout -- everything else (library banners, progress) goes to stderr -- carrying the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_contract_keys():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "tiny",
                        "--steps", "1", "--warmup", "0", "--cpu-budget", "2"], capture_output=True, text=True, timeout=300,
                       cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.split("\n") if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True and d["n_gpus"] == 1
    assert d["value"] > 0 and d["steps"] == 1 and d["config"]["workload"] == "tiny"
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]


def test_ours_arm_fails_loudly_without_a_gpu():
    """No CPU fallback: on a box without CUDA the product arm must exit non-zero instead of timing something else."""
    import torch
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", "tiny", "--steps", "1", "--warmup", "0",
                        "--no-cpu-baseline"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode != 0
    assert not [l for l in r.stdout.split("\n") if l.strip().startswith("{")]


def test_reference_arm_under_torchrun_prints_once_and_uses_the_host_cores():
    """N > 1 launch of the reference arm (the driver uses torchrun for every N > 1): rank 0 alone measures and prints,
    the other rank exits 0 without output; torchrun's OMP_NUM_THREADS=1 default must not throttle the CPU arm."""
    env = {k: v for k, v in os.environ.items() if k != "OMP_NUM_THREADS"}
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "bench.py"),
                        "--impl", "reference", "--gpus", "2", "--workload", "tiny", "--steps", "1", "--warmup", "0",
                        "--cpu-budget", "2"], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.split("\n") if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    assert d["cpu_baseline"]["cores"] == min(ncpu, 32)          # fixed thread policy of the CPU arm (bench.CPU_ARM_MAX_THREADS)
    assert d["cpu_baseline"]["tile_stride"] == 1 and d["config"] == {"workload": "tiny", "P": 6000, "faces": 2000, "K": 3, "width": 320,
                                                                        "height": 240, "sh_degree": 3, "cameras": 4}

