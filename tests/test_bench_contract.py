"""bench.py's contract as far as a CPU box can check it: the reference arm (the reference's own CPU decimator,
oracle/_ref or the C restatement) prints one JSON line with the agreed keys, under torchrun only rank 0 works,
and the GPU arm refuses to run without a device instead of falling back."""
import json
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")


def one_json_line(stdout):
    lines = [l for l in stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, stdout[-2000:]
    return json.loads(lines[0])


def check_reference_line(d, n_gpus):
    assert d["impl"] == "reference" and d["metric"] == "dsd_mbit_per_s" and d["unit"] == "Mbit/s"
    assert d["n_gpus"] == n_gpus and d["steps"] == 1 and d["warmup"] >= 3 and d["higher_is_better"] is True
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "workload" in d["config"] and "model" not in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "Mbit/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the default workload also carries the config the metric's x realtime is quoted on, dither off and on
    if "configs0" in d:
        for k in ("dither_off", "dither_on"):
            assert d["configs0"][k]["value"] > 0 and d["configs0"][k]["x_realtime"] > 1 and d["configs0"][k]["kind"] in ("reference", "port")


def test_reference_arm_line():
    r = subprocess.run([sys.executable, BENCH, "--impl", "reference", "--steps", "1", "--warmup", "1", "--seconds", "4"],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    check_reference_line(one_json_line(r.stdout), 1)


def test_reference_arm_under_torchrun_only_rank0_prints():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", str(port), BENCH, "--gpus", "2", "--impl", "reference", "--steps", "1",
                        "--warmup", "1", "--seconds", "4"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    check_reference_line(one_json_line(r.stdout), 2)


def test_gpu_arm_refuses_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("this box has a GPU")
    r = subprocess.run([sys.executable, BENCH, "--steps", "1", "--seconds", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0 and "no CUDA device" in (r.stderr + r.stdout)
    assert not [l for l in r.stdout.splitlines() if l.startswith("{")]
