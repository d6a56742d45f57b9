"""The bench.py contract the driver depends on: one JSON line per arm, same `config` in both arms,
the keys of the measurement contract present.  CPU: the reference arm (oracle on host threads);
GPU: our arm at --small sizes."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
             "scaling", "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"}


def _run(args):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True,
                       text=True, timeout=900, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, p.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line():
    d = _run(["--impl", "reference", "--small", "--steps", "1", "--warmup", "0"])
    assert d["impl"] == "reference" and BASE_KEYS <= set(d)
    assert d["value"] > 0 and d["unit"] == "LOS/s" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert "workload" in d["config"] and "model" not in d["config"]


@pytest.mark.gpu
def test_our_arm_line_small():
    d = _run(["--small", "--steps", "1", "--warmup", "1"])
    assert BASE_KEYS | {"roofline", "gpu_launches", "clocks", "kernels", "lut_build", "voigt"} <= set(d)
    assert d["value"] > 0 and d["e2e"]["value"] > 0 and d["gpu_launches"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert d["batch"]["finite_positive"] and d["batch"]["e2e_equals_device_result"]
    assert d["roofline"]["bound"] == "tensor" and 0 < d["roofline"]["frac"] < 1.2
    assert d["cpu_baseline"]["value"] > 0
    r = _run(["--impl", "reference", "--small", "--steps", "1", "--warmup", "0"])
    assert r["config"] == d["config"]                     # both arms measure the same workload
    assert list(d)[-1] == "gpu_launches" or "value" in list(d)[-14:]   # headline numbers in the tail
