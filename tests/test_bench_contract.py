"""bench.py contract on a CPU-only box: the reference arm prints one JSON line with the agreed keys (tiny workload),
and the product arm refuses to run without a CUDA device instead of falling back."""
import json
import os
import subprocess
import sys

from conftest import ROOT, _have_gpu

import pytest


def _run(args, timeout=300):
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                          timeout=timeout, cwd=ROOT)


@pytest.mark.parametrize("workload", ["diffuse-64^3-uniform-192dir", "point-32^3-uniform-1src"])
def test_reference_arm_prints_one_json_line(workload):
    r = _run(["--impl", "reference", "--workload", workload, "--steps", "1", "--warmup", "0", "--cpu-threads", "2"])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "higher_is_better", "scaling", "dtype", "data",
              "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["config"]["workload"] == workload
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == 2
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0


@pytest.mark.skipif(_have_gpu(), reason="only meaningful without a CUDA device")
def test_product_arm_has_no_cpu_fallback():
    r = _run(["--workload", "diffuse-64^3-uniform-192dir", "--steps", "1", "--warmup", "3", "--no-cpu-baseline"])
    assert r.returncode != 0
    assert "no CUDA device" in (r.stderr + r.stdout)
