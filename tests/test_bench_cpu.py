"""bench.py's CPU arm (`--impl reference`: the oracle on the host cores) prints the contract's JSON
line for the default workload and for configs[0]; the product arm refuses to run without a GPU
instead of falling back.  CPU only."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step",
        "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"}


def run(*args):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + list(args),
                       capture_output=True, text=True, cwd=ROOT, timeout=600)
    return p


@pytest.mark.parametrize("extra", [["--cpu-crop-s", "0.5"], ["--workload", "tamy"]])
def test_reference_arm_line(extra):
    p = run("--impl", "reference", "--steps", "1", "--warmup", "0", *extra)
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert KEYS <= set(line), KEYS - set(line)
    assert line["impl"] == "reference" and line["higher_is_better"] is True
    assert line["metric"] == "gem_tf_bins_iters_per_s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["value"] == line["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in line["config"]
    if "tamy" in extra:
        assert line["config"]["tf_bins"] == 1025 * 1122 and "tamy.wav" in line["data"]


def test_product_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    p = run("--steps", "1", "--warmup", "3", "--no-cpu-baseline")
    assert p.returncode != 0  # no CPU fallback: the CUDA path fails loudly
