"""A short run of tests/harness/gpu_soak.py (randomised mixed traffic, fuzzed documents, reloads, window roll-overs; every
decision, counter and metric row against the oracle)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [11, 12])
def test_soak_short(seed):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "harness", "gpu_soak.py"), "8", str(seed)], capture_output=True,
                         text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0 and "soak ok" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
