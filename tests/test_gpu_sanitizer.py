"""compute-sanitizer over every kernel family (the reference has no sanitizer runs at all): memcheck for the
predicated / speculative loads (out-of-box rows, edge lanes, x-face cache, ghost planes) and racecheck for the shared-memory
hand-offs (cp.async slots, the block kernel's parked info line, the TMA stage ring)."""
import os
import shutil
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("tool", ["memcheck", "racecheck"])
def test_kernels_are_clean_under_compute_sanitizer(tool):
    exe = shutil.which("compute-sanitizer") or "/usr/local/cuda/bin/compute-sanitizer"
    if not os.path.exists(exe):
        pytest.skip("compute-sanitizer not installed")
    r = subprocess.run([exe, "--tool", tool, "--error-exitcode", "77", "--print-limit", "5", sys.executable,
                        os.path.join(ROOT, "tests", "sanitizer_case.py")], capture_output=True, text=True, timeout=1500,
                       env=dict(os.environ, SANITIZER_SKIP_TMA="1" if tool == "racecheck" else "0"))
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0 and "sanitizer case done" in r.stdout, tail
    clean = "ERROR SUMMARY: 0 errors" if tool == "memcheck" else "RACECHECK SUMMARY: 0 hazards displayed (0 errors, 0 warnings)"
    assert clean in r.stdout + r.stderr, tail
