"""Memory-safety run with the instrumented library (`make -C uni-encoder-code_b200/csrc checked`):
device-side asserts on every corner offset of the fast and the generic kernels.  compute-sanitizer
is not available on the GPU pool, so this is the bounds check; it re-runs the parity tests that
stress borders, ragged / odd shapes, out-of-range and non-finite locations in a subprocess that
loads libmsda_b200_checked.so through MSDA_B200_LIB."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHECKED = os.path.join(ROOT, "uni-encoder-code_b200", "lib", "libmsda_b200_checked.so")


def test_parity_suite_passes_with_device_side_bounds_asserts():
    if os.environ.get("MSDA_B200_LIB"):
        pytest.skip("already inside the instrumented run")
    if not os.path.exists(CHECKED):
        res = subprocess.run(["make", "-C", os.path.join(ROOT, "uni-encoder-code_b200", "csrc"), "checked", "-j4"],
                             capture_output=True, text=True)
        assert res.returncode == 0, res.stdout + res.stderr
    env = dict(os.environ, MSDA_B200_LIB=CHECKED)
    res = subprocess.run(
        [sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_parity_gpu.py"),
         os.path.join(ROOT, "tests", "test_fused_gpu.py"), "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider",
         "-k", "golden or random_shapes or lattice or non_finite or oracle or fused_matches"],
        capture_output=True, text=True, env=env, cwd=ROOT, timeout=1500)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-2000:]
    assert "passed" in res.stdout
