"""A torch-free C++ consumer of the C ABI, built with nvcc on the GPU box and checked against the C
oracle: the boundary really is plain pointers + sizes + a cudaStream_t."""
import os
import shutil
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_consumer_of_the_abi_matches_the_c_oracle(pkg, oracle, tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    oracle.build()
    lib_dir = os.path.dirname(pkg._lib.LIB_PATH)
    ora_dir = os.path.join(ROOT, "oracle", "_build")
    exe = str(tmp_path / "cabi_parity")
    subprocess.run([nvcc, "-O2", "-o", exe, os.path.join(ROOT, "tests", "cabi", "cabi_parity.cu"),
                    "-I", os.path.join(ROOT, "include"), "-L", lib_dir, "-lmsda_b200", "-L", ora_dir,
                    "-lmsda_oracle", "-Xlinker", f"-rpath={lib_dir}", "-Xlinker", f"-rpath={ora_dir}"],
                   check=True, capture_output=True, text=True)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "CABI PARITY OK" in res.stdout, res.stdout + res.stderr
