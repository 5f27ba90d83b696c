"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol
include/msda_b200.h declares; host-side argument validation mirrors the reference's error
behaviour; the product package never touches oracle/ and has no CPU fallback."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "msda_b200.h")
PKG_DIR = os.path.join(ROOT, "uni-encoder-code_b200")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(msda_b200_\w+)\s*\(", text)))


def test_header_symbols_all_exported(pkg):
    names = declared_symbols()
    assert len(names) >= 10
    lib = ctypes.CDLL(pkg._lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/msda_b200.h but not exported"
    assert set(names) == set(pkg._lib.SYMBOLS), "ctypes table and header disagree"


def test_abi_version_and_error_strings(pkg):
    lib = pkg._lib.lib
    assert lib.msda_b200_abi_version() == 2          # round 2: new entry points, group_norm gained channel_bias
    assert lib.msda_b200_error_string(0) == b"success"
    for code in (-1, -2, -3, -4):
        assert b"msda_b200" in lib.msda_b200_error_string(code)
    assert lib.msda_b200_launch_count() >= 0


def test_null_and_shape_errors_without_gpu(pkg):
    lib = pkg._lib.lib
    # NULL pointers and non-positive sizes are rejected before any CUDA call
    assert lib.msda_b200_forward_f32(None, None, None, None, None, 1, 1, 1, 1, 1, 1, 1, None, None) == -1
    one = ctypes.c_void_p(16)
    assert lib.msda_b200_forward_f32(one, one, one, one, one, 0, 1, 1, 1, 1, 1, 1, one, None) == -2
    assert lib.msda_b200_forward_f32(one, one, one, one, one, 1, 1, 1, 1, 17, 1, 1, one, None) == -3
    assert lib.msda_b200_backward_f32(one, one, one, one, one, one, 1, 1, 1, 1, 1, 1, -1, one, one, one, None) == -2


def test_options_roundtrip(pkg):
    pkg.set_option("tile_order", 1)
    assert pkg.get_option("tile_order") == 1
    pkg.set_option("tile_order", 0)
    with pytest.raises(RuntimeError):
        pkg.set_option("no_such_option", 1)


def _cpu_inputs(pkg):
    inp = pkg.synthetic.make_inputs([(2, 3), (4, 6)], batch=2, heads=2, channels=32, points=2,
                                    num_query=5, mode="uniform", seed=0)
    return (inp["value"], inp["spatial_shapes"], inp["level_start_index"], inp["sampling_locations"],
            inp["attention_weights"])


def test_cpu_tensors_raise_like_reference(pkg):
    # ops/src/ms_deform_attn.h:43,65: AT_ERROR("Not implemented on the CPU")
    args = _cpu_inputs(pkg)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        pkg.ms_deform_attn_forward(*args, 128)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        pkg.ms_deform_attn_backward(*args, torch.zeros(2, 5, 64), 128)
    with pytest.raises(RuntimeError, match="Not implemented on the CPU"):
        pkg.MSDeformAttnFunction.apply(*args, 128)


def test_dropin_module_exports_exactly_the_reference_names(pkg):
    pkg.install_dropin()
    import MultiScaleDeformableAttention as MSDA
    assert sorted(MSDA.__all__) == ["ms_deform_attn_backward", "ms_deform_attn_forward"]
    assert MSDA.ms_deform_attn_forward is pkg.ms_deform_attn_forward


def test_product_never_imports_oracle():
    for dirpath, _dirs, files in os.walk(PKG_DIR):
        if "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.lower().replace("no cpu or pytorch fallback", ""), \
                    f"{f} mentions oracle/: product code must not depend on test infrastructure"


def test_missing_library_fails_loudly(tmp_path):
    """With the .so absent the package import raises (no silent fallback)."""
    import shutil
    dst = tmp_path / "pkgcopy"
    shutil.copytree(PKG_DIR, dst, ignore=shutil.ignore_patterns("lib", "build", "__pycache__"))
    code = (
        "import importlib.util, sys\n"
        f"p = r'{dst}'\n"
        "spec = importlib.util.spec_from_file_location('x_pkg', p + '/__init__.py', submodule_search_locations=[p])\n"
        "m = importlib.util.module_from_spec(spec); sys.modules['x_pkg'] = m\n"
        "try:\n    spec.loader.exec_module(m)\nexcept RuntimeError as e:\n    print('RAISED', e); sys.exit(0)\n"
        "sys.exit(1)\n")
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert res.returncode == 0 and "RAISED" in res.stdout, res.stdout + res.stderr
