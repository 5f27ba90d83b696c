"""CPU tests: the oracle (oracle/) against the golden vectors produced by the reference's own
ms_deform_attn_core_pytorch (tests/golden/make_golden.py), and hand-computed integer KATs."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_NAMES, load_golden, rel_err

REF_FUNC = "/root/reference/model/modeling/pixel_decoder/ops/functions/ms_deform_attn_func.py"


def test_golden_fixtures_present():
    assert len(GOLDEN_NAMES) >= 10, GOLDEN_NAMES


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_c_oracle_fp64_matches_reference_golden(oracle, name):
    g = load_golden(name)
    args = (g["value"], g["spatial_shapes"], g["level_start_index"], g["sampling_locations"],
            g["attention_weights"])
    out = oracle.forward(*args, dtype=np.float64)
    assert np.abs(out - g["ref_output"]).max() <= 1e-12
    gv, gl, gw = oracle.backward(g["grad_output"], *args, dtype=np.float64)
    assert rel_err(gv, g["ref_grad_value"]) <= 1e-12
    assert rel_err(gw, g["ref_grad_attn_weight"]) <= 1e-12
    if bool(g["loc_grad_pinned"]):
        assert rel_err(gl, g["ref_grad_sampling_loc"]) <= 1e-11
    if name == "all_oob":
        assert not out.any() and not gv.any() and not gl.any() and not gw.any()


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_c_oracle_fp32_within_north_star_tolerance(oracle, name):
    g = load_golden(name)
    args = (g["value"], g["spatial_shapes"], g["level_start_index"], g["sampling_locations"],
            g["attention_weights"])
    out = oracle.forward(*args, dtype=np.float32)
    assert np.abs(out.astype(np.float64) - g["ref_output"]).max() <= 1e-5


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_grid_sample_restatement_matches_reference_golden(oracle, name):
    g = load_golden(name)
    t = lambda k: torch.from_numpy(g[k]).double()
    out, gv, gl, gw = oracle.core_grid_sample_grads(
        t("value"), torch.from_numpy(g["spatial_shapes"]), t("sampling_locations"),
        t("attention_weights"), t("grad_output"))
    assert (out.numpy() - g["ref_output"]).__abs__().max() <= 1e-12
    assert rel_err(gv.numpy(), g["ref_grad_value"]) <= 1e-12
    assert rel_err(gw.numpy(), g["ref_grad_attn_weight"]) <= 1e-12
    assert rel_err(gl.numpy(), g["ref_grad_sampling_loc"]) <= 1e-11


@pytest.mark.skipif(not os.path.exists(REF_FUNC), reason="reference checkout not present")
def test_restatements_match_live_reference(oracle, pkg):
    """In the build container the reference itself is importable: compare on fresh inputs."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("ref_func_live", REF_FUNC)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    inp = pkg.synthetic.make_inputs([(3, 4), (5, 9), (11, 17)], batch=2, heads=4, mode="model", seed=123)
    v, loc, w = (inp[k].double() for k in ("value", "sampling_locations", "attention_weights"))
    ref = mod.ms_deform_attn_core_pytorch(v, inp["spatial_shapes"], loc, w)
    mine = oracle.core_grid_sample(v, inp["spatial_shapes"], loc, w)
    assert (ref - mine).abs().max().item() <= 1e-12
    c_out = oracle.forward(inp["value"], inp["spatial_shapes"], inp["level_start_index"],
                           inp["sampling_locations"], inp["attention_weights"], dtype=np.float64)
    assert np.abs(c_out - ref.numpy()).max() <= 1e-12


def test_integer_known_answers(oracle):
    """Hand-computed decomposition on a 4x8 level (H=4, W=8), M=2, D=4, level_start=10."""
    shapes = np.array([[4, 8]], np.int64)
    lsi = np.array([10], np.int64)
    M, D, S = 2, 4, 10 + 32
    pts = [
        # (x, y)              valid h_low w_low cmask
        ((0.5 / 8, 0.5 / 4),   1, 0, 0, 0b1111),   # centre of pixel (0,0): x=0,y=0 exactly
        ((0.0, 0.0),           1, -1, -1, 0b1000), # x=y=-0.5: only corner (0,0) = corner 3
        ((1.0, 1.0),           1, 3, 7, 0b0001),   # x=7.5,y=3.5: only (3,7) = corner 0
        ((-1.0 / 8, 0.5),      0, 1, -2, 0),       # x=-1.5 -> invalid
        ((1.0 + 0.5 / 8, 0.5), 0, 1, 8, 0),        # x=8.0 -> not < W -> invalid
        ((0.95, 0.3),          1, 0, 7, 0b0101),   # x=7.1,y=0.7: w_high=8 out
        ((2.25 / 8, 3.75 / 4), 1, 3, 1, 0b0011),   # x=1.75,y=3.25: h_high=4 out
    ]
    loc = np.zeros((1, len(pts), M, 1, 1, 2), np.float32)
    for q, (xy, *_rest) in enumerate(pts):
        loc[0, q, :, 0, 0] = xy
    idx, off = oracle.indices(shapes, lsi, loc, (1, S, M, D), dtype=np.float32)
    for q, (_xy, valid, h_low, w_low, cmask) in enumerate(pts):
        for m in range(M):
            got = idx[0, q, m, 0, 0]
            assert got[0] == valid, (q, got)
            if valid:
                assert tuple(got[1:]) == (h_low, w_low, cmask), (q, got)
                for k in range(4):
                    h, w = h_low + (k >> 1), w_low + (k & 1)
                    want = ((10 + h * 8 + w) * M + m) * D if (cmask >> k) & 1 else -1
                    assert off[0, q, m, 0, 0, k] == want, (q, m, k)
            else:
                assert got[3] == 0 and (off[0, q, m, 0, 0] == -1).all()


def test_oracle_linearity_and_zero_weight(oracle, pkg):
    inp = pkg.synthetic.make_inputs([(4, 5), (8, 10)], batch=1, heads=2, channels=8, points=3,
                                    num_query=13, mode="uniform", seed=5)
    a = (inp["spatial_shapes"], inp["level_start_index"], inp["sampling_locations"])
    o1 = oracle.forward(inp["value"], *a, inp["attention_weights"])
    o2 = oracle.forward(2 * inp["value"], *a, inp["attention_weights"])
    assert np.allclose(o2, 2 * o1, rtol=0, atol=1e-12)
    o0 = oracle.forward(inp["value"], *a, 0 * inp["attention_weights"])
    assert not o0.any()


def test_fp32_geometry_mode(oracle, pkg):
    """geometry=float32: on power-of-two levels the fp32 product loc*W is exact (only `- 0.5` can
    round, and only for coordinates in (-0.5, 0)), so it reproduces pure fp64 to ~1e-7; elsewhere it
    agrees with the all-fp32 restatement to accumulation rounding."""
    inp = pkg.synthetic.make_inputs([(4, 8), (8, 16), (16, 32)], 1, heads=2, mode="model", seed=3)
    a = (inp["value"], inp["spatial_shapes"], inp["level_start_index"], inp["sampling_locations"],
         inp["attention_weights"])
    assert np.abs(oracle.forward(*a) - oracle.forward(*a, geometry=np.float32)).max() <= 1e-6
    for x, y in zip(oracle.backward(inp["grad_output"], *a),
                    oracle.backward(inp["grad_output"], *a, geometry=np.float32)):
        assert rel_err(y, x) <= 1e-6
    inp = pkg.synthetic.make_inputs([(12, 39), (24, 78)], 1, heads=2, mode="model", seed=4)
    a = (inp["value"], inp["spatial_shapes"], inp["level_start_index"], inp["sampling_locations"],
         inp["attention_weights"])
    mixed = oracle.forward(*a, geometry=np.float32)
    assert np.abs(oracle.forward(*a, dtype=np.float32) - mixed).max() <= 2e-6
    assert np.abs(oracle.forward(*a) - mixed).max() <= 1e-4


def test_reference_binary_rounds_the_pixel_coordinate_with_one_fma():
    """The bit-exact contract depends on how nvcc contracts the reference's `loc * size - 0.5`
    (cuh:290-291): the reference op built for sm_100 (baseline/build_reference_cuda.py) must show a
    single `FFMA ..., -0.5` per coordinate in its forward and D=32 backward kernels, which is what
    decompose() and the fp32-geometry oracle reproduce (profiles/r2_reference_sass_ffma.md)."""
    import os
    import shutil
    import subprocess
    so = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref", "ref_msda_cuda.so")
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not (os.path.exists(so) and os.path.exists(cuobjdump)):
        pytest.skip("reference CUDA op or cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", so], capture_output=True, text=True, check=True).stdout
    per_kernel, cur = {}, None
    for line in sass.splitlines():
        if "Function :" in line:
            cur = line.split("Function :")[1].strip()
        elif cur and "FFMA" in line and "-0.5" in line:
            per_kernel[cur] = per_kernel.get(cur, 0) + 1
    fwd = [k for k in per_kernel if "ms_deformable_im2col_gpu_kernelIf" in k]
    bwd = [k for k in per_kernel if "blocksize_aware_reduce_v1IfLj32" in k]
    assert fwd and bwd, sorted(per_kernel)[:5]
    assert all(per_kernel[k] >= 2 for k in fwd + bwd)      # x and y
