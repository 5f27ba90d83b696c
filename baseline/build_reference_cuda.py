#!/usr/bin/env python
"""Optional same-GPU comparison baseline: the REFERENCE's own CUDA op, compiled for sm_100.

Run in the build container (needs /root/reference; no GPU required):

    python baseline/build_reference_cuda.py

The reference's extension (model/modeling/pixel_decoder/ops/src/) does not compile against torch 2.11
as shipped: `AT_DISPATCH_FLOATING_TYPES(value.type(), ...)` at ops/src/cuda/ms_deform_attn_cuda.cu:69
and :139 needs `value.scalar_type()`.  This script copies the sources to a temporary directory
OUTSIDE the repository, applies that two-line change there, builds with torch.utils.cpp_extension for
compute capability 10.0 and keeps only the resulting `baseline/_ref/ref_msda_cuda.so` (git-ignored,
shipped to the GPU box by gpurun).  No reference source enters the repository, and nothing in the
product or in bench.py's arms depends on this file: it is used by tools/bench_vs_reference_cuda.py and
by one skippable GPU test as an additional parity / speed reference.
"""
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SRC = "/root/reference/model/modeling/pixel_decoder/ops/src"
OUT = os.path.join(ROOT, "baseline", "_ref")


def main():
    if not os.path.isdir(REF_SRC):
        raise SystemExit("reference checkout not present; nothing built")
    os.environ["TORCH_CUDA_ARCH_LIST"] = "10.0"
    from torch.utils.cpp_extension import load
    with tempfile.TemporaryDirectory() as tmp:
        src = os.path.join(tmp, "src")
        shutil.copytree(REF_SRC, src)
        cu = os.path.join(src, "cuda", "ms_deform_attn_cuda.cu")
        text = open(cu).read()
        patched = text.replace("AT_DISPATCH_FLOATING_TYPES(value.type(), ", "AT_DISPATCH_FLOATING_TYPES(value.scalar_type(), ")
        assert patched.count("value.scalar_type()") == 2, "reference source changed"
        open(cu, "w").write(patched)
        build = os.path.join(tmp, "build")
        os.makedirs(build)
        mod = load(name="ref_msda_cuda",
                   sources=[os.path.join(src, "vision.cpp"), os.path.join(src, "cpu", "ms_deform_attn_cpu.cpp"), cu],
                   extra_include_paths=[src], extra_cflags=["-DWITH_CUDA", "-O3"],
                   extra_cuda_cflags=["-DWITH_CUDA", "-O3", "-DCUDA_HAS_FP16=1", "-D__CUDA_NO_HALF_OPERATORS__",
                                      "-D__CUDA_NO_HALF_CONVERSIONS__", "-D__CUDA_NO_HALF2_OPERATORS__"],
                   build_directory=build, with_cuda=True, verbose=False)
        os.makedirs(OUT, exist_ok=True)
        shutil.copy(mod.__file__, os.path.join(OUT, "ref_msda_cuda.so"))
    print("built", os.path.join(OUT, "ref_msda_cuda.so"))


if __name__ == "__main__":
    main()
