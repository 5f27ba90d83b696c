#!/usr/bin/env python
"""Stage the handful of UNMODIFIED reference Python files of the hot path under git-ignored
``baseline/_ref/py/`` so that they travel to the GPU box with the gpurun snapshot (like
``baseline/_ref/ref_msda_cuda.so``): /root/reference itself does not exist there.

    python baseline/stage_reference_py.py          # build container only (needs /root/reference)

Staged (byte-identical copies, directory layout of model/modeling kept):
  pixel_decoder/msdeformattn.py                      MSDeformAttnPixelDecoder, encoder (pd.py)
  pixel_decoder/ops/functions/{__init__,ms_deform_attn_func}.py   MSDeformAttnFunction, core_pytorch
  pixel_decoder/ops/modules/{__init__,ms_deform_attn}.py           MSDeformAttn
  transformer_decoder/{position_encoding,transformer}.py           imported by pd.py

Users: tests/ref_import.py (the reference stack running on the drop-in shim, on a B200) and
bench.py's reference arm (the reference's own ms_deform_attn_core_pytorch as the CPU baseline).
Nothing under uni-encoder-code_b200/ reads these files, and none of them enters the git history.
"""
import hashlib
import json
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_MODELING = "/root/reference/model/modeling"
OUT = os.path.join(ROOT, "baseline", "_ref", "py", "modeling")
FILES = [
    "pixel_decoder/msdeformattn.py",
    "pixel_decoder/ops/functions/__init__.py",
    "pixel_decoder/ops/functions/ms_deform_attn_func.py",
    "pixel_decoder/ops/modules/__init__.py",
    "pixel_decoder/ops/modules/ms_deform_attn.py",
    "transformer_decoder/position_encoding.py",
    "transformer_decoder/transformer.py",
]


def main():
    if not os.path.isdir(REF_MODELING):
        raise SystemExit("reference checkout not present; nothing staged")
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(REF_MODELING, rel), os.path.join(OUT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = hashlib.sha256(open(dst, "rb").read()).hexdigest()
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump({"source": REF_MODELING, "sha256": manifest}, f, indent=1)
    print("staged", len(FILES), "files under", OUT)


if __name__ == "__main__":
    main()
