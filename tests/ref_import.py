"""Import pieces of the reference (build container only) without running its package __init__
files, which need detectron2.  Test infrastructure; returns None-equivalents when /root/reference
is absent (the GPU box)."""
import importlib
import os
import sys
import types

REF = "/root/reference"
REF_MODELING = os.path.join(REF, "model", "modeling")
STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_stubs")


def available():
    return os.path.isdir(REF_MODELING)


def _pkg(name, path):
    if name not in sys.modules:
        m = types.ModuleType(name)
        m.__path__ = [path]
        sys.modules[name] = m
    return sys.modules[name]


def load():
    """-> namespace with MSDeformAttn, MSDeformAttnTransformerEncoderOnly, MSDeformAttnPixelDecoder,
    ms_deform_attn_core_pytorch, PositionEmbeddingSine of the reference."""
    if STUBS not in sys.path:
        sys.path.insert(0, STUBS)
    _pkg("refmodeling", REF_MODELING)
    _pkg("refmodeling.transformer_decoder", os.path.join(REF_MODELING, "transformer_decoder"))
    _pkg("refmodeling.pixel_decoder", os.path.join(REF_MODELING, "pixel_decoder"))
    pd = importlib.import_module("refmodeling.pixel_decoder.msdeformattn")
    func = importlib.import_module("refmodeling.pixel_decoder.ops.functions.ms_deform_attn_func")
    pe = importlib.import_module("refmodeling.transformer_decoder.position_encoding")
    ns = types.SimpleNamespace(
        MSDeformAttn=pd.MSDeformAttn,
        MSDeformAttnTransformerEncoderOnly=pd.MSDeformAttnTransformerEncoderOnly,
        MSDeformAttnPixelDecoder=pd.MSDeformAttnPixelDecoder,
        ms_deform_attn_core_pytorch=func.ms_deform_attn_core_pytorch,
        PositionEmbeddingSine=pe.PositionEmbeddingSine,
        ShapeSpec=importlib.import_module("detectron2.layers").ShapeSpec,
        pd_module=pd)
    return ns
