"""Import pieces of the reference without running its package __init__ files, which need
detectron2.  Test infrastructure.

Where the files come from: /root/reference in the build container; on the GPU box, which has no
checkout, the byte-identical copies that baseline/stage_reference_py.py (run by build()) staged
under git-ignored baseline/_ref/py/ and that travel with the gpurun snapshot."""
import importlib
import os
import sys
import types

REF = "/root/reference"
_HERE = os.path.dirname(os.path.abspath(__file__))
_STAGED = os.path.join(os.path.dirname(_HERE), "baseline", "_ref", "py", "modeling")
REF_MODELING = os.path.join(REF, "model", "modeling")
if not os.path.isdir(REF_MODELING):
    REF_MODELING = _STAGED
STUBS = os.path.join(_HERE, "ref_stubs")


def available():
    return os.path.isfile(os.path.join(REF_MODELING, "pixel_decoder", "msdeformattn.py"))


def source():
    """'checkout' (build container) or 'staged' (copies shipped to the GPU box)."""
    return "staged" if REF_MODELING == _STAGED else "checkout"


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
