"""Drop-in for the reference's compiled extension module ``MultiScaleDeformableAttention``
(built by ops/setup.py from ops/src/vision.cpp; imported at
ops/functions/ms_deform_attn_func.py:23).

Put this directory on ``sys.path`` (or call ``install_dropin()``) and the reference's
``MSDeformAttnFunction`` / ``MSDeformAttn`` / pixel decoder run unchanged on the sm_100a
kernels.  Exactly the two names of vision.cpp:19-20 are exported.
"""
import importlib.util as _ilu
import os as _os
import sys as _sys

_NAME = "uni_encoder_code_b200"
if _NAME not in _sys.modules:
    _pkg = _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__)))
    _spec = _ilu.spec_from_file_location(_NAME, _os.path.join(_pkg, "__init__.py"),
                                         submodule_search_locations=[_pkg])
    _mod = _ilu.module_from_spec(_spec)
    _sys.modules[_NAME] = _mod
    try:
        _spec.loader.exec_module(_mod)
    except BaseException:
        _sys.modules.pop(_NAME, None)
        raise

from uni_encoder_code_b200.ops import ms_deform_attn_backward, ms_deform_attn_forward  # noqa: E402,F401

__all__ = ["ms_deform_attn_forward", "ms_deform_attn_backward"]
