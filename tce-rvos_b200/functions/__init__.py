# mirrors models/ops/functions/__init__.py:9-10 of the reference
from .ms_deform_attn_func import MSDeformAttnFunction
from .ms_deform_attn_func import MSDeformAttnFunction3D
from .ms_deform_attn_func import MSDeformAttnPrologueFunction
from .ms_deform_attn_func import MSDeformAttnValueEpilogue
from .ms_deform_attn_func import decoder_sample_topk

__all__ = ["MSDeformAttnFunction", "MSDeformAttnFunction3D", "MSDeformAttnPrologueFunction",
           "MSDeformAttnValueEpilogue", "decoder_sample_topk"]
