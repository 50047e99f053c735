# mirrors models/ops/functions/__init__.py:9-10 of the reference
from .ms_deform_attn_func import MSDeformAttnFunction
from .ms_deform_attn_func import MSDeformAttnFunction3D
from .ms_deform_attn_func import MSDeformAttnPrologueFunction

__all__ = ["MSDeformAttnFunction", "MSDeformAttnFunction3D", "MSDeformAttnPrologueFunction"]
