# mirrors models/ops/modules/__init__.py:9 of the reference
from .ms_deform_attn import MSDeformAttn

__all__ = ["MSDeformAttn"]
