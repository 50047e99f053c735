"""tce-rvos_b200 — B200-native (sm_100a) multi-scale deformable attention, the hot path of
haliphinx/TCE-RVOS, behind the reference's own operator boundary.

The directory name contains a hyphen, so import it through the shim at the repo root
(``import tce_rvos_b200``) or with ``importlib.import_module("tce-rvos_b200")``.

Public surface (mirrors the reference's models/ops package):
    functions.MSDeformAttnFunction, functions.MSDeformAttnFunction3D   (autograd Functions)
    modules.MSDeformAttn                                               (nn.Module)
    functions.MSDeformAttnValueEpilogue, functions.decoder_sample_topk (the passes either side of the op)
    capi                                                               (ctypes view of the C ABI)
    install_drop_in()   put the extension on sys.path under the reference's module name so that the
                        unmodified reference code imports it
"""
from ._ext import install_path as install_drop_in, load as load_extension, native_dir  # noqa: F401
from .functions import (MSDeformAttnFunction, MSDeformAttnFunction3D, MSDeformAttnPrologueFunction,  # noqa: F401
                        MSDeformAttnValueEpilogue, decoder_sample_topk)
from .modules import MSDeformAttn  # noqa: F401

__all__ = ["MSDeformAttnFunction", "MSDeformAttnFunction3D", "MSDeformAttnPrologueFunction", "MSDeformAttn",
           "MSDeformAttnValueEpilogue", "decoder_sample_topk",
           "install_drop_in", "load_extension", "native_dir"]
__version__ = "0.1.0"
