"""Loader of the native code.  There is NO fallback: if the CUDA extension is missing the import
fails loudly, and CPU tensors are rejected by the extension itself ("Not implemented on the CPU",
same behaviour as the reference, models/ops/src/ms_deform_attn.h:39)."""
from __future__ import annotations

import importlib
import sys
import sysconfig
from pathlib import Path  # noqa: E402

import os

# MSDA_NATIVE_DIR: developer override used by tuning sweeps to load an alternative BUILD of the same
# sources (tools/variants.py); it never points at a different implementation.
NATIVE_DIR = Path(os.environ.get("MSDA_NATIVE_DIR") or (Path(__file__).resolve().parent / "_native"))
MODULE_NAME = "MultiScaleDeformableAttention_update"   # models/ops/setup.py:53
CAPI_NAME = "libmsda_b200.so"

_module = None


def native_dir() -> Path:
    return NATIVE_DIR


def capi_path() -> Path:
    return NATIVE_DIR / CAPI_NAME


def module_path() -> Path:
    return NATIVE_DIR / f"{MODULE_NAME}{sysconfig.get_config_var('EXT_SUFFIX')}"


def install_path() -> None:
    """Put the directory holding MultiScaleDeformableAttention_update.*.so on sys.path so that the
    reference's own `import MultiScaleDeformableAttention_update as MSDA`
    (models/ops/functions/ms_deform_attn_func.py:18) resolves to the B200 build."""
    p = str(NATIVE_DIR)
    if p not in sys.path:
        sys.path.insert(0, p)


def load():
    """Import and return the torch binding; raises ImportError with build instructions if absent."""
    global _module
    if _module is not None:
        return _module
    if not module_path().exists() or not capi_path().exists():
        raise ImportError(
            f"{MODULE_NAME}: native extension not built (expected {module_path()} and {capi_path()}). "
            "Build it with `python tce-rvos_b200/build.py` (needs nvcc, no GPU). "
            "There is no CPU or PyTorch fallback for this operator.")
    import torch  # noqa: F401  (libtorch must be loaded before the extension)
    install_path()
    mod = importlib.import_module(MODULE_NAME)
    if not getattr(mod, "__b200_native__", False):
        raise ImportError(
            f"{MODULE_NAME} resolved to {getattr(mod, '__file__', '?')}, which is not the B200 build; "
            f"put {NATIVE_DIR} first on sys.path")
    _module = mod
    return mod
