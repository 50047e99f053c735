"""ctypes view of the C ABI (include/msda_b200.h) — what a non-torch host (or a test) binds.

Nothing here computes anything: it declares the prototypes of libmsda_b200.so and offers two thin
helpers that pass torch tensors' ``data_ptr()`` (device API) or numpy arrays' buffers (host API).
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

from ._ext import capi_path

MSDA_F32, MSDA_BF16, MSDA_F16, MSDA_F64 = 0, 1, 2, 3
MSDA_E_NULL, MSDA_E_SHAPE, MSDA_E_DTYPE, MSDA_E_ALIGN = -1, -2, -3, -4
MSDA_E_WORKSPACE, MSDA_E_UNSUPPORTED, MSDA_E_IM2COL_STEP = -5, -6, -7


class msda_shape(C.Structure):
    _fields_ = [("n_frames", C.c_int64), ("spatial_size", C.c_int64), ("n_queries", C.c_int64),
                ("n_heads", C.c_int32), ("channels", C.c_int32), ("n_levels", C.c_int32),
                ("n_points", C.c_int32), ("value_dtype", C.c_int32), ("loc_dtype", C.c_int32)]


_P, _I64P, _SH = C.c_void_p, C.c_void_p, C.POINTER(msda_shape)

# name -> (restype, argtypes); every symbol include/msda_b200.h declares
PROTOTYPES = {
    "msda_abi_version": (C.c_int, []),
    "msda_error_string": (C.c_char_p, [C.c_int]),
    "msda_kernel_name": (C.c_char_p, [_SH, C.c_int]),
    "msda_algorithmic_bytes": (C.c_int64, [_SH, C.c_int]),
    "msda_check_im2col_step": (C.c_int, [C.c_int64, C.c_int]),
    "msda_launch_count": (C.c_int64, []),
    "msda_value_epilogue": (C.c_int, [C.c_int64, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "msda_topk_samples": (C.c_int, [C.POINTER(msda_shape), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p]),
    "msda_forward": (C.c_int, [_SH, _P, _I64P, _I64P, _P, _P, _P, _P]),
    "msda_backward_workspace_bytes": (C.c_size_t, [_SH]),
    "msda_backward": (C.c_int, [_SH, _P, _I64P, _I64P, _P, _P, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "msda_forward_3d": (C.c_int, [_SH, C.c_int64, _P, _I64P, _I64P, _P, _P, _P, _P]),
    "msda_backward_3d": (C.c_int, [_SH, C.c_int64, _P, _I64P, _I64P, _P, _P, _P, _P, _P, _P, _P, C.c_size_t, _P]),
    "msda_prologue_supported": (C.c_int, [_SH, C.c_int]),
    "msda_prologue_forward": (C.c_int, [_SH, _P, _I64P, _I64P, _P, _P, _P, C.c_int, _P, _P, _P, _P]),
    "msda_prologue_backward": (C.c_int, [_SH, _P, _I64P, _I64P, _P, _P, _P, C.c_int, _P, _P, _P, _P, _P, _P,
                                         C.c_size_t, _P]),
    "msda_host_ctx_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int]),
    "msda_host_ctx_destroy": (None, [C.c_void_p]),
    "msda_forward_host": (C.c_int, [C.c_void_p, _SH, _P, _I64P, _I64P, _P, _P, _P]),
    "msda_forward_backward_host": (C.c_int, [C.c_void_p, _SH, _P, _I64P, _I64P, _P, _P, _P, _P, _P, _P, _P]),
}

_lib = None


def library_path() -> Path:
    return capi_path()


def load() -> C.CDLL:
    """dlopen libmsda_b200.so and attach prototypes.  Fails loudly when the library is missing."""
    global _lib
    if _lib is not None:
        return _lib
    path = capi_path()
    if not path.exists():
        raise ImportError(f"{path} not built; run `python tce-rvos_b200/build.py` (no fallback exists)")
    lib = C.CDLL(str(path))
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)        # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class MsdaError(RuntimeError):
    def __init__(self, code: int, what: str):
        self.code = code
        msg = load().msda_error_string(code).decode()
        super().__init__(f"{what}: {msg} (code {code})")


def check(code: int, what: str) -> None:
    if code != 0:
        raise MsdaError(code, what)


_TORCH_DT = None


def dtype_code(dt) -> int:
    global _TORCH_DT
    import numpy as np
    try:
        import torch
        if _TORCH_DT is None:
            _TORCH_DT = {torch.float32: MSDA_F32, torch.bfloat16: MSDA_BF16, torch.float16: MSDA_F16,
                         torch.float64: MSDA_F64}
        if dt in _TORCH_DT:
            return _TORCH_DT[dt]
    except ImportError:
        pass
    return {np.dtype("float32"): MSDA_F32, np.dtype("float16"): MSDA_F16, np.dtype("float64"): MSDA_F64}[np.dtype(dt)]


def make_shape(N, S, Lq, M, D, L, P, value_dtype, loc_dtype) -> msda_shape:
    return msda_shape(int(N), int(S), int(Lq), int(M), int(D), int(L), int(P), int(value_dtype), int(loc_dtype))


def shape_of(value, sampling_loc, n_levels) -> msda_shape:
    N, S, M, D = value.shape
    Lq, P = sampling_loc.shape[1], sampling_loc.shape[4]
    return make_shape(N, S, Lq, M, D, n_levels, P, dtype_code(value.dtype), dtype_code(sampling_loc.dtype))


# ---- device API on torch tensors (raw pointers + the current stream) --------------------------

def _stream_ptr(t):
    import torch
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def forward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, out=None):
    """msda_forward on CUDA torch tensors through the C ABI (no torch binding involved)."""
    import torch
    lib = load()
    if not value.is_cuda:
        raise RuntimeError("Not implemented on the CPU")
    sh = shape_of(value, sampling_loc, spatial_shapes.shape[0])
    if out is None:
        out = torch.empty(value.shape[0], sampling_loc.shape[1], value.shape[2] * value.shape[3],
                          dtype=value.dtype, device=value.device)
    with torch.cuda.device(value.device):
        check(lib.msda_forward(C.byref(sh), value.data_ptr(), spatial_shapes.data_ptr(),
                               level_start_index.data_ptr(), sampling_loc.data_ptr(), attn_weight.data_ptr(),
                               out.data_ptr(), _stream_ptr(value)), "msda_forward")
    return out


def backward(value, spatial_shapes, level_start_index, sampling_loc, attn_weight, grad_out, use_workspace=True):
    """msda_backward on CUDA torch tensors through the C ABI."""
    import torch
    lib = load()
    if not value.is_cuda:
        raise RuntimeError("Not implemented on the CPU")
    sh = shape_of(value, sampling_loc, spatial_shapes.shape[0])
    gv = torch.empty_like(value)
    gl = torch.empty_like(sampling_loc)
    ga = torch.empty_like(attn_weight)
    ws_bytes = lib.msda_backward_workspace_bytes(C.byref(sh)) if use_workspace else 0
    ws = torch.empty(ws_bytes // 4, dtype=torch.float32, device=value.device) if ws_bytes else None
    with torch.cuda.device(value.device):
        check(lib.msda_backward(C.byref(sh), value.data_ptr(), spatial_shapes.data_ptr(),
                                level_start_index.data_ptr(), sampling_loc.data_ptr(), attn_weight.data_ptr(),
                                grad_out.data_ptr(), gv.data_ptr(), gl.data_ptr(), ga.data_ptr(),
                                ws.data_ptr() if ws is not None else None, ws_bytes, _stream_ptr(value)),
              "msda_backward")
    return gv, gl, ga


def value_epilogue(src, padding_mask=None, out_dtype=None, out=None):
    """msda_value_epilogue on CUDA torch tensors: src [..., width] (fp32 / bf16), padding_mask
    [...] bool or None -> masked (and converted) copy; out=src masks in place."""
    import torch
    lib = load()
    if not src.is_cuda:
        raise RuntimeError("Not implemented on the CPU")
    width = src.shape[-1]
    rows = src.numel() // width if width else 0
    out_dtype = out_dtype or src.dtype
    if out is None:
        out = torch.empty(src.shape, dtype=out_dtype, device=src.device)
    mask = None
    if padding_mask is not None:
        mask = padding_mask.contiguous().view(torch.uint8) if padding_mask.dtype == torch.bool else padding_mask.contiguous()
        assert mask.numel() == rows and mask.dtype == torch.uint8
    with torch.cuda.device(src.device):
        check(lib.msda_value_epilogue(rows, width, src.data_ptr(), dtype_code(src.dtype),
                                      mask.data_ptr() if mask is not None else None, out.data_ptr(),
                                      dtype_code(out.dtype), _stream_ptr(src)), "msda_value_epilogue")
    return out


def topk_samples(sampling_loc, attn_weight, valid_ratios=None, k=30, want_index=True):
    """msda_topk_samples on CUDA torch tensors -> (samples_keep [N,Lq,k,2], weights_keep [N,Lq,k],
    index_keep [N,Lq,k] int64 or None)."""
    import torch
    lib = load()
    if not sampling_loc.is_cuda:
        raise RuntimeError("Not implemented on the CPU")
    N, Lq, M, L, P, _ = sampling_loc.shape
    sh = make_shape(N, 1, Lq, M, 1, L, P, MSDA_F32, dtype_code(sampling_loc.dtype))
    samples = torch.empty(N, Lq, k, 2, dtype=sampling_loc.dtype, device=sampling_loc.device)
    weights = torch.empty(N, Lq, k, dtype=attn_weight.dtype, device=sampling_loc.device)
    index = torch.empty(N, Lq, k, dtype=torch.int64, device=sampling_loc.device) if want_index else None
    with torch.cuda.device(sampling_loc.device):
        check(lib.msda_topk_samples(C.byref(sh), int(k), sampling_loc.data_ptr(), attn_weight.data_ptr(),
                                    valid_ratios.data_ptr() if valid_ratios is not None else None,
                                    samples.data_ptr(), weights.data_ptr(),
                                    index.data_ptr() if index is not None else None, _stream_ptr(sampling_loc)),
              "msda_topk_samples")
    return samples, weights, index


# ---- host API on numpy arrays / pinned torch CPU tensors ---------------------------------------

class HostContext:
    """Owns device staging + streams for the msda_*_host entry points."""

    def __init__(self, device: int = 0):
        self._lib = load()
        self._ctx = C.c_void_p()
        check(self._lib.msda_host_ctx_create(C.byref(self._ctx), int(device)), "msda_host_ctx_create")

    def close(self):
        if self._ctx:
            self._lib.msda_host_ctx_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _ptr(a):
        if hasattr(a, "data_ptr"):
            return a.data_ptr()
        return a.ctypes.data

    def forward(self, shape: msda_shape, value, spatial_shapes, level_start_index, sampling_loc, attn_weight, out):
        check(self._lib.msda_forward_host(self._ctx, C.byref(shape), self._ptr(value), self._ptr(spatial_shapes),
                                          self._ptr(level_start_index), self._ptr(sampling_loc),
                                          self._ptr(attn_weight), self._ptr(out)), "msda_forward_host")
        return out

    def forward_backward(self, shape: msda_shape, value, spatial_shapes, level_start_index, sampling_loc,
                         attn_weight, grad_out, out, grad_value, grad_loc, grad_aw):
        check(self._lib.msda_forward_backward_host(
            self._ctx, C.byref(shape), self._ptr(value), self._ptr(spatial_shapes), self._ptr(level_start_index),
            self._ptr(sampling_loc), self._ptr(attn_weight), self._ptr(grad_out), self._ptr(out),
            self._ptr(grad_value), self._ptr(grad_loc), self._ptr(grad_aw)), "msda_forward_backward_host")
        return out, grad_value, grad_loc, grad_aw
