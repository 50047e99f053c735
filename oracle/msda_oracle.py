"""Python side of the oracle.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Parity status: PINNED against the reference's own ms_deform_attn_core_pytorch through the golden
vectors in tests/golden/ (made by oracle/gen_golden.py, which imports the reference in the build
container; checked by tests/test_oracle.py).

Two independent restatements live here:

* ``core_pytorch``  — the reference's CPU path, models/ops/functions/ms_deform_attn_func.py:67-87:
  per level, a bilinear ``grid_sample`` (zeros padding, align_corners=False) of the level's value
  map at ``2*loc-1``, weighted by the attention weights and summed over levels x points.  Its
  backward is torch autograd, exactly how the reference's own test differentiates it.  This is the
  function bench.py times as the CPU baseline.
* ``c_forward`` / ``c_backward`` — ctypes calls into oracle/msda_oracle.c, which restates the CUDA
  kernels' arithmetic (gating, zero padding, gradient formulas; 2-D and the fork's 3-D variant).

Also here: the seeded input generators of SURVEY.md 8(d) and the parity metrics of 8(c).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "_build" / "libmsda_oracle.so"
_lib = None



# ------------------------------------------------------------------------------------------------
# restatement 1: the reference's pure-PyTorch path
# ------------------------------------------------------------------------------------------------
def core_pytorch(value, spatial_shapes, sampling_locations, attention_weights):
    """value [N,S,M,D], spatial_shapes [(H,W)...] (list or tensor), sampling_locations
    [N,Lq,M,L,P,2] in [0,1] (x,y), attention_weights [N,Lq,M,L,P]  ->  [N,Lq,M*D]."""
    N, S, M, D = value.shape
    _, Lq, _, L, P, _ = sampling_locations.shape
    sizes = [(int(h), int(w)) for h, w in (spatial_shapes.tolist() if torch.is_tensor(spatial_shapes) else spatial_shapes)]
    # heads become batch entries: [N*M, D, S]
    maps = value.permute(0, 2, 3, 1).reshape(N * M, D, S)
    # grid_sample wants [-1, 1]; [N*M, Lq, L, P, 2]
    grid = (2.0 * sampling_locations - 1.0).permute(0, 2, 1, 3, 4, 5).reshape(N * M, Lq, L, P, 2)
    weights = attention_weights.permute(0, 2, 1, 3, 4).reshape(N * M, 1, Lq, L, P)
    out = value.new_zeros(N * M, D, Lq)
    start = 0
    for lvl, (h, w) in enumerate(sizes):
        level_map = maps[:, :, start:start + h * w].reshape(N * M, D, h, w)
        start += h * w
        sampled = F.grid_sample(level_map, grid[:, :, lvl], mode="bilinear", padding_mode="zeros",
                                align_corners=False)            # [N*M, D, Lq, P]
        out = out + (sampled * weights[:, :, :, lvl]).sum(-1)
    return out.reshape(N, M * D, Lq).transpose(1, 2).contiguous()


def core_pytorch_fwd_bwd(value, spatial_shapes, loc, aw, grad_out):
    """forward + autograd backward through ``core_pytorch``; returns (out, g_value, g_loc, g_aw)."""
    v = value.detach().clone().requires_grad_(True)
    l = loc.detach().clone().requires_grad_(True)
    a = aw.detach().clone().requires_grad_(True)
    out = core_pytorch(v, spatial_shapes, l, a)
    gv, gl, ga = torch.autograd.grad(out, (v, l, a), grad_out)
    return out.detach(), gv, gl, ga


# ------------------------------------------------------------------------------------------------
# restatement 2: C (the CUDA kernels' arithmetic)
# ------------------------------------------------------------------------------------------------
def build_c_oracle(force: bool = False) -> Path:
    src = _HERE / "msda_oracle.c"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        res = subprocess.run(["make", "-C", str(_HERE)] + (["-B"] if force else []), capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("oracle build failed:\n" + res.stdout + res.stderr)
    return _LIB_PATH


def _load_c():
    global _lib
    if _lib is None:
        build_c_oracle()
        _lib = C.CDLL(str(_LIB_PATH))
    return _lib


def _np(t, dtype):
    if torch.is_tensor(t):
        t = t.detach().cpu().numpy()
    return np.ascontiguousarray(t, dtype=dtype)


def _dims(value, loc):
    N, S, M, D = value.shape
    _, Lq, _, L, P, LD = loc.shape
    return N, S, Lq, M, D, L, P, LD


def c_forward(value, spatial_shapes, level_start_index, loc, aw, dtype=np.float64, group=0):
    lib = _load_c()
    value, loc, aw = _np(value, dtype), _np(loc, dtype), _np(aw, dtype)
    shapes, lsi = _np(spatial_shapes, np.int64), _np(level_start_index, np.int64)
    N, S, Lq, M, D, L, P, LD = _dims(value, loc)
    out = np.empty((N, Lq, M * D), dtype=dtype)
    fn = lib.msda_oracle_forward_f64 if dtype == np.float64 else lib.msda_oracle_forward_f32
    fn.restype = None
    fn(value.ctypes.data_as(C.c_void_p), shapes.ctypes.data_as(C.c_void_p), lsi.ctypes.data_as(C.c_void_p),
       loc.ctypes.data_as(C.c_void_p), aw.ctypes.data_as(C.c_void_p),
       C.c_int64(N), C.c_int64(S), C.c_int64(Lq), C.c_int(M), C.c_int(D), C.c_int(L), C.c_int(P),
       C.c_int(1 if LD == 3 else 0), C.c_int64(group), out.ctypes.data_as(C.c_void_p))
    return out


def c_backward(value, spatial_shapes, level_start_index, loc, aw, grad_out, dtype=np.float64, group=0):
    lib = _load_c()
    value, loc, aw, grad_out = _np(value, dtype), _np(loc, dtype), _np(aw, dtype), _np(grad_out, dtype)
    shapes, lsi = _np(spatial_shapes, np.int64), _np(level_start_index, np.int64)
    N, S, Lq, M, D, L, P, LD = _dims(value, loc)
    gv, gl, ga = np.empty_like(value), np.empty_like(loc), np.empty_like(aw)
    fn = lib.msda_oracle_backward_f64 if dtype == np.float64 else lib.msda_oracle_backward_f32
    fn.restype = None
    fn(value.ctypes.data_as(C.c_void_p), shapes.ctypes.data_as(C.c_void_p), lsi.ctypes.data_as(C.c_void_p),
       loc.ctypes.data_as(C.c_void_p), aw.ctypes.data_as(C.c_void_p), grad_out.ctypes.data_as(C.c_void_p),
       C.c_int64(N), C.c_int64(S), C.c_int64(Lq), C.c_int(M), C.c_int(D), C.c_int(L), C.c_int(P),
       C.c_int(1 if LD == 3 else 0), C.c_int64(group),
       gv.ctypes.data_as(C.c_void_p), gl.ctypes.data_as(C.c_void_p), ga.ctypes.data_as(C.c_void_p))
    return gv, gl, ga


# ------------------------------------------------------------------------------------------------
# module prologue (models/ops/modules/ms_deform_attn.py:99-109), for the fused-kernel tests
# ------------------------------------------------------------------------------------------------
def prologue(offsets, logits, reference_points, spatial_shapes):
    """offsets [N,Lq,M,L,P,2], logits [N,Lq,M,L*P], reference_points [N,Lq,L,2|4] ->
    (sampling_locations [N,Lq,M,L,P,2], attention_weights [N,Lq,M,L,P])."""
    N, Lq, M, L, P, _ = offsets.shape
    aw = F.softmax(logits, -1).view(N, Lq, M, L, P)
    shapes = torch.as_tensor(spatial_shapes, dtype=offsets.dtype, device=offsets.device)
    if reference_points.shape[-1] == 2:
        normalizer = torch.stack([shapes[:, 1], shapes[:, 0]], -1)
        loc = reference_points[:, :, None, :, None, :] + offsets / normalizer[None, None, None, :, None, :]
    else:
        loc = reference_points[:, :, None, :, None, :2] + offsets / P * reference_points[:, :, None, :, None, 2:] * 0.5
    return loc, aw


# ------------------------------------------------------------------------------------------------
# rows f2 / f3: the passes either side of the op, restated with the reference's own torch calls
# ------------------------------------------------------------------------------------------------
def value_epilogue(value, padding_mask=None, out_dtype=None):
    """models/ops/modules/ms_deform_attn.py:96-97 (+ the optional down-cast):
    ``value.masked_fill(input_padding_mask[..., None], float(0))``."""
    if padding_mask is not None:
        value = value.masked_fill(padding_mask[..., None], float(0))
    return value.to(out_dtype) if out_dtype is not None else value


def decoder_sample_topk(sampling_locations, attention_weights, valid_ratios=None, k=30):
    """models/tce_deformable_transformer.py:752-758, line for line, except that the selection is a
    STABLE descending sort (ties -> lower index) where the reference calls topk, whose order among
    equal weights is unspecified.  Returns (samples_keep, weights_keep, top_idx)."""
    N, Len_q = sampling_locations.shape[:2]
    if valid_ratios is not None:
        sampling_locations = sampling_locations / valid_ratios[:, None, None, :, None, :]
    weights_flat = attention_weights.reshape(N, Len_q, -1)
    samples_flat = sampling_locations.reshape(N, Len_q, -1, 2)
    top_idx = torch.sort(weights_flat.float(), dim=2, descending=True, stable=True)[1][..., :k]
    weights_keep = torch.gather(weights_flat, 2, top_idx)
    samples_keep = torch.gather(samples_flat, 2, top_idx.unsqueeze(-1).repeat(1, 1, 1, 2))
    return samples_keep, weights_keep, top_idx


def decoder_sample_topk_reference_calls(sampling_locations, attention_weights, valid_ratios, k=30):
    """The same lines with the reference's literal torch.topk call (ties unspecified); used to pin the
    restatement above wherever the k-th weight is unique."""
    N, Len_q = sampling_locations.shape[:2]
    sampling_locations = sampling_locations / valid_ratios[:, None, None, :, None, :]
    weights_flat = attention_weights.view(N, Len_q, -1)
    samples_flat = sampling_locations.view(N, Len_q, -1, 2)
    top_weights, top_idx = weights_flat.topk(k, dim=2)
    weights_keep = torch.gather(weights_flat, 2, top_idx)
    samples_keep = torch.gather(samples_flat, 2, top_idx.unsqueeze(-1).repeat(1, 1, 1, 2))
    return samples_keep, weights_keep, top_idx


# ------------------------------------------------------------------------------------------------
# seeded synthetic inputs: defined in tools/synthetic.py (shared with the benchmarks, which
# must not import the checker); re-exported here for the tests
# ------------------------------------------------------------------------------------------------
from tools.synthetic import (LEVELS_A, LEVELS_B, LEVELS_C, LEVELS_TEST_PY, level_tensors,  # noqa: E402,F401
                             make_inputs, reference_points_grid)


# ------------------------------------------------------------------------------------------------
# parity metrics
# ------------------------------------------------------------------------------------------------
def max_rel(a, ref):
    """max |a-ref| / max|ref| — error relative to the tensor's scale (elementwise relative error is
    meaningless where outputs cancel to ~0)."""
    a, ref = torch.as_tensor(a).double(), torch.as_tensor(ref).double()
    scale = ref.abs().max().clamp_min(1e-300)
    return float(((a - ref).abs().max() / scale))


def assert_close(a, ref, tol, what=""):
    """The parity criterion of this repo, written out: the error of EVERY element is held against the tensor's scale,
    |a - ref| <= tol * max|ref|  (= max_rel(a, ref) <= tol), which is also torch.allclose(a, ref, rtol=tol,
    atol=tol * max|ref|) — asserted as well so that the elementwise form is on record.  A purely elementwise relative
    bound is meaningless for this operator: outputs and gradients are signed sums that cancel to ~0 for some elements."""
    a, ref = torch.as_tensor(a).double(), torch.as_tensor(ref).double()
    scale = float(ref.abs().max().clamp_min(1e-300))
    err = float((a - ref).abs().max()) / scale
    assert err <= tol, (what, "max|a-b|/max|b|", err, "tol", tol)
    assert torch.allclose(a, ref, rtol=tol, atol=tol * scale), (what, "allclose(rtol=tol, atol=tol*scale)")
    return err


def rel_l2(a, ref):
    a, ref = torch.as_tensor(a).double(), torch.as_tensor(ref).double()
    return float((a - ref).norm() / ref.norm().clamp_min(1e-300))


def boundary_mask(loc, levels, eps_px=1e-4):
    """True where a sampling point lies within eps_px of a bilinear cell boundary in x or y (or of
    the -1 / H gating edges): grad_sampling_loc is discontinuous there, so fp32 and fp64 can pick
    different cells (SURVEY.md 8c caveat).  loc [N,Lq,M,L,P,2|3] -> bool [N,Lq,M,L,P]."""
    loc = torch.as_tensor(loc).double()
    wh = torch.tensor([[w, h] for h, w in levels], dtype=torch.float64)
    px = loc[..., :2] * wh[None, None, None, :, None, :] - 0.5
    frac = px - torch.floor(px)
    near = (frac < eps_px) | (frac > 1 - eps_px)
    return near.any(-1)


def grad_loc_errors(g, ref, loc, levels, eps_px=1e-4):
    """(max_rel over non-boundary points, rel_l2 over non-boundary points, boundary fraction).
    Points within eps_px of a bilinear cell boundary are excluded from both figures: there the
    gradient jumps, and a last-bit difference in loc*W-0.5 selects the other cell."""
    g, ref = torch.as_tensor(g).double(), torch.as_tensor(ref).double()
    mask = boundary_mask(loc, levels, eps_px)
    keep = (~mask)[..., None].expand_as(ref[..., :2])
    if ref.shape[-1] == 3:
        keep = torch.cat([keep, (~mask)[..., None]], -1)
    scale = ref.abs().max().clamp_min(1e-300)
    diff = (g - ref) * keep
    err = diff.abs().max() / scale
    l2 = diff.norm() / (ref * keep).norm().clamp_min(1e-300)
    return float(err), float(l2), float(mask.double().mean())
