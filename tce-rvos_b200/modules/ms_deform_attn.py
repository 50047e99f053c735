"""MSDeformAttn module — host-side mirror of the reference's
models/ops/modules/ms_deform_attn.py:31-117.

Same constructor, parameter names (state_dict keys sampling_offsets / attention_weights /
value_proj / output_proj .{weight,bias}, :56-59), initialisation (:63-77), forward signature (:79)
and fork-specific 3-tuple return ``(output, sampling_locations, attention_weights)`` (:117), so a
reference checkpoint loads and the reference's transformer runs unchanged.

What differs:
  * the core op is the B200 extension (no CPU path);
  * the `assert ... == Len_in` of :93 forces a device->host sync on every call; here it only runs
    when ``check_shapes`` is True (default: once per distinct shape tensor, then cached);
  * ``fused_prologue`` (default False = reference composition): when True the softmax and the
    reference-point + offset arithmetic of :100-109 are computed inside the gather kernel from the
    raw Linear outputs.  The returned sampling_locations / attention_weights are then materialised
    by the kernel only when ``return_sampling`` is True (the decoder needs them,
    tce_deformable_transformer.py:752-758; encoder and FTF discard them) and carry no gradient.
  * ``fused_value_epilogue`` (default False): the masked_fill of :96-97 runs as one pass of the
    extension instead of an out-of-place torch op, and with ``value_dtype = torch.bfloat16`` the same
    pass stores the value as bf16 so that the 16-bit gather kernels are used (bf16 value and output,
    fp32 locations and weights; 1e-2 accuracy class).
"""
from __future__ import annotations

import math
import weakref
import warnings

import torch
import torch.nn.functional as F
from torch import nn
from torch.nn.init import constant_, xavier_uniform_

from ..functions import MSDeformAttnFunction, MSDeformAttnPrologueFunction, MSDeformAttnValueEpilogue
from .._ext import load as _load_ext

_CHECKED_SHAPES = {}      # (id, version, Len_in) -> weakref of a validated spatial_shapes tensor (check_shapes='once')


def _is_power_of_2(n):
    if (not isinstance(n, int)) or (n < 0):
        raise ValueError("invalid input for _is_power_of_2: {} (type: {})".format(n, type(n)))
    return (n & (n - 1) == 0) and n != 0


class MSDeformAttn(nn.Module):
    def __init__(self, d_model=256, n_levels=4, n_heads=8, n_points=4):
        super().__init__()
        if d_model % n_heads != 0:
            raise ValueError('d_model must be divisible by n_heads, but got {} and {}'.format(d_model, n_heads))
        _d_per_head = d_model // n_heads
        if not _is_power_of_2(_d_per_head):
            warnings.warn("You'd better set d_model in MSDeformAttn to make the dimension of each attention head "
                          "a power of 2 which is more efficient in our CUDA implementation.")

        self.im2col_step = 64   # accepted and validated for compatibility; the kernels never chunk

        self.d_model = d_model
        self.n_levels = n_levels
        self.n_heads = n_heads
        self.n_points = n_points

        self.sampling_offsets = nn.Linear(d_model, n_heads * n_levels * n_points * 2)
        self.attention_weights = nn.Linear(d_model, n_heads * n_levels * n_points)
        self.value_proj = nn.Linear(d_model, d_model)
        self.output_proj = nn.Linear(d_model, d_model)

        # B200 additions (not parameters, not in state_dict)
        self.fused_prologue = False
        self.return_sampling = True
        self.fused_value_epilogue = False
        self.value_dtype = None             # None = keep value_proj's dtype | torch.bfloat16
        self.check_shapes = "once"          # True | False | "once"

        self._reset_parameters()

    def _reset_parameters(self):
        constant_(self.sampling_offsets.weight.data, 0.)
        thetas = torch.arange(self.n_heads, dtype=torch.float32) * (2.0 * math.pi / self.n_heads)
        grid_init = torch.stack([thetas.cos(), thetas.sin()], -1)
        grid_init = (grid_init / grid_init.abs().max(-1, keepdim=True)[0]).view(self.n_heads, 1, 1, 2) \
            .repeat(1, self.n_levels, self.n_points, 1)
        for i in range(self.n_points):
            grid_init[:, :, i, :] *= i + 1
        with torch.no_grad():
            self.sampling_offsets.bias = nn.Parameter(grid_init.view(-1))
        constant_(self.attention_weights.weight.data, 0.)
        constant_(self.attention_weights.bias.data, 0.)
        xavier_uniform_(self.value_proj.weight.data)
        constant_(self.value_proj.bias.data, 0.)
        xavier_uniform_(self.output_proj.weight.data)
        constant_(self.output_proj.bias.data, 0.)

    def _validate_shapes(self, input_spatial_shapes, Len_in):
        if self.check_shapes is False:
            return
        if self.check_shapes == "once":
            # keyed on the tensor OBJECT and its version counter, not on its address: the caching allocator hands
            # the same pointer to a later, different shape table, which must be checked again (the kernels trust
            # H*W and level_start_index).  The model builds the table once per forward and passes the same tensor
            # to every layer (tce_deformable_transformer.py:173-181); the cache is shared by all MSDeformAttn
            # instances, so one sync per forward remains.
            key = (id(input_spatial_shapes), input_spatial_shapes._version, int(Len_in))
            hit = _CHECKED_SHAPES.get(key)
            if hit is not None and hit() is input_spatial_shapes:
                return
            if len(_CHECKED_SHAPES) > 64:
                _CHECKED_SHAPES.clear()
            _CHECKED_SHAPES[key] = weakref.ref(input_spatial_shapes)
        assert (input_spatial_shapes[:, 0] * input_spatial_shapes[:, 1]).sum() == Len_in

    def forward(self, query, reference_points, input_flatten, input_spatial_shapes, input_level_start_index,
                input_padding_mask=None):
        """Arguments and return value exactly as the reference (:79-117):
        query (N, Lq, C); reference_points (N, Lq, n_levels, 2|4); input_flatten (N, S, C);
        input_spatial_shapes (n_levels, 2); input_level_start_index (n_levels,);
        input_padding_mask (N, S) bool or None  ->  (output (N, Lq, C), sampling_locations,
        attention_weights)."""
        N, Len_q, _ = query.shape
        N, Len_in, _ = input_flatten.shape
        self._validate_shapes(input_spatial_shapes, Len_in)

        value = self.value_proj(input_flatten)
        if self.fused_value_epilogue and value.is_cuda and (input_padding_mask is not None or self.value_dtype is not None):
            value = MSDeformAttnValueEpilogue.apply(value, input_padding_mask, self.value_dtype)
        elif input_padding_mask is not None:
            value = value.masked_fill(input_padding_mask[..., None], float(0))
        value = value.view(N, Len_in, self.n_heads, self.d_model // self.n_heads)
        sampling_offsets = self.sampling_offsets(query).view(
            N, Len_q, self.n_heads, self.n_levels, self.n_points, 2)
        attention_weights = self.attention_weights(query).view(
            N, Len_q, self.n_heads, self.n_levels * self.n_points)
        if reference_points.shape[-1] not in (2, 4):
            raise ValueError('Last dim of reference_points must be 2 or 4, but get {} instead.'
                             .format(reference_points.shape[-1]))

        if self.fused_prologue and value.is_cuda:
            ref = reference_points.to(sampling_offsets.dtype).contiguous()
            if _load_ext().ms_deform_attn_prologue_supported(value, input_spatial_shapes, sampling_offsets, ref):
                res = MSDeformAttnPrologueFunction.apply(
                    value.contiguous(), input_spatial_shapes, input_level_start_index,
                    sampling_offsets.contiguous(), attention_weights.contiguous(), ref,
                    bool(self.return_sampling))
                if self.return_sampling:
                    output, sampling_locations, attention_weights = res
                else:
                    output, sampling_locations, attention_weights = res, None, None
                return self.output_proj(output.to(query.dtype)), sampling_locations, attention_weights

        attention_weights = F.softmax(attention_weights, -1).view(
            N, Len_q, self.n_heads, self.n_levels, self.n_points)
        if reference_points.shape[-1] == 2:
            offset_normalizer = torch.stack([input_spatial_shapes[..., 1], input_spatial_shapes[..., 0]], -1)
            sampling_locations = reference_points[:, :, None, :, None, :] \
                + sampling_offsets / offset_normalizer[None, None, None, :, None, :]
        else:
            sampling_locations = reference_points[:, :, None, :, None, :2] \
                + sampling_offsets / self.n_points * reference_points[:, :, None, :, None, 2:] * 0.5
        output = MSDeformAttnFunction.apply(
            value, input_spatial_shapes, input_level_start_index, sampling_locations, attention_weights,
            self.im2col_step)
        output = self.output_proj(output.to(query.dtype))
        return output, sampling_locations, attention_weights
