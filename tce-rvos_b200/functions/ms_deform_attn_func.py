"""autograd Functions over the B200 extension.

Host-side mirror of the reference's models/ops/functions/ms_deform_attn_func.py:
  MSDeformAttnFunction    :21-41  same apply() signature, same 6 (or 7) returned gradients
  MSDeformAttnFunction3D  :43-64  the fork's cross-frame variant
The reference's third member, ms_deform_attn_core_pytorch (:67-87, "for debug and test only"),
is NOT shipped in the product package: it is test infrastructure and lives in oracle/.

New here: MSDeformAttnPrologueFunction, which fuses the module's softmax + location prologue
(models/ops/modules/ms_deform_attn.py:99-109) into the gather kernel; MSDeformAttnValueEpilogue, the
masked_fill (+ optional bf16 down-cast) that follows value_proj (:95-98) as one pass; and
decoder_sample_topk, the decoder's top-30 sample selection (models/tce_deformable_transformer.py:752-758)
as one launch.
"""
from __future__ import annotations

import torch
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from .._ext import load as _load_ext


class MSDeformAttnFunction(Function):
    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations,
                attention_weights, im2col_step, is_3d=False):
        MSDA = _load_ext()
        ctx.im2col_step = im2col_step
        output = MSDA.ms_deform_attn_forward(
            value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights,
            ctx.im2col_step, is_3d)
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index, sampling_locations,
                              attention_weights)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        MSDA = _load_ext()
        value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights = \
            ctx.saved_tensors
        # the reference hard-codes is_3d=False here (ms_deform_attn_func.py:37-39)
        grad_value, grad_sampling_loc, grad_attn_weight = MSDA.ms_deform_attn_backward(
            value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights,
            grad_output.contiguous(), ctx.im2col_step, False)
        # forward() is normally applied with 6 arguments (ms_deform_attn.py:113-114); autograd also
        # accepts a trailing None for an explicitly passed is_3d.
        return grad_value, None, None, grad_sampling_loc, grad_attn_weight, None, None


class MSDeformAttnFunction3D(Function):
    @staticmethod
    def forward(ctx, value, value_spatial_shapes, value_level_start_index, sampling_locations,
                attention_weights, im2col_step, is_3d=True):
        MSDA = _load_ext()
        ctx.im2col_step = im2col_step
        ctx.is_3d = is_3d
        output = MSDA.ms_deform_attn_forward(
            value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights,
            ctx.im2col_step, ctx.is_3d)
        ctx.save_for_backward(value, value_spatial_shapes, value_level_start_index, sampling_locations,
                              attention_weights)
        return output

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        MSDA = _load_ext()
        value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights = \
            ctx.saved_tensors
        grad_value, grad_sampling_loc, grad_attn_weight = MSDA.ms_deform_attn_backward(
            value, value_spatial_shapes, value_level_start_index, sampling_locations, attention_weights,
            grad_output.contiguous(), ctx.im2col_step, ctx.is_3d)
        return grad_value, None, None, grad_sampling_loc, grad_attn_weight, None, None


class MSDeformAttnPrologueFunction(Function):
    """out (and optionally the materialised sampling_locations / attention_weights) from the RAW
    Linear outputs: sampling offsets [N,Lq,M,L,P,2], logits [N,Lq,M,L*P], reference points
    [N,Lq,L,2|4].  The materialised tensors are returned for inspection (the decoder's top-k sample
    selection, tce_deformable_transformer.py:752-758) and are not differentiable."""

    @staticmethod
    def forward(ctx, value, spatial_shapes, level_start_index, offsets, logits, reference_points,
                materialize=False):
        MSDA = _load_ext()
        outs = MSDA.ms_deform_attn_prologue_forward(
            value, spatial_shapes, level_start_index, offsets, logits, reference_points, bool(materialize))
        ctx.save_for_backward(value, spatial_shapes, level_start_index, offsets, logits, reference_points)
        ctx.materialize = bool(materialize)
        if materialize:
            out, loc, aw = outs
            ctx.mark_non_differentiable(loc, aw)
            return out, loc, aw
        return outs[0]

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output, *unused):
        MSDA = _load_ext()
        value, spatial_shapes, level_start_index, offsets, logits, reference_points = ctx.saved_tensors
        need_ref = ctx.needs_input_grad[5]
        grads = MSDA.ms_deform_attn_prologue_backward(
            value, spatial_shapes, level_start_index, offsets, logits, reference_points,
            grad_output.contiguous(), bool(need_ref))
        grad_ref = grads[3] if need_ref else None
        return grads[0], None, None, grads[1], grads[2], grad_ref, None


class MSDeformAttnValueEpilogue(Function):
    """value_proj output -> value the sampling kernels read: rows under the padding mask become zero
    (models/ops/modules/ms_deform_attn.py:96-97) and the tensor is optionally stored as bf16, in ONE
    pass.  Backward is the same pass on the incoming gradient (same mask, conversion back)."""

    @staticmethod
    def forward(ctx, value, padding_mask=None, out_dtype=None):
        MSDA = _load_ext()
        out_dtype = out_dtype or value.dtype
        ctx.in_dtype = value.dtype
        ctx.save_for_backward(padding_mask if padding_mask is not None else torch.empty(0, device=value.device))
        ctx.has_mask = padding_mask is not None
        return MSDA.ms_deform_attn_value_epilogue(value.contiguous(), padding_mask, out_dtype, False)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        MSDA = _load_ext()
        (mask,) = ctx.saved_tensors
        grad = MSDA.ms_deform_attn_value_epilogue(grad_output.contiguous(), mask if ctx.has_mask else None,
                                                  ctx.in_dtype, False)
        return grad, None, None


def decoder_sample_topk(sampling_locations, attention_weights, valid_ratios=None, k=30):
    """(samples_keep [N,Lq,k,2], weights_keep [N,Lq,k], top_idx [N,Lq,k]) as computed by the decoder
    loop of the reference (models/tce_deformable_transformer.py:752-758):

        sampling_locations / valid_ratios[:, None, None, :, None, :]
        top_weights, top_idx = weights_flat.topk(k, dim=2); gather(weights), gather(samples)

    One launch; not differentiable (the reference never back-propagates through these tensors:
    they leave the transformer as `inter_samples` and are dropped, models/tce_rvos.py:312).  Equal
    weights are ordered by the lower (head, level, point) index."""
    MSDA = _load_ext()
    with torch.no_grad():
        return tuple(MSDA.ms_deform_attn_topk_samples(sampling_locations.contiguous(), attention_weights.contiguous(),
                                                      valid_ratios, int(k)))
