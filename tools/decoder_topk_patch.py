"""Harness-side wiring of the fused decoder sample selection.

The reference's decoder loop (models/tce_deformable_transformer.py:718-790) selects, per layer, the 30
highest-weighted sampling points of every query with five tensor ops (:752-758: divide by the valid
ratios, topk, gather, repeat, gather).  `decoder_sample_topk` of this repo does that in one launch.  The
loop is a method body, so there is no seam to swap the five ops without touching the reference file;
this module therefore supplies a replacement `forward` for `DeformableTransformerDecoder` — installed by
tools/model_dropin.py --fused-topk on the class object at run time, the reference's files stay unedited —
that walks the layers the same way (reference-point scaling :739-744, layer call :745-746, box
refinement :761-771, intermediates :773-776) and calls the fused selection where the reference has :752-758.
"""
import importlib

import torch


def fused_decoder_forward(self, tgt, reference_points, src, src_spatial_shapes, src_level_start_index,
                          src_valid_ratios, query_pos=None, src_padding_mask=None):
    pkg = importlib.import_module("tce-rvos_b200")
    _inverse_sigmoid = importlib.import_module("util.misc").inverse_sigmoid      # the reference's own (util/misc.py:555)
    hidden, refs, kept = [], [], []
    out = tgt
    samples_keep = None
    for lid, layer in enumerate(self.layers):
        ratios = src_valid_ratios if reference_points.shape[-1] == 2 else torch.cat([src_valid_ratios, src_valid_ratios], -1)
        out, sampling_locations, attention_weights = layer(out, query_pos, reference_points[:, :, None] * ratios[:, None],
                                                           src, src_spatial_shapes, src_level_start_index, src_padding_mask)
        samples_keep, _, _ = pkg.decoder_sample_topk(sampling_locations, attention_weights, src_valid_ratios, 30)
        if self.bbox_embed is not None:
            delta = self.bbox_embed[lid](out)
            if reference_points.shape[-1] == 4:
                new_ref = (delta + _inverse_sigmoid(reference_points)).sigmoid()
            else:
                new_ref = delta
                new_ref[..., :2] = delta[..., :2] + _inverse_sigmoid(reference_points)
                new_ref = new_ref.sigmoid()
            reference_points = new_ref.detach()
        if self.return_intermediate:
            hidden.append(out)
            refs.append(reference_points)
            kept.append(samples_keep)
    if self.return_intermediate:
        return torch.stack(hidden), torch.stack(refs), torch.stack(kept)
    return out, reference_points, samples_keep


def install():
    """Replace DeformableTransformerDecoder.forward on the imported reference module; returns the original."""
    mod = importlib.import_module("models.tce_deformable_transformer")
    original = mod.DeformableTransformerDecoder.forward
    mod.DeformableTransformerDecoder.forward = fused_decoder_forward
    return original


def uninstall(original):
    mod = importlib.import_module("models.tce_deformable_transformer")
    mod.DeformableTransformerDecoder.forward = original
