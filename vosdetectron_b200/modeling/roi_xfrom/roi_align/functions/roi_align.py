"""Drop-in for lib/modeling/roi_xfrom/roi_align/functions/roi_align.py:7-48.

``RoIAlignFunction(aligned_height, aligned_width, spatial_scale, sampling_ratio)(features, rois)``
keeps the reference's constructor-then-call form, but is a plain callable that
forwards to a static ``torch.autograd.Function`` (legacy instance-style Functions
raise on current PyTorch).  Differentiable w.r.t. ``features`` only; CPU tensors
raise NotImplementedError like the reference (:29-30).
"""
import torch
from torch.autograd import Function

from ..... import ops


class _RoIAlign(Function):
    @staticmethod
    def forward(ctx, features, rois, aligned_height, aligned_width, spatial_scale, sampling_ratio):
        ctx.params = (aligned_height, aligned_width, spatial_scale, sampling_ratio)
        ctx.feature_size = tuple(features.shape)
        ctx.channels_last = ops._is_channels_last(features)      # the gradient comes back in the map's memory order
        ctx.save_for_backward(rois)
        return ops.roi_align_forward(features, rois, aligned_height, aligned_width, spatial_scale,
                                     sampling_ratio)

    @staticmethod
    def backward(ctx, grad_output):
        (rois,) = ctx.saved_tensors
        ah, aw, scale, sr = ctx.params
        if not grad_output.is_cuda:
            raise NotImplementedError("RoIAlign backward needs CUDA tensors")
        grad_input = ops.roi_align_backward(grad_output.contiguous(), rois, ctx.feature_size, ah, aw, scale, sr,
                                            channels_last=ctx.channels_last)
        return grad_input, None, None, None, None, None


class RoIAlignFunction(object):
    def __init__(self, aligned_height, aligned_width, spatial_scale, sampling_ratio):
        self.aligned_width = int(aligned_width)
        self.aligned_height = int(aligned_height)
        self.spatial_scale = float(spatial_scale)
        self.sampling_ratio = int(sampling_ratio)

    def __call__(self, features, rois):
        if not features.is_cuda:
            raise NotImplementedError
        return _RoIAlign.apply(features, rois.detach(), self.aligned_height, self.aligned_width,
                               self.spatial_scale, self.sampling_ratio)

    forward = __call__


class _RoIAlignML(Function):
    """All FPN levels in one launch (fast variant used by roi_feature_transform)."""

    @staticmethod
    def forward(ctx, rois, roi_level, out_index, aligned_height, aligned_width, sampling_ratio, scales,
                *level_features):
        ctx.params = (aligned_height, aligned_width, sampling_ratio, tuple(scales))
        ctx.shapes = [tuple(f.shape) for f in level_features]
        ctx.channels_last = len(level_features) > 0 and all(ops._is_channels_last(f) for f in level_features)
        ctx.save_for_backward(rois, roi_level, out_index if out_index is not None else torch.empty(0))
        ctx.has_index = out_index is not None
        return ops.roi_align_ml_forward(level_features, scales, rois, roi_level, aligned_height,
                                        aligned_width, sampling_ratio, out_index)

    @staticmethod
    def backward(ctx, grad_output):
        rois, roi_level, out_index = ctx.saved_tensors
        ah, aw, sr, scales = ctx.params
        grads = ops.roi_align_ml_backward(grad_output.contiguous(), ctx.shapes, scales, rois, roi_level,
                                          ah, aw, sr, out_index if ctx.has_index else None,
                                          channels_last=ctx.channels_last)
        return (None, None, None, None, None, None, None) + tuple(grads)


def roi_align_multilevel(level_features, level_scales, rois, roi_level, aligned_height, aligned_width,
                         sampling_ratio, out_index=None):
    return _RoIAlignML.apply(rois.detach(), roi_level, out_index, int(aligned_height), int(aligned_width),
                             int(sampling_ratio), tuple(float(s) for s in level_scales), *level_features)
