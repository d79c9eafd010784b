"""``nn.Module`` front ends of the RoIAlign op: drop-ins for the three classes of
lib/modeling/roi_xfrom/roi_align/modules/roi_align.py:6-45 (same constructor arguments, same outputs).

    RoIAlign(h, w, scale, sr)(features, rois)      -> (R, C, h, w)
    RoIAlignAvg(h, w, scale, sr)(features, rois)   -> (R, C, h, w): pooled at (h+1) x (w+1), then a 2x2 stride-1 mean
    RoIAlignMax(h, w, scale, sr)(features, rois)   -> (R, C, h, w): pooled at (h+1) x (w+1), then a 2x2 stride-1 max

All three run the CUDA RoIAlign of this package (NCHW or channels-last features, see ops.roi_align_forward);
the 2x2 reductions of the Avg / Max variants are plain torch pooling on the pooled (tiny) blob.
"""
import torch.nn.functional as F
from torch import nn

from ..functions.roi_align import RoIAlignFunction


class _PooledRoIs(nn.Module):
    """Holds the four hyper-parameters; ``pooled(features, rois, grow)`` samples a (h+grow) x (w+grow) grid."""

    def __init__(self, aligned_height, aligned_width, spatial_scale, sampling_ratio):
        super().__init__()
        self.aligned_height, self.aligned_width = int(aligned_height), int(aligned_width)
        self.spatial_scale, self.sampling_ratio = float(spatial_scale), int(sampling_ratio)

    def pooled(self, features, rois, grow=0):
        op = RoIAlignFunction(self.aligned_height + grow, self.aligned_width + grow, self.spatial_scale,
                              self.sampling_ratio)
        return op(features, rois)

    def extra_repr(self):
        return "%dx%d, scale=%g, sampling_ratio=%d" % (self.aligned_height, self.aligned_width, self.spatial_scale,
                                                       self.sampling_ratio)


class RoIAlign(_PooledRoIs):
    def forward(self, features, rois):
        return self.pooled(features, rois)


class RoIAlignAvg(_PooledRoIs):
    def forward(self, features, rois):
        return F.avg_pool2d(self.pooled(features, rois, grow=1), kernel_size=2, stride=1)


class RoIAlignMax(_PooledRoIs):
    def forward(self, features, rois):
        return F.max_pool2d(self.pooled(features, rois, grow=1), kernel_size=2, stride=1)
