"""Drop-in for lib/modeling/roi_xfrom/roi_align/modules/roi_align.py:6-45."""
from torch.nn.functional import avg_pool2d, max_pool2d
from torch.nn.modules.module import Module

from ..functions.roi_align import RoIAlignFunction


class _Base(Module):
    def __init__(self, aligned_height, aligned_width, spatial_scale, sampling_ratio):
        super().__init__()
        self.aligned_width = int(aligned_width)
        self.aligned_height = int(aligned_height)
        self.spatial_scale = float(spatial_scale)
        self.sampling_ratio = int(sampling_ratio)

    def _pool(self, features, rois, extra):
        return RoIAlignFunction(self.aligned_height + extra, self.aligned_width + extra,
                                self.spatial_scale, self.sampling_ratio)(features, rois)


class RoIAlign(_Base):
    def forward(self, features, rois):
        return self._pool(features, rois, 0)


class RoIAlignAvg(_Base):
    """(h+1)x(w+1) RoIAlign followed by a 2x2 stride-1 average (modules/roi_align.py:20-32)."""

    def forward(self, features, rois):
        return avg_pool2d(self._pool(features, rois, 1), kernel_size=2, stride=1)


class RoIAlignMax(_Base):
    def forward(self, features, rois):
        return max_pool2d(self._pool(features, rois, 1), kernel_size=2, stride=1)
