"""``FlowAlign(spatial_scale)`` module: drop-in for lib_vos/vos_model/flow_align/modules/flow_align.py:5-38.

The optical flow arrives at image resolution; a feature map at ``spatial_scale`` needs it at its own resolution and
in its own pixel units.  The reference does that with a frozen 2 -> 2 convolution whose kernel and stride are
``k = 1 / spatial_scale`` and whose diagonal weights are ``spatial_scale ** 3`` (= box mean over k x k, times
``spatial_scale`` to rescale the displacement).  The layer is kept under the same attribute name and with the same
weights so reference checkpoints load unchanged; the warp itself is the CUDA FlowAlign of this package.
"""
import torch
from torch import nn

from ..functions.flow_align import FlowAlignFunction

_SCALES = (1.0, 0.5, 0.25, 0.125, 0.0625, 0.03125, 1.0 / 64.0)


def _frozen_flow_pool(spatial_scale):
    """2 -> 2 strided convolution equal to ``spatial_scale * avg_pool2d(flow, 1 / spatial_scale)``."""
    if spatial_scale not in _SCALES:
        raise AssertionError("spatial_scale must be one of %s" % (_SCALES,))
    k = int(1.0 / spatial_scale)
    conv = nn.Conv2d(2, 2, kernel_size=k, stride=k, bias=False)
    with torch.no_grad():
        conv.weight.zero_()
        for ch in (0, 1):                      # x displacement -> x, y displacement -> y
            conv.weight[ch, ch].fill_(spatial_scale ** 3)
    conv.weight.requires_grad_(False)          # "This layer should not be trained." (:26-28)
    return conv


class FlowAlign(nn.Module):
    def __init__(self, spatial_scale):
        super().__init__()
        self.spatial_scale = spatial_scale
        self.feature_size = None               # attribute of the reference module (unused there as well)
        self.conv_flow_downsample = _frozen_flow_pool(spatial_scale)

    def forward(self, features, flows):
        level_flow = flows if self.spatial_scale == 1.0 else self.conv_flow_downsample(flows)
        return FlowAlignFunction.apply(features, level_flow)
