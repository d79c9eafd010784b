"""Drop-in for lib_vos/vos_model/flow_align/modules/flow_align.py:5-38 (``FlowAlign(spatial_scale)``).

The full-resolution flow is brought to the feature map's resolution by the reference's fixed,
non-trainable strided convolution (kernel = stride = 1/spatial_scale, weights spatial_scale**3 on the
diagonal: a box mean that also rescales the displacement, :12-25), then the map is warped.
"""
import torch
import torch.nn as nn

from ..functions.flow_align import FlowAlignFunction


class FlowAlign(nn.Module):
    def __init__(self, spatial_scale):
        super(FlowAlign, self).__init__()
        self.spatial_scale = spatial_scale
        self.feature_size = None
        self.conv_flow_downsample = self._flow_downsample_convolutional_layer(spatial_scale)

    def _flow_downsample_convolutional_layer(self, spatial_scale):
        assert spatial_scale <= 1.0 and spatial_scale in [1.0, 0.5, 0.25, 0.125, 0.0625, 0.03125, 1. / 64.]
        inv_scale = int(1.0 / spatial_scale)
        conv = nn.Conv2d(2, 2, kernel_size=inv_scale, stride=inv_scale, padding=0, dilation=1, groups=1, bias=False)
        weights = torch.zeros(conv.weight.shape)
        for idx in range(weights.shape[0]):
            weights[idx, idx, :, :] = spatial_scale ** 3
        conv.weight = torch.nn.Parameter(weights, requires_grad=False)   # "This layer should not be trained."
        return conv

    def forward(self, features, flows):
        _flows = self.conv_flow_downsample(flows) if self.spatial_scale != 1.0 else flows
        return FlowAlignFunction.apply(features, _flows)
