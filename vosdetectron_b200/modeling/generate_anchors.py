"""Anchor enumeration (host, model-build time) -- mirrors the public surface of
lib/modeling/generate_anchors.py:54-123: ``generate_anchors(stride, sizes, aspect_ratios)``
returns an (A,4) float64 array.  Values are integers / half-integers, so the fp64 -> fp32
rounding the reference does later (lib/utils/boxes.py:164) is exact; the CUDA decode kernel
receives these fp64 values and shifts them in fp64 like generate_proposals.py:69-89."""
import numpy as np


def _window_anchor(w, h, ctr):
    return [ctr - 0.5 * (w - 1), ctr - 0.5 * (h - 1), ctr + 0.5 * (w - 1), ctr + 0.5 * (h - 1)]


def generate_anchors(stride=16, sizes=(32, 64, 128, 256, 512), aspect_ratios=(0.5, 1, 2)):
    stride = float(stride)
    scales = np.asarray(sizes, dtype=np.float64) / stride
    ctr = 0.5 * (stride - 1.0)                   # centre of the (0, 0, stride-1, stride-1) window
    area = stride * stride
    out = []
    for ratio in np.asarray(aspect_ratios, dtype=np.float64):
        w = np.round(np.sqrt(area / ratio))      # ratio enumeration: rounded widths / heights
        h = np.round(w * ratio)
        out.extend(_window_anchor(w * s, h * s, ctr) for s in scales)   # scale enumeration
    return np.asarray(out, dtype=np.float64)


def fpn_level_anchors(level, start_size=32, aspect_ratios=(0.5, 1, 2), k_min=2):
    """One anchor size per FPN level, as fpn_rpn_outputs builds them (lib/modeling/FPN.py:343-350)."""
    return generate_anchors(2.0 ** level, (start_size * 2.0 ** (level - k_min),), aspect_ratios)
