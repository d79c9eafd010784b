"""Field of anchors of the RPN label assignment: lib/roi_data/data_utils.py:39-102 (``FieldOfAnchors``,
``get_field_of_anchors``).  Host code, run once per (stride, sizes, ratios) and cached like the reference's."""
import collections

import numpy as np

from ..config import get_cfg
from ..modeling.generate_anchors import generate_anchors

FieldOfAnchors = collections.namedtuple(
    'FieldOfAnchors', ['field_of_anchors', 'num_cell_anchors', 'stride', 'field_size', 'octave', 'aspect'])

_cache = {}


def get_field_of_anchors(stride, anchor_sizes, anchor_aspect_ratios, octave=None, aspect=None, cfg=None):
    """Every cell anchor at every position of the largest training blob: (field_size**2 * A, 4) float32 in
    (y, x, anchor) order, field_size = ceil(COARSEST_STRIDE * ceil(TRAIN.MAX_SIZE / COARSEST_STRIDE) / stride)."""
    c = cfg or get_cfg()
    key = (float(stride), tuple(anchor_sizes), tuple(anchor_aspect_ratios), c.fpn_coarsest_stride, c.train_max_size)
    if key in _cache:
        return _cache[key]
    cell = generate_anchors(stride=stride, sizes=anchor_sizes, aspect_ratios=anchor_aspect_ratios)
    A = cell.shape[0]
    fpn_max = c.fpn_coarsest_stride * np.ceil(c.train_max_size / float(c.fpn_coarsest_stride))
    field = int(np.ceil(fpn_max / float(stride)))
    sh = np.arange(0, field) * stride
    sx, sy = np.meshgrid(sh, sh)
    shifts = np.stack([sx.ravel(), sy.ravel(), sx.ravel(), sy.ravel()], axis=1)
    foa = FieldOfAnchors(
        field_of_anchors=(shifts[:, None, :] + cell[None, :, :]).reshape(field * field * A, 4).astype(np.float32),
        num_cell_anchors=A, stride=stride, field_size=field, octave=octave, aspect=aspect)
    _cache[key] = foa
    return foa
