"""Drop-in for Generalized_RCNN.roi_feature_transform (lib/modeling/model_builder.py:252-324;
mirrors lib_vos/vos_modeling/vos_model_builder.py:449-521 and
generalized_rcnn_predictor_with_boxes.py:231-303), method='RoIAlign' only.

The reference runs one RoIAlign launch + one host->device roi copy per FPN level, then
torch.cat and an index_select with rois_idx_restore_int32.  Here all levels go through ONE
kernel that writes each RoI's features straight to its un-shuffled row.
"""
import numpy as np
import torch

from .roi_xfrom.roi_align.functions.roi_align import RoIAlignFunction, roi_align_multilevel
from ..config import get_cfg


def roi_feature_transform(blobs_in, rpn_ret, blob_rois='rois', method='RoIAlign',
                          resolution=7, spatial_scale=1. / 16., sampling_ratio=0, cfg=None):
    """Same arguments as the reference method (minus self).  ``blobs_in`` is a list of FPN maps
    in the reference's reversed order (coarsest first, model_builder.py:272-273) with
    ``spatial_scale`` a matching list, or a single tensor with a scalar scale.  ``rpn_ret`` holds
    ndarrays (reference style) or CUDA tensors."""
    if method != 'RoIAlign':
        raise NotImplementedError("only method='RoIAlign' is on the region pipeline "
                                  "(RoIPoolF / RoICrop are unused by the supported configs)")
    if not isinstance(blobs_in, list):
        rois = rpn_ret[blob_rois]
        if not isinstance(rois, torch.Tensor):
            rois = torch.from_numpy(np.ascontiguousarray(rois, dtype=np.float32))
        rois = rois.to(blobs_in.device)
        return RoIAlignFunction(resolution, resolution, spatial_scale, sampling_ratio)(blobs_in, rois)

    cfg = cfg or get_cfg()
    k_min, k_max = cfg.roi_min_level, cfg.roi_max_level
    assert len(blobs_in) == k_max - k_min + 1
    dev = blobs_in[0].device
    feats = [blobs_in[k_max - lvl] for lvl in range(k_min, k_max + 1)]          # finest first
    scales = [spatial_scale[k_max - lvl] for lvl in range(k_min, k_max + 1)]
    per_level = [rpn_ret['%s_fpn%d' % (blob_rois, lvl)] for lvl in range(k_min, k_max + 1)]
    restore = rpn_ret[blob_rois + '_idx_restore_int32']
    if isinstance(per_level[0], torch.Tensor):
        rois = torch.cat([r.to(dev) for r in per_level], dim=0)
        level = torch.cat([torch.full((len(r),), i, dtype=torch.int32, device=dev)
                           for i, r in enumerate(per_level)])
        restore = restore.to(dev).long()
        out_index = torch.empty_like(restore, dtype=torch.int32)
        out_index[restore] = torch.arange(restore.numel(), dtype=torch.int32, device=dev)
    else:
        rois_h = np.concatenate([np.asarray(r, dtype=np.float32).reshape(-1, 5) for r in per_level], axis=0)
        level_h = np.concatenate([np.full((len(r),), i, dtype=np.int32) for i, r in enumerate(per_level)])
        restore_h = np.asarray(restore).astype(np.int64)
        oi = np.empty(restore_h.shape[0], dtype=np.int32)
        oi[restore_h] = np.arange(restore_h.shape[0], dtype=np.int32)
        packed = np.concatenate([rois_h.view(np.int32).reshape(-1), level_h, oi])   # one H2D copy
        p = torch.from_numpy(packed).to(dev, non_blocking=True)
        R = rois_h.shape[0]
        rois = p[:5 * R].view(torch.float32).view(R, 5)
        level = p[5 * R:6 * R]
        out_index = p[6 * R:7 * R]
    return roi_align_multilevel(feats, scales, rois, level, resolution, resolution, sampling_ratio, out_index)
