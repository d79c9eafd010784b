"""Drop-in for lib/modeling/collect_and_distribute_fpn_rpn_proposals.py:18-138.

``collect(inputs, is_training)`` and ``distribute(rois, label_blobs)`` keep the reference's
ndarray-in / dict-of-ndarray-out signatures; the top-N, the level map and the per-level split
run in csrc/collect.cu.  The training branch of the op (:57-82) goes through the mirrors of
json_dataset.add_proposals and roi_data.fast_rcnn.add_fast_rcnn_blobs in this package (overlaps,
sampling, regression targets and the FPN split on the device; see roi_data/fast_rcnn.py for the
sampler's RNG contract).
"""
import numpy as np
import torch
from torch import nn

from .. import ops
from ..config import get_cfg


def _blob_names(prefix, k_min, k_max):
    # roi_data/fast_rcnn.py:36-105 with is_training=False and FPN.MULTILEVEL_ROIS
    return [prefix] + ['%s_fpn%d' % (prefix, l) for l in range(k_min, k_max + 1)] + [prefix + '_idx_restore_int32']


def _pack_levels(roi_inputs, score_inputs):
    """Per-level (R_l,5)/(R_l,1) ndarrays -> padded (L,1,cap,5)/(L,1,cap) device tensors + counts.
    The batch index stays in column 0, so one group spans the whole minibatch like the reference."""
    cap = max(1, max(len(r) for r in roi_inputs))
    L = len(roi_inputs)
    rois = np.zeros((L, 1, cap, 5), dtype=np.float32)
    probs = np.zeros((L, 1, cap), dtype=np.float32)
    count = np.zeros((L, 1), dtype=np.int32)
    for i, (r, s) in enumerate(zip(roi_inputs, score_inputs)):
        n = len(r)
        rois[i, 0, :n] = r
        probs[i, 0, :n] = np.asarray(s, dtype=np.float32).reshape(-1)
        count[i, 0] = n
    return torch.from_numpy(rois).cuda(), torch.from_numpy(probs).cuda(), torch.from_numpy(count).cuda()


def _to_blobs(rois_h, level_h, order_h, restore_h, k_min, k_max, prefix='rois'):
    blobs = {prefix: rois_h}
    for lvl in range(k_min, k_max + 1):
        blobs['%s_fpn%d' % (prefix, lvl)] = rois_h[order_h[level_h[order_h] == lvl]]
    blobs[prefix + '_idx_restore_int32'] = restore_h.astype(np.int32, copy=False)
    return blobs


def collect(inputs, is_training, cfg=None):
    """:91-106 -- list [rois_fpn2..6, probs_fpn2..6] of ndarrays -> (post,5) ndarray."""
    cfg = cfg or get_cfg()
    post = cfg.collect_post_topN(is_training)
    num_lvls = cfg.rpn_max_level - cfg.rpn_min_level + 1
    rois, probs, count = _pack_levels(inputs[:num_lvls], inputs[num_lvls:])
    out = ops.collect_distribute_cuda(rois, probs, count, post, 1, cfg.roi_min_level, cfg.roi_max_level,
                                      cfg.roi_canonical_scale, cfg.roi_canonical_level)
    n = int(out["count"].item())
    return out["rois"][0, :n].cpu().numpy()


def distribute(rois, label_blobs, cfg=None):
    """:109-138 -- (R,5) ndarray -> {'rois', 'rois_fpn2'..'rois_fpn5', 'rois_idx_restore_int32'}."""
    cfg = cfg or get_cfg()
    k_min, k_max = cfg.roi_min_level, cfg.roi_max_level
    rois = np.ascontiguousarray(rois, dtype=np.float32)
    if rois.shape[0] == 0:
        return _to_blobs(rois, np.zeros(0, np.int32), np.zeros(0, np.int64), np.zeros(0, np.int32), k_min, k_max)
    level, _, order, restore = ops.distribute_cuda(torch.from_numpy(rois).cuda(), k_min, k_max,
                                                   cfg.roi_canonical_scale, cfg.roi_canonical_level)
    return _to_blobs(rois, level.cpu().numpy(), order.cpu().numpy().astype(np.int64), restore.cpu().numpy(),
                     k_min, k_max)


def collect_and_distribute(inputs, is_training, cfg=None):
    """distribute(collect(...)) in a single launch / single round trip."""
    cfg = cfg or get_cfg()
    post = cfg.collect_post_topN(is_training)
    num_lvls = cfg.rpn_max_level - cfg.rpn_min_level + 1
    rois, probs, count = _pack_levels(inputs[:num_lvls], inputs[num_lvls:])
    k_min, k_max = cfg.roi_min_level, cfg.roi_max_level
    out = ops.collect_distribute_cuda(rois, probs, count, post, 1, k_min, k_max,
                                      cfg.roi_canonical_scale, cfg.roi_canonical_level)
    n = int(out["count"].item())
    return _to_blobs(out["rois"][0, :n].cpu().numpy(), out["level"][0, :n].cpu().numpy(),
                     out["order"][0, :n].cpu().numpy().astype(np.int64), out["restore"][0, :n].cpu().numpy(),
                     k_min, k_max)


class CollectAndDistributeFpnRpnProposalsOp(nn.Module):
    def __init__(self, cfg=None):
        super().__init__()
        self._cfg = cfg

    def forward(self, inputs, roidb, im_info, rand_keys=None):
        """Inference: distribute(collect(inputs)).  Training (:57-82): add_proposals on the roidb entries, then
        add_fast_rcnn_blobs (sampling, regression targets, FPN split); ``rand_keys`` is the sampler's RNG contract
        (roi_data/fast_rcnn.py), None = keys drawn from numpy's global generator."""
        cfg = self._cfg or get_cfg()
        if not self.training:
            return collect_and_distribute(inputs, False, self._cfg)
        from ..datasets import json_dataset
        from ..roi_data import fast_rcnn
        rois = collect(inputs, True, self._cfg)
        info = im_info.detach().cpu().numpy() if isinstance(im_info, torch.Tensor) else np.asarray(im_info)
        im_scales = info[:, 2]
        json_dataset.add_proposals(roidb, rois, im_scales, crowd_thresh=0)
        blobs = {k: [] for k in fast_rcnn.get_fast_rcnn_blob_names(True, cfg)}
        fast_rcnn.add_fast_rcnn_blobs(blobs, im_scales, roidb, rand_keys, cfg)
        return blobs
