"""Drop-in for lib/modeling/generate_proposals.py:13-168.

``GenerateProposalsOp(anchors, spatial_scale).forward(rpn_cls_prob, rpn_bbox_pred, im_info)``
returns ``(rois ndarray (R,5) float32, roi_probs ndarray (R,1) float32)`` exactly like the
reference (:102) -- the mode (TRAIN/TEST settings) follows ``self.training`` (:106) -- but the
top-k, decode, clip, filter and NMS run on the GPU and there is ONE device->host copy at the
very end instead of three up front (:59-66).  ``forward_cuda`` is the tensor-in/tensor-out
variant that never leaves the device.
"""
import numpy as np
import torch
from torch import nn

from .. import ops
from ..config import get_cfg


class GenerateProposalsOp(nn.Module):
    def __init__(self, anchors, spatial_scale, cfg=None, check_nan=True):
        super().__init__()
        self._anchors = np.asarray(anchors, dtype=np.float64)
        self._num_anchors = self._anchors.shape[0]
        self._feat_stride = 1. / spatial_scale
        self._cfg = cfg
        self._check_nan = check_nan

    def _mode(self):
        return (self._cfg or get_cfg()).mode(self.training)

    def forward_cuda(self, rpn_cls_prob, rpn_bbox_pred, im_info):
        """Device tensors in -> (rois (N,cap,5), probs (N,cap), count (N,)) device tensors out."""
        m = self._mode()
        rois, probs, count = ops.generate_proposals_cuda(
            [(rpn_cls_prob, rpn_bbox_pred, self._anchors, self._feat_stride)], im_info,
            m.pre_nms_topN, m.post_nms_topN, m.nms_thresh, m.min_size)
        return rois[0], probs[0], count[0]

    def forward(self, rpn_cls_prob, rpn_bbox_pred, im_info):
        if not rpn_cls_prob.is_cuda:
            raise NotImplementedError("GenerateProposalsOp needs CUDA tensors (no CPU path)")
        dev = rpn_cls_prob.device
        scores = rpn_cls_prob.detach().float()
        deltas = rpn_bbox_pred.detach().float()
        flag = ops.any_nan_cuda(deltas) if self._check_nan else None
        info = torch.as_tensor(np.asarray(im_info.detach().cpu() if isinstance(im_info, torch.Tensor) else im_info,
                                          dtype=np.float32)).to(dev)
        rois, probs, count = self.forward_cuda(scores, deltas, info)
        count_h = count.cpu().numpy()                       # the one synchronising copy
        if flag is not None and int(flag.item()):
            raise ValueError('bbox_deltas nan')            # generate_proposals.py:62-63
        rois_h, probs_h = rois.cpu().numpy(), probs.cpu().numpy()
        out_r = [rois_h[i, :count_h[i]] for i in range(rois_h.shape[0])]
        out_p = [probs_h[i, :count_h[i], None] for i in range(rois_h.shape[0])]
        return (np.concatenate(out_r, axis=0).astype(np.float32, copy=False),
                np.concatenate(out_p, axis=0).astype(np.float32, copy=False))
