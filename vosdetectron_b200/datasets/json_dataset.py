"""Drop-in for the proposal merge of lib/datasets/json_dataset.py: ``add_proposals(roidb, rois, scales, crowd_thresh)``
(:413-427) -> ``_merge_proposal_boxes_into_roidb`` (:429-490) + ``_add_class_assignments`` (:513-532), the call the
training branch of CollectAndDistributeFpnRpnProposalsOp makes (collect_and_distribute_fpn_rpn_proposals.py:57-71;
lib_vos: imdb/vos/davis_db.py has the same function).

The proposal -> ground-truth overlaps, their row maxima / arg-maxima and the class lookup run on the device
(csrc/overlaps.cu: bit-identical to the compiled cython_bbox); the roidb entries are updated in place with the
reference's keys and dtypes.  ``gt_overlaps`` is kept dense (the reference re-wraps it into a scipy CSR matrix after
every merge; every reader on this path calls ``.toarray()`` first).  crowd_thresh > 0 (``_filter_crowd_proposals``) needs
pycocotools' RLE IoU and is not part of this path (the op passes 0)."""
import numpy as np
import torch

from .. import ops


def add_proposals(roidb, rois, scales, crowd_thresh=0):
    if crowd_thresh > 0:
        raise NotImplementedError("_filter_crowd_proposals (pycocotools iou) is outside the region pipeline")
    rois = np.ascontiguousarray(rois, dtype=np.float32)
    for i, entry in enumerate(roidb):
        inv = np.float32(1.) / np.float32(scales[i])
        boxes = np.ascontiguousarray(rois[rois[:, 0] == i, 1:] * inv, dtype=np.float32)
        n = boxes.shape[0]
        gt_inds = np.where(entry['gt_classes'] > 0)[0]
        K = entry['gt_overlaps'].shape[1]
        dense = entry['gt_overlaps'].toarray() if hasattr(entry['gt_overlaps'], 'toarray') else np.asarray(entry['gt_overlaps'])
        add = np.zeros((n, K), dtype=dense.dtype)
        b2g = -np.ones(n, dtype=entry['box_to_gt_ind_map'].dtype)
        if len(gt_inds) > 0 and n > 0:
            mx, cls, am = ops.label_proposals_cuda(
                torch.from_numpy(boxes).cuda(),
                torch.from_numpy(np.ascontiguousarray(entry['boxes'][gt_inds], dtype=np.float32)).cuda(),
                torch.from_numpy(np.ascontiguousarray(entry['gt_classes'][gt_inds], dtype=np.int32)).cuda())
            mx, cls, am = mx.cpu().numpy(), cls.cpu().numpy(), am.cpu().numpy()
            pos = np.where(mx > 0)[0]
            add[pos, cls[pos]] = mx[pos]
            b2g[pos] = gt_inds[am[pos]]
        entry['boxes'] = np.append(entry['boxes'], boxes.astype(entry['boxes'].dtype, copy=False), axis=0)
        entry['gt_classes'] = np.append(entry['gt_classes'], np.zeros(n, dtype=entry['gt_classes'].dtype))
        entry['seg_areas'] = np.append(entry['seg_areas'], np.zeros(n, dtype=entry['seg_areas'].dtype))
        entry['gt_overlaps'] = np.append(dense, add, axis=0)
        entry['is_crowd'] = np.append(entry['is_crowd'], np.zeros(n, dtype=entry['is_crowd'].dtype))
        entry['box_to_gt_ind_map'] = np.append(entry['box_to_gt_ind_map'], b2g)
    _add_class_assignments(roidb)


def _add_class_assignments(roidb):
    for entry in roidb:
        ov = entry['gt_overlaps'].toarray() if hasattr(entry['gt_overlaps'], 'toarray') else entry['gt_overlaps']
        entry['max_overlaps'] = ov.max(axis=1)
        entry['max_classes'] = ov.argmax(axis=1)
