"""Drop-ins for the detection post-processing of lib_vos/tools/vos_test.py (imported there as ``vos_test``):

``box_results_with_nms_and_limit(scores, boxes, prev_cls_boxes=None)``  vos_test.py:748-865 -- the lib/ version
    (threshold, per-class NMS, over-all-classes limit: four launches, see core/test.py) plus the VOS extras:
    NMS across classes (:816-837, one more vosd_nms launch), the per-class top-k (:839-843) and the filter against
    the previous frame's box (:845-860; at most one box per class, host arithmetic on <= 100 rows exactly as written
    there).
``nms_with_mask_iou(cls_boxes, cls_segms, iou_th, max_per_class)``  vos_test.py:985-1029 -- the O(R^2) loop over
    decoded full-frame masks becomes AND + POPC over bit-packed masks on the device (vosd_rle_to_bits +
    vosd_mask_iou_nms); only the RLE string parsing stays on the host.
``nms_with_mask_iou_cuda(scores, packed, iou_th)`` is the tensor-in / tensor-out variant for masks that never left
    the device (the bit-packed output of vosd_paste_masks_packed).
``segm_results`` is the same function as in lib/core/test.py (vos_test.py:867-921 is an identical copy).
"""
import numpy as np
import torch

from .. import ops
from ..config import get_cfg
from ..utils import boxes as box_utils
from . import test as core_test
from .test import segm_results  # noqa: F401  (re-export, vos_test.py:867)


def bb_intersection_over_union(boxA, boxB):
    """vos_test.py:961-982, kept verbatim in its arithmetic (NumPy scalar types decide the rounding)."""
    xA = max(boxA[0], boxB[0])
    yA = max(boxA[1], boxB[1])
    xB = min(boxA[2], boxB[2])
    yB = min(boxA[3], boxB[3])
    interArea = max(0, xB - xA + 1) * max(0, yB - yA + 1)
    boxAArea = (boxA[2] - boxA[0] + 1) * (boxA[3] - boxA[1] + 1)
    boxBArea = (boxB[2] - boxB[0] + 1) * (boxB[3] - boxB[1] + 1)
    return interArea / float(boxAArea + boxBArea - interArea)


def box_results_with_nms_and_limit(scores, boxes, prev_cls_boxes=None, cfg=None):
    cfg = cfg or get_cfg()
    K = cfg.num_classes
    _, _, cls_boxes = core_test.box_results_with_nms_and_limit(scores, boxes, cfg, _per_class_limit=False)
    if cfg.test_nms_cross_class > 0.:                                        # :816-837
        all_dets = np.vstack([cls_boxes[j] for j in range(1, K)])
        class_ids = np.vstack([np.ones(shape=(len(cls_boxes[j]), 1)) * j for j in range(1, K)])
        keep = box_utils.nms(all_dets, cfg.test_nms_cross_class)
        all_dets = all_dets[keep, :]
        class_ids = class_ids[keep, :]
        for j in range(1, K):
            idx_j = np.where(class_ids == j)[0]
            cls_boxes[j] = all_dets[idx_j, :]
    if cfg.test_num_det_per_class_pre > 0:                                   # :839-843
        for j in range(1, K):
            keep = np.argsort(-cls_boxes[j][:, -1])[:cfg.test_num_det_per_class_pre]
            cls_boxes[j] = cls_boxes[j][keep, :]
    if cfg.test_nms_small_box_iou > 0:                                       # :845-860
        for j in range(1, K):
            if prev_cls_boxes is not None:
                assert len(prev_cls_boxes[j]) < 2, 'number of prev boxes should <2.'
                if len(prev_cls_boxes[j]) == 1:
                    if prev_cls_boxes[j][0][-1] < cfg.test_nms_small_box_score_threshold:
                        continue
                    prev_cls_box = prev_cls_boxes[j][0][:-1]
                    index_to_remove = []
                    for id_box in range(len(cls_boxes[j]) - 1, -1, -1):
                        iou = bb_intersection_over_union(prev_cls_box, cls_boxes[j][id_box][:-1])
                        if iou < cfg.test_nms_small_box_iou:
                            index_to_remove.append(id_box)
                    cls_boxes[j] = np.delete(cls_boxes[j], index_to_remove, 0)
    im_results = np.vstack([cls_boxes[j] for j in range(1, K)])
    return im_results[:, -1], im_results[:, :-1], cls_boxes


def convert_from_cls_format(cls_boxes, cls_segms, cls_keyps):
    """vos_test.py:922-945."""
    box_list = [b for b in cls_boxes if len(b) > 0]
    boxes = np.concatenate(box_list) if len(box_list) > 0 else None
    segms = [s for slist in cls_segms for s in slist] if cls_segms is not None else None
    keyps = [k for klist in cls_keyps for k in klist] if cls_keyps is not None else None
    classes = []
    for j in range(len(cls_boxes)):
        classes += [j] * len(cls_boxes[j])
    return boxes, segms, keyps, classes


def nms_with_mask_iou_cuda(scores, packed, iou_th):
    """scores (R) and bit-packed masks (R, bytes) on the device -> (order (R) int32 = detection index per
    descending-score position, removed (R) int32 per position, num_keep (1) int32), all on the device."""
    order = torch.argsort(scores, descending=True, stable=True).to(torch.int32)
    removed, num = ops.mask_iou_nms_cuda(packed, order, iou_th)
    return order, removed, num


def nms_with_mask_iou(cls_boxes, cls_segms, iou_th=0.9, max_per_class=1, cfg=None):
    cfg = cfg or get_cfg()
    if not isinstance(cls_boxes, list):
        raise TypeError("cls_boxes must be the per-class list format")        # the reference leaves `boxes` unbound
    boxes, segms, _, classes = convert_from_cls_format(cls_boxes, cls_segms, None)
    if boxes is None:
        return cls_boxes, cls_segms
    classes = np.array(classes)
    sorted_inds = np.argsort(-boxes[:, -1])                                   # :995
    h, w = (int(v) for v in segms[0]['size'])
    runs = [core_test.rle_counts_from_string(s['counts']) for s in segms]     # the parse half of mask_util.decode
    packed = ops.rle_to_bits_cuda(runs, h * w)
    order = torch.from_numpy(sorted_inds.astype(np.int32)).to(packed.device)
    removed, _ = ops.mask_iou_nms_cuda(packed, order, float(iou_th))
    keep = np.flatnonzero(removed.cpu().numpy() == 0)                         # positions in the sorted order (:1012)
    boxes = boxes[sorted_inds][keep, :]
    classes = classes[sorted_inds][keep]
    new_segms = []
    for k in keep:                                                            # encode(decode(rle)) is the same RLE (:1017-1020)
        s = segms[int(sorted_inds[k])]
        c = s['counts']
        new_segms.append({'size': [h, w], 'counts': c.decode('ascii') if isinstance(c, bytes) else c})
    num_classes = cfg.num_classes
    out_boxes = [[] for _ in range(num_classes)]
    out_segms = [[] for _ in range(num_classes)]
    for i in range(len(keep)):                                                # :1022-1028
        clss = int(classes[i])
        if len(out_boxes[clss]) < max_per_class:
            out_boxes[clss].append(boxes[i, :])
            out_segms[clss].append(new_segms[i])
    return out_boxes, out_segms
