"""Drop-in for the Fast R-CNN label assignment of lib/roi_data/fast_rcnn.py: ``get_fast_rcnn_blob_names`` (:36-105),
``add_fast_rcnn_blobs(blobs, im_scales, roidb)`` (:108-129), ``_sample_rois`` (:132-213), ``_add_multilevel_rois``
(:262-290), plus the ``mask_rois`` / ``roi_has_mask_int32`` part of ``roi_data.mask_rcnn.add_mask_rcnn_blobs``
(mask_rcnn.py:34-102) when MODEL.MASK_ON.  Same blob names, dtypes and row order as the reference.

Device work: the fg / bg index selection of every image in one launch (``vosd_sample_rois``), the regression targets
(``vosd_bbox_targets``) and the FPN level split (``vosd_distribute``).  The blobs come back as ndarrays (one D2H per blob
at the very end), like every other ndarray-returning shim of this package.

RNG CONTRACT.  The reference samples with ``npr.choice(inds, size, replace=False)`` from NumPy's global generator.  Here
the sampler consumes ONE UNIFORM KEY PER BOX (``rand_keys``: list of (num_boxes_i,) float32 arrays, one per image) and
"choice" = the `size` candidates with the smallest keys in ascending key order (ties: lower index first).  With
``rand_keys=None`` the keys are drawn from ``numpy.random.random_sample`` (the generator the reference uses): the same
distribution as the reference's draw, not the same sample.  tests/golden/make_golden_labels.py runs the unmodified
reference under exactly this contract.

Not built here: ``masks_int32`` (polygon rasterisation by pycocotools, absent from this image), keypoint blobs."""
import numpy as np
import torch

from .. import ops
from ..config import get_cfg


def get_fast_rcnn_blob_names(is_training=True, cfg=None):
    c = cfg or get_cfg()
    names = ['rois']
    if is_training:
        names += ['labels_int32', 'bbox_targets', 'bbox_inside_weights', 'bbox_outside_weights']
        if c.mask_on:
            names += ['mask_rois', 'roi_has_mask_int32', 'masks_int32']
    lv = range(c.roi_min_level, c.roi_max_level + 1)
    names += ['rois_fpn%d' % l for l in lv] + ['rois_idx_restore_int32']
    if is_training and c.mask_on:
        names += ['mask_rois_fpn%d' % l for l in lv] + ['mask_rois_idx_restore_int32']
    return names


def _distribute(blobs, name, c):
    rois = blobs[name]
    if rois.shape[0] == 0:
        for l in range(c.roi_min_level, c.roi_max_level + 1):
            blobs['%s_fpn%d' % (name, l)] = np.zeros((0, 5), np.float32)
        blobs[name + '_idx_restore_int32'] = np.zeros(0, np.int32)
        return
    level, _, order, restore = ops.distribute_cuda(torch.from_numpy(np.ascontiguousarray(rois)).cuda(), c.roi_min_level,
                                                   c.roi_max_level, c.roi_canonical_scale, c.roi_canonical_level)
    level, order = level.cpu().numpy(), order.cpu().numpy().astype(np.int64)
    for l in range(c.roi_min_level, c.roi_max_level + 1):
        blobs['%s_fpn%d' % (name, l)] = rois[order[level[order] == l]]
    blobs[name + '_idx_restore_int32'] = restore.cpu().numpy().astype(np.int32, copy=False)


def add_fast_rcnn_blobs(blobs, im_scales, roidb, rand_keys=None, cfg=None):
    """blobs: dict of lists keyed by get_fast_rcnn_blob_names() (filled in place, like the reference); roidb entries as
    json_dataset.add_proposals leaves them (boxes, max_overlaps, max_classes, gt_classes, box_to_gt_ind_map).
    Returns True (the reference's `valid`)."""
    c = cfg or get_cfg()
    B = len(roidb)
    R = int(c.train_batch_size_per_im)
    fg_per = int(np.round(c.train_fg_fraction * R))
    nmax = max(1, max(e['boxes'].shape[0] for e in roidb))
    ov_h = np.full((B, nmax), -1.0, np.float32)
    key_h = np.ones((B, nmax), np.float32)
    for i, e in enumerate(roidb):
        n = e['boxes'].shape[0]
        ov_h[i, :n] = e['max_overlaps']
        key_h[i, :n] = np.random.random_sample(n) if rand_keys is None else np.asarray(rand_keys[i], np.float32)
    nb = torch.tensor([e['boxes'].shape[0] for e in roidb], dtype=torch.int32).cuda()
    keep, nfg, nkeep = ops.sample_rois_cuda(torch.from_numpy(ov_h).cuda(), torch.from_numpy(key_h).cuda(), nb, R, fg_per,
                                            c.train_fg_thresh, c.train_bg_thresh_hi, c.train_bg_thresh_lo)
    keep_h, nfg_h, nkeep_h = keep.cpu().numpy(), nfg.cpu().numpy(), nkeep.cpu().numpy()
    out = {k: [] for k in ('labels_int32', 'rois', 'bbox_targets', 'bbox_inside_weights', 'bbox_outside_weights',
                           'mask_rois', 'roi_has_mask_int32')}
    for i, e in enumerate(roidb):
        k = keep_h[i, :nkeep_h[i]].astype(np.int64)
        labels = e['max_classes'][k].astype(np.int32)
        labels[nfg_h[i]:] = 0                                           # fast_rcnn.py:166
        boxes = np.ascontiguousarray(e['boxes'][k], dtype=np.float32)
        gt_inds = np.where(e['gt_classes'] > 0)[0]
        gt_assign = gt_inds[e['box_to_gt_ind_map'][k]] if len(gt_inds) else np.zeros(len(k), np.int64)
        gt_rows = np.ascontiguousarray(e['boxes'][gt_assign], dtype=np.float32) if len(k) else np.zeros((0, 4), np.float32)
        if len(k):
            t, iw, ow = ops.bbox_targets_cuda(torch.from_numpy(boxes).cuda(), torch.from_numpy(gt_rows).cuda(),
                                              torch.from_numpy(labels).cuda(), c.num_classes, c.bbox_reg_weights,
                                              c.cls_agnostic_bbox_reg)
            t, iw, ow = t.cpu().numpy(), iw.cpu().numpy(), ow.cpu().numpy()
        else:
            Kc = 2 if c.cls_agnostic_bbox_reg else c.num_classes
            t = iw = ow = np.zeros((0, 4 * Kc), np.float32)
        s = np.float32(im_scales[i])
        rois = np.hstack((i * np.ones((len(k), 1), np.float32), boxes * s)).astype(np.float32)
        out['labels_int32'].append(labels)
        out['rois'].append(rois)
        out['bbox_targets'].append(t)
        out['bbox_inside_weights'].append(iw)
        out['bbox_outside_weights'].append(ow)
        if c.mask_on:
            fg = np.where(labels > 0)[0]
            has = (labels > 0).astype(np.int32)
            if fg.size > 0:
                rf = boxes[fg].copy()
            else:                                                        # mask_rcnn.py:72-86: first bg roi, ignore mask
                rf = boxes[np.where(labels == 0)[0][0]].reshape((1, -1)).copy()
                has[0] = 1
            out['mask_rois'].append(np.hstack((i * np.ones((rf.shape[0], 1), np.float32), rf * s)).astype(np.float32))
            out['roi_has_mask_int32'].append(has)
    for name, parts in out.items():
        if parts and name in blobs:
            blobs[name] = np.concatenate(parts)
    _distribute(blobs, 'rois', c)
    if c.mask_on:
        _distribute(blobs, 'mask_rois', c)
        if 'masks_int32' in blobs and isinstance(blobs['masks_int32'], list):
            blobs.pop('masks_int32')        # polygon rasterisation (pycocotools) is not part of this package
    return True
