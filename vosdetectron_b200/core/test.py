"""Drop-ins for segm_results (lib/core/test.py:801-855; identical copy in
lib_vos/tools/vos_test.py:867-921) and box_results_with_nms_and_limit (lib/core/test.py:733-797).

``segm_results(cls_boxes, masks, ref_boxes, im_h, im_w)`` keeps the reference signature and
returns ``cls_segms``: per class a list of COCO RLE dicts.  The expand / resize / threshold /
paste / RLE work is ONE kernel over all detections (csrc/paste.cu: paste_rle_kernel); ``rle_encode``
below is the host restatement kept for callers that already hold a dense mask.
``paste_masks`` returns the dense (R, im_h, im_w) uint8 masks without the RLE step.
"""
import numpy as np
import torch

from .. import ops
from ..config import get_cfg


def box_results_with_nms_and_limit(scores, boxes, cfg=None, _per_class_limit=True):
    """Drop-in for lib/core/test.py:733-797: ``scores`` (R,K), ``boxes`` (R,4K) ndarrays (or CUDA tensors) ->
    ``(scores, boxes, cls_boxes)`` ndarrays, cls_boxes[j] = (n_j,5) [x1,y1,x2,y2,score].  One upload, four
    launches (threshold + sort per class, IoU bitmask, greedy reduce, over-all-classes limit), one download."""
    cfg = cfg or get_cfg()
    if cfg.test_soft_nms or cfg.test_bbox_vote:
        raise NotImplementedError("TEST.SOFT_NMS / TEST.BBOX_VOTE (off by default) are not implemented on the device")
    K = cfg.num_classes
    s = scores if isinstance(scores, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(scores, dtype=np.float32))
    b = boxes if isinstance(boxes, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(boxes, dtype=np.float32))
    R = int(s.shape[0])
    assert s.shape[1] == K and b.shape[1] == 4 * K
    cls_boxes = [[] for _ in range(K)]
    if R == 0:
        for j in range(1, K):
            cls_boxes[j] = np.zeros((0, 5), dtype=np.float32)
        return np.zeros((0,), np.float32), np.zeros((0, 4), np.float32), cls_boxes
    dets, count, cls_count = ops.box_results_cuda(s.cuda()[None], b.cuda()[None], cfg.test_score_thresh, cfg.test_nms,
                                                  cfg.test_detections_per_im, cap=R * (K - 1))
    n = int(count[0].item())
    d = dets[0, :n].cpu().numpy()
    cc = cls_count[0].cpu().numpy()
    start = 0
    for j in range(1, K):
        cls_boxes[j] = d[start:start + cc[j], :5].copy()
        start += int(cc[j])
    if _per_class_limit and cfg.test_num_det_per_class > 0:  # test.py:785-788 (the lib_vos copy has no such branch)
        for j in range(1, K):
            keep = np.argsort(-cls_boxes[j][:, -1])[:cfg.test_num_det_per_class]
            cls_boxes[j] = cls_boxes[j][keep, :]
    im_results = np.vstack([cls_boxes[j] for j in range(1, K)])
    return im_results[:, -1], im_results[:, :-1], cls_boxes


def rle_encode(mask):
    """COCO compressed RLE of a (H,W) uint8 mask (column-major runs, LEB128-like string) --
    what pycocotools.mask.encode returns for one mask, with 'counts' as ascii str."""
    h, w = mask.shape
    flat = np.asarray(mask, dtype=np.uint8).ravel(order='F')
    change = np.flatnonzero(flat[1:] != flat[:-1]) + 1
    bounds = np.concatenate(([0], change, [flat.size]))
    counts = np.diff(bounds).tolist()
    if flat.size and flat[0] == 1:
        counts = [0] + counts
    out = []
    for i, x in enumerate(counts):
        if i > 2:
            x -= counts[i - 2]
        more = True
        while more:
            c = x & 0x1f
            x >>= 5
            more = (x != -1) if (c & 0x10) else (x != 0)
            if more:
                c |= 0x20
            out.append(chr(c + 48))
    return {'size': [int(h), int(w)], 'counts': ''.join(out)}


def rle_counts_from_string(s):
    """'counts' string of a COCO RLE -> list of run lengths (pycocotools common/maskApi.c rleFrString)."""
    if isinstance(s, bytes):
        s = s.decode('ascii')
    counts, p, n = [], 0, len(s)
    while p < n:
        x, k, more = 0, 0, True
        while more:
            c = ord(s[p]) - 48
            x |= (c & 0x1f) << (5 * k)
            more = bool(c & 0x20)
            p += 1
            k += 1
            if not more and (c & 0x10):
                x |= -1 << (5 * k)
        if len(counts) > 2:
            x += counts[-2]
        counts.append(x)
    return counts


def _class_order(cls_boxes, num_classes):
    """Mask channel of every detection in the reference's class-major running order (:816-850)."""
    cls = []
    for j in range(1, num_classes):
        cls += [j] * int(len(cls_boxes[j]))
    return np.asarray(cls, dtype=np.int32)


def _device_inputs(cls_boxes, masks, ref_boxes, cfg):
    cls = _class_order(cls_boxes, cfg.num_classes)
    R = cls.shape[0]
    assert R == masks.shape[0]                                    # test.py:854
    if R == 0:
        return cls, None, None, None
    if isinstance(masks, torch.Tensor) and masks.is_cuda:
        m = masks
        c = torch.from_numpy(cls if cfg.mrcnn_cls_specific_mask else np.zeros_like(cls)).to(m.device)
    else:
        masks = np.asarray(masks, dtype=np.float32)
        sel = masks[np.arange(R), cls if cfg.mrcnn_cls_specific_mask else 0]      # (R,M,M): upload only what is read
        m = torch.from_numpy(np.ascontiguousarray(sel[:, None])).cuda()
        c = None
    b = torch.from_numpy(np.ascontiguousarray(ref_boxes, dtype=np.float32)).to(m.device)
    return cls, m, c, b


def paste_masks(cls_boxes, masks, ref_boxes, im_h, im_w, cfg=None):
    """Dense (R, im_h, im_w) uint8 masks of segm_results (the `im_mask` canvases, test.py:825-841) + mask channels."""
    cfg = cfg or get_cfg()
    cls, m, c, b = _device_inputs(cls_boxes, masks, ref_boxes, cfg)
    if m is None:
        return np.zeros((0, im_h, im_w), dtype=np.uint8), cls
    out = ops.paste_masks_cuda(m, c, b, int(im_h), int(im_w), cfg.mrcnn_thresh_binarize)
    return out.cpu().numpy(), cls


def segm_results(cls_boxes, masks, ref_boxes, im_h, im_w, cfg=None):
    """lib/core/test.py:801-855 end to end on the device: expand / resize / threshold / paste / RLE encode are ONE
    kernel (vosd_paste_rle); only the 'counts' strings (a few hundred bytes per detection) come back, the dense
    (im_h, im_w) canvases of the reference are never materialised."""
    cfg = cfg or get_cfg()
    cls, m, c, b = _device_inputs(cls_boxes, masks, ref_boxes, cfg)
    cls_segms = [[] for _ in range(cfg.num_classes)]
    if m is None:
        return cls_segms
    rles = ops.rle_results(m, c, b, int(im_h), int(im_w), cfg.mrcnn_thresh_binarize)
    for i, j in enumerate(cls):
        cls_segms[int(j)].append(rles[i])
    return cls_segms
