"""Drop-in for the RPN label assignment of lib/roi_data/rpn.py: ``get_rpn_blob_names`` (:36-61),
``add_rpn_blobs(blobs, im_scales, roidb)`` (:64-140) and ``_get_rpn_blobs`` (:143-270).  Same blob names, shapes, dtypes
and quirks as the reference (see ``rpn_labels_cuda``).

Device work per image: the anchor x gt overlap matrix with its row max / argmax (``vosd_bbox_overlaps``), the foreground
subsample (``vosd_sample_rois`` on the label vector), the regression targets (``vosd_bbox_targets``); the label rules
between them are elementwise device ops with no host synchronisation.  One D2H per blob at the end.

RNG CONTRACT (the reference draws from NumPy's global generator):
  * ``npr.choice(fg_inds, size, replace=False)`` -- the foreground anchors to DISABLE -- is the `size` candidates with
    the smallest keys (ties: lower anchor index first), one uniform float32 key per anchor of the field
    (``rand_keys[i]``: (total_anchors,)), as in roi_data/fast_rcnn.py;
  * ``npr.randint(len(bg_inds), size=num_bg)`` -- positions in the ascending list of background anchors, drawn WITH
    replacement -- is ``floor(u[:num_bg] * len(bg_inds))`` (float64 product) of the first num_bg uniforms of
    ``rand_bg[i]``: (RPN_BATCH_SIZE_PER_IM,) float32 in [0, 1).
With ``rand_keys`` / ``rand_bg`` = None they are drawn from ``numpy.random.random_sample``: the reference's distribution,
not its sample.  tests/golden/make_golden_rpn_labels.py runs the unmodified reference under exactly this contract."""
import numpy as np
import torch

from .. import ops
from ..config import get_cfg
from . import data_utils

_NAMES = ('rpn_labels_int32_wide', 'rpn_bbox_targets_wide', 'rpn_bbox_inside_weights_wide',
          'rpn_bbox_outside_weights_wide')


def get_rpn_blob_names(is_training=True, cfg=None):
    c = cfg or get_cfg()
    names = ['im_info']
    if is_training:
        names += ['roidb']
        if c.fpn_on and c.multilevel_rpn:
            for lvl in range(c.rpn_min_level, c.rpn_max_level + 1):
                names += ['%s_fpn%d' % (n, lvl) for n in _NAMES]
        else:
            names += list(_NAMES)
    return names


def _fields(c):
    if c.fpn_on and c.multilevel_rpn:
        return [data_utils.get_field_of_anchors(2. ** lvl, (c.rpn_anchor_start_size * 2. ** (lvl - c.rpn_min_level),),
                                                c.rpn_aspect_ratios, cfg=c)
                for lvl in range(c.rpn_min_level, c.rpn_max_level + 1)]
    return [data_utils.get_field_of_anchors(c.rpn_stride, c.rpn_sizes, c.rpn_single_aspect_ratios, cfg=c)]


_anchor_cache = {}


def _device_anchors(foas, device):
    key = (tuple(id(f) for f in foas), str(device))
    if key not in _anchor_cache:
        _anchor_cache[key] = torch.from_numpy(np.concatenate([f.field_of_anchors for f in foas])).to(device)
    return _anchor_cache[key]


def rpn_labels_cuda(all_anchors, gt_boxes, im_height, im_width, keys, rand_bg, cfg=None):
    """_get_rpn_blobs (rpn.py:143-230) for one image, on the device, before the per-level split.
    all_anchors (T,4) fp32, gt_boxes (G,4) fp32 (scaled to the blob), keys (T) fp32, rand_bg (RPN_BATCH_SIZE_PER_IM) fp32
    -> labels (T) int32 in {-1, 0, 1}, bbox_targets / inside / outside weights (T,4) fp32.

    Kept as the reference has them: a gt box that overlaps no inside anchor marks EVERY zero-overlap inside anchor
    foreground (`anchor_by_gt_overlap == gt_to_anchor_max` with max 0); no background is labelled at all when there are
    not more than num_bg candidates; a background draw may land on a gt-forced foreground anchor and turn it into
    background, and that anchor keeps its regression target (fg_inds is taken before the draw) with zero weights."""
    c = cfg or get_cfg()
    a = all_anchors
    dev = a.device
    T = int(a.size(0))
    G = 0 if gt_boxes is None else int(gt_boxes.size(0))
    st = float(c.train_rpn_straddle_thresh)
    if st >= 0:
        inside = ((a[:, 0] >= -st) & (a[:, 1] >= -st) & (a[:, 2] < float(im_width) + st)
                  & (a[:, 3] < float(im_height) + st))
    else:
        inside = torch.ones(T, dtype=torch.bool, device=dev)
    R = int(c.train_rpn_batch_size_per_im)
    num_fg = int(c.train_rpn_fg_fraction * R)
    fg = torch.zeros(T, dtype=torch.bool, device=dev)
    if G > 0:
        ov, mx, am = ops.bbox_overlaps_cuda(a, gt_boxes, want_matrix=True)
        # gt -> best inside anchor (ties included)
        gt_max = torch.where(inside[:, None], ov, torch.full_like(ov, -1.0)).max(dim=0).values
        fg = inside & ((ov == gt_max[None, :]).any(dim=1) | (mx >= float(c.train_rpn_positive_overlap)))
        bg_c = inside & (mx < float(c.train_rpn_negative_overlap))
    else:
        mx = am = None
        bg_c = inside
    # foreground subsample: the reference disables the (n_fg - num_fg) smallest keys, i.e. keeps the num_fg LARGEST keys
    # with ties going to the HIGHER index: vosd_sample_rois on the flipped, negated arrays
    if num_fg > 0:
        score = fg.flip(0).to(torch.float32).view(1, T)
        nk = (-keys).flip(0).contiguous().view(1, T)
        nb = torch.full((1,), T, dtype=torch.int32, device=dev)
        keep, _, nkeep = ops.sample_rois_cuda(score, nk, nb, num_fg, num_fg, 0.5, -1.0, -1.0)
        valid = torch.arange(num_fg, device=dev) < nkeep[0]
        idx = torch.where(valid, T - 1 - keep[0].long(), torch.full((num_fg,), T, dtype=torch.long, device=dev))
        fg_keep = torch.zeros(T + 1, dtype=torch.bool, device=dev)
        fg_keep[idx] = True
        fg = fg_keep[:T]
    else:
        fg = torch.zeros_like(fg)
    n_fg = fg.sum()
    # background: positions floor(u * n_bg) in the ascending candidate list, only if there are more than num_bg of them
    num_bg = R - n_fg
    csum = torch.cumsum(bg_c.to(torch.int32), 0)
    n_bg = csum[-1]
    u = rand_bg[:R].to(torch.float64)
    pos = torch.floor(u * n_bg.to(torch.float64)).to(torch.int64).clamp_(max=torch.iinfo(torch.int32).max)
    pick = torch.searchsorted(csum, (pos + 1).to(torch.int32))
    use = (torch.arange(R, device=dev) < num_bg) & (n_bg > num_bg)
    bg = torch.zeros(T + 1, dtype=torch.bool, device=dev)
    bg[torch.where(use, pick.clamp_(max=T), torch.full_like(pick, T))] = True
    bg = bg[:T]
    labels = torch.full((T,), -1, dtype=torch.int32, device=dev)
    labels[fg] = 1
    labels[bg] = 0
    # targets of the anchors that were foreground BEFORE the background draw; weights of those still foreground after
    targets = torch.zeros((T, 4), dtype=torch.float32, device=dev)
    if G > 0:
        t, _, _ = ops.bbox_targets_cuda(a, gt_boxes[am.long()].contiguous(), fg.to(torch.int32), 2, (1.0, 1.0, 1.0, 1.0),
                                        class_agnostic=True)
        targets = t[:, 4:8].contiguous()
    pos1 = (labels == 1)[:, None]
    inside_w = pos1.to(torch.float32).expand(T, 4).contiguous()
    n_ex = (labels >= 0).sum().to(torch.float64)                  # 1.0 / num_examples is a float64 quotient, stored as fp32
    outside_w = torch.where((labels >= 0)[:, None], (1.0 / n_ex).to(torch.float32),
                            torch.zeros((), device=dev)).expand(T, 4).contiguous()
    return labels, targets, inside_w, outside_w


def _split(foas, labels, targets, inside_w, outside_w):
    out, s = [], 0
    for f in foas:
        H = W = f.field_size
        A = f.num_cell_anchors
        e = s + H * W * A
        out.append({
            'rpn_labels_int32_wide': labels[s:e].reshape(1, H, W, A).transpose(0, 3, 1, 2),
            'rpn_bbox_targets_wide': targets[s:e].reshape(1, H, W, A * 4).transpose(0, 3, 1, 2),
            'rpn_bbox_inside_weights_wide': inside_w[s:e].reshape(1, H, W, A * 4).transpose(0, 3, 1, 2),
            'rpn_bbox_outside_weights_wide': outside_w[s:e].reshape(1, H, W, A * 4).transpose(0, 3, 1, 2)})
        s = e
    return out


def _get_rpn_blobs(im_height, im_width, foas, all_anchors, gt_boxes, keys=None, rand_bg=None, cfg=None):
    """rpn.py:143-270: list (one dict per field of anchors; the dict itself for a single field) of ndarray blobs.
    all_anchors: (T,4) ndarray or device tensor; gt_boxes (G,4) ndarray."""
    c = cfg or get_cfg()
    dev = torch.device("cuda", torch.cuda.current_device())
    a = all_anchors if isinstance(all_anchors, torch.Tensor) else torch.from_numpy(
        np.ascontiguousarray(all_anchors, dtype=np.float32)).to(dev)
    T = int(a.size(0))
    R = int(c.train_rpn_batch_size_per_im)
    k = np.random.random_sample(T) if keys is None else keys
    u = np.random.random_sample(R) if rand_bg is None else rand_bg
    g = np.ascontiguousarray(gt_boxes, dtype=np.float32).reshape(-1, 4)
    lab, tg, iw, ow = rpn_labels_cuda(a, torch.from_numpy(g).to(dev) if len(g) else None, im_height, im_width,
                                      torch.from_numpy(np.asarray(k, np.float32)).to(dev),
                                      torch.from_numpy(np.asarray(u, np.float32)).to(dev), cfg=c)
    out = _split(foas, lab.cpu().numpy(), tg.cpu().numpy(), iw.cpu().numpy(), ow.cpu().numpy())
    return out[0] if len(out) == 1 else out


def add_rpn_blobs(blobs, im_scales, roidb, rand_keys=None, rand_bg=None, cfg=None):
    """blobs: dict of lists keyed by get_rpn_blob_names() (filled in place, then concatenated, like the reference);
    roidb entries with 'height', 'width', 'boxes', 'gt_classes', 'is_crowd'.  Returns True."""
    c = cfg or get_cfg()
    foas = _fields(c)
    fpn = c.fpn_on and c.multilevel_rpn
    dev = torch.device("cuda", torch.cuda.current_device())
    anchors = _device_anchors(foas, dev)
    for i, entry in enumerate(roidb):
        scale = im_scales[i]
        im_h = np.round(entry['height'] * scale)
        im_w = np.round(entry['width'] * scale)
        gt_inds = np.where((entry['gt_classes'] > 0) & (entry['is_crowd'] == 0))[0]
        gt_rois = entry['boxes'][gt_inds, :] * scale
        blobs['im_info'].append(np.array([[im_h, im_w, scale]], dtype=np.float32))
        rb = _get_rpn_blobs(im_h, im_w, foas, anchors, gt_rois, None if rand_keys is None else rand_keys[i],
                            None if rand_bg is None else rand_bg[i], cfg=c)
        if fpn:
            for j, lvl in enumerate(range(c.rpn_min_level, c.rpn_max_level + 1)):
                for k, v in rb[j].items():
                    blobs['%s_fpn%d' % (k, lvl)].append(v)
        else:
            for k, v in rb.items():
                blobs[k].append(v)
    for k, v in blobs.items():
        if isinstance(v, list) and len(v) > 0 and k != 'data_flow':
            blobs[k] = np.concatenate(v)
    valid = ['has_visible_keypoints', 'boxes', 'segms', 'seg_areas', 'gt_classes', 'gt_overlaps', 'is_crowd',
             'box_to_gt_ind_map', 'gt_keypoints']
    if c.identity_training:
        valid += ['gt_overlaps_id', 'instance_id', 'global_instance_id']
    blobs['roidb'] = [{k: e[k] for k in valid if k in e} for e in roidb]
    return True
