"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the per-frame region pipeline.

NumPy restatement of the reference's Python/NumPy hot path plus ctypes bindings
to ``oracle/oracle.c`` (NMS, RoIAlign fwd/bwd, cv2-style resize).  Only tests/,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this module; ``vosdetectron_b200`` never does.

Parity status: the reference has no tests / golden vectors for this path
("parity unpinned by upstream tests"), so the oracle is
pinned against *outputs of the reference itself run in the build container*:
``tests/golden/*.npz`` (made by ``tests/golden/make_golden.py``) and the live
comparison in ``tests/test_oracle_vs_reference.py``.  The one embedded
known-answer of the reference (the 9 stride-16 anchors quoted in
lib/modeling/generate_anchors.py:26-51) is checked in tests/test_oracle.py.

Every function cites the reference file:line it follows.  dtypes follow the
reference exactly (float64 anchors, float32 everything downstream, int64 index
arrays) because the rounding points are part of the contract.
"""
import ctypes
import math
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
# lib/core/config.py:1009 -- an np.float64 *scalar*, not a Python float.  That matters:
# under NumPy >= 2 (NEP 50) ``np.minimum(float32_array, np.float64_scalar)`` is float64, so
# the reference as run in this image carries dw/dh, exp() and pred_w/pred_h in float64 and
# rounds to float32 only when storing into pred_boxes (boxes.py:195-203); under NumPy 1.x
# they stay float32.  The oracle reproduces whichever NumPy it runs under, like the reference.
BBOX_XFORM_CLIP = np.log(1000. / 16.)


def build_c(force=False):
    """gcc-compile oracle.c -> oracle/liboracle.so (git-ignored, travels with gpurun)."""
    so = os.path.join(HERE, "liboracle.so")
    src = os.path.join(HERE, "oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-fopenmp", "-mfma",
                               "-ffp-contract=off", src, "-o", so, "-lm"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = ctypes.CDLL(build_c())
        fp = ctypes.POINTER(ctypes.c_float)
        L.orc_nms.restype = ctypes.c_int64
        L.orc_nms.argtypes = [fp, ctypes.c_int64, ctypes.c_float, ctypes.POINTER(ctypes.c_int64)]
        sig = [fp] + [ctypes.c_int] * 4 + [fp] + [ctypes.c_int] * 3 + [ctypes.c_float, ctypes.c_int, fp, ctypes.c_int]
        L.orc_roialign_fwd.argtypes = sig
        L.orc_roialign_bwd.argtypes = sig
        L.orc_resize_linear.argtypes = [fp, ctypes.c_int, ctypes.c_int, fp, ctypes.c_int, ctypes.c_int]
        L.orc_flow_align_fwd.argtypes = [fp, fp] + [ctypes.c_int] * 4 + [fp, ctypes.c_int]
        L.orc_flow_align_bwd.argtypes = [fp, fp, fp] + [ctypes.c_int] * 4 + [fp, fp, ctypes.c_int]
        L.orc_num_threads.restype = ctypes.c_int
        _LIB = L
    return _LIB


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


# --------------------------------------------------------------------------- #
# (a1) anchors -- lib/modeling/generate_anchors.py:54-123
# --------------------------------------------------------------------------- #
def generate_anchors(stride=16, sizes=(32, 64, 128, 256, 512), aspect_ratios=(0.5, 1, 2)):
    """(A,4) float64 anchors: ratio enumeration (rounded w/h) about the centre of
    the [0, stride-1] window, then scale enumeration (generate_anchors.py:66-123)."""
    scales = np.asarray(sizes, dtype=np.float64) / stride
    ratios = np.asarray(aspect_ratios, dtype=np.float64)
    base = float(stride)                       # window (0,0,stride-1,stride-1): w = h = stride
    ctr = 0.5 * (base - 1.0)
    rows = []
    for r in ratios:
        w_r = np.round(np.sqrt(base * base / r))
        h_r = np.round(w_r * r)
        for s in scales:
            ws, hs = w_r * s, h_r * s
            rows.append([ctr - 0.5 * (ws - 1), ctr - 0.5 * (hs - 1),
                         ctr + 0.5 * (ws - 1), ctr + 0.5 * (hs - 1)])
    return np.asarray(rows, dtype=np.float64)


def fpn_anchors(level, start_size=32, aspect_ratios=(0.5, 1, 2), k_min=2):
    """Per-level anchors as built at lib/modeling/FPN.py:343-350."""
    return generate_anchors(2.0 ** level, (start_size * 2.0 ** (level - k_min),), aspect_ratios)


# --------------------------------------------------------------------------- #
# (a4) bbox_transform -- lib/utils/boxes.py:156-205
# --------------------------------------------------------------------------- #
def bbox_transform(boxes, deltas, weights=(1.0, 1.0, 1.0, 1.0), clip=BBOX_XFORM_CLIP):
    if boxes.shape[0] == 0:
        return np.zeros((0, deltas.shape[1]), dtype=deltas.dtype)
    b = boxes.astype(deltas.dtype, copy=False)               # :164 fp64 anchors -> fp32
    w = b[:, 2] - b[:, 0] + 1.0
    h = b[:, 3] - b[:, 1] + 1.0
    cx = b[:, 0] + 0.5 * w
    cy = b[:, 1] + 0.5 * h
    wx, wy, ww, wh = weights
    dx = deltas[:, 0::4] / wx
    dy = deltas[:, 1::4] / wy
    dw = np.minimum(deltas[:, 2::4] / ww, clip)               # :180-181
    dh = np.minimum(deltas[:, 3::4] / wh, clip)
    pcx = dx * w[:, None] + cx[:, None]                       # separate mul, add (no FMA)
    pcy = dy * h[:, None] + cy[:, None]
    pw = np.maximum(np.exp(dw) * w[:, None], 1.0)             # :188-193
    ph = np.maximum(np.exp(dh) * h[:, None], 1.0)
    out = np.zeros(deltas.shape, dtype=deltas.dtype)
    out[:, 0::4] = pcx - 0.5 * pw
    out[:, 1::4] = pcy - 0.5 * ph
    out[:, 2::4] = pcx + 0.5 * pw - 1
    out[:, 3::4] = pcy + 0.5 * ph - 1
    return out


# (a5) clip_tiled_boxes -- lib/utils/boxes.py:138-153 (im_shape = [h, w], fp32)
def clip_tiled_boxes(boxes, im_shape):
    boxes[:, 0::4] = np.maximum(np.minimum(boxes[:, 0::4], im_shape[1] - 1), 0)
    boxes[:, 1::4] = np.maximum(np.minimum(boxes[:, 1::4], im_shape[0] - 1), 0)
    boxes[:, 2::4] = np.maximum(np.minimum(boxes[:, 2::4], im_shape[1] - 1), 0)
    boxes[:, 3::4] = np.maximum(np.minimum(boxes[:, 3::4], im_shape[0] - 1), 0)
    return boxes


# (a6) _filter_boxes -- lib/modeling/generate_proposals.py:171-182
def filter_boxes(boxes, min_size, im_info):
    min_size = min_size * im_info[2]
    ws = boxes[:, 2] - boxes[:, 0] + 1
    hs = boxes[:, 3] - boxes[:, 1] + 1
    xc = boxes[:, 0] + ws / 2.
    yc = boxes[:, 1] + hs / 2.
    return np.where((ws >= min_size) & (hs >= min_size) & (xc < im_info[1]) & (yc < im_info[0]))[0]


# (a7) nms -- lib/utils/boxes.py:329-333 -> lib/utils/cython_nms.pyx:37-87
def nms(dets, thresh):
    if dets.shape[0] == 0:
        return []
    d = np.ascontiguousarray(dets, dtype=np.float32)
    keep = np.empty(d.shape[0], dtype=np.int64)
    n = lib().orc_nms(_fp(d), d.shape[0], np.float32(thresh), keep.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)))
    return keep[:n].copy()


def topk_order(scores, k):
    """Indices of the k largest scores, best first (generate_proposals.py:131-139).
    Ties: the reference's argpartition/argsort are unstable; the oracle and the
    CUDA path both define ties as lower flat index first."""
    s = scores.ravel()
    order = np.argsort(-s, kind='stable')
    if k <= 0 or k >= s.size:
        return order
    return order[:k]


# (a3) proposals_for_one_image -- lib/modeling/generate_proposals.py:104-168
def proposals_for_one_image(im_info, all_anchors, bbox_deltas, scores,
                            pre_nms_topN, post_nms_topN, nms_thresh, min_size):
    deltas = bbox_deltas.transpose((1, 2, 0)).reshape((-1, 4))      # (4A,H,W)->(H*W*A,4)
    sc = scores.transpose((1, 2, 0)).reshape((-1, 1))
    order = topk_order(sc, pre_nms_topN)
    deltas, anchors, sc = deltas[order, :], all_anchors[order, :], sc[order]
    props = bbox_transform(anchors, deltas, (1.0, 1.0, 1.0, 1.0))
    props = clip_tiled_boxes(props, im_info[:2])
    keep = filter_boxes(props, min_size, im_info)
    props, sc = props[keep, :], sc[keep]
    if nms_thresh > 0:
        keep = nms(np.hstack((props, sc)), nms_thresh)
        if post_nms_topN > 0:
            keep = keep[:post_nms_topN]
        props, sc = props[keep, :], sc[keep]
    return props, sc


def all_anchors_for_level(anchors, H, W, feat_stride):
    """(H*W*A, 4) float64 shifted anchors in (H, W, A) order (generate_proposals.py:69-89)."""
    sx = np.arange(0, W) * feat_stride
    sy = np.arange(0, H) * feat_stride
    sx, sy = np.meshgrid(sx, sy, copy=False)
    shifts = np.vstack((sx.ravel(), sy.ravel(), sx.ravel(), sy.ravel())).transpose()
    A, K = anchors.shape[0], shifts.shape[0]
    return (anchors[np.newaxis, :, :] + shifts[:, np.newaxis, :]).reshape((K * A, 4))


# (a2) GenerateProposalsOp.forward -- lib/modeling/generate_proposals.py:20-102
def generate_proposals(rpn_cls_prob, rpn_bbox_pred, im_info, anchors, spatial_scale,
                       pre_nms_topN, post_nms_topN, nms_thresh, min_size):
    scores = np.asarray(rpn_cls_prob, dtype=np.float32)
    deltas = np.asarray(rpn_bbox_pred, dtype=np.float32)
    if np.any(np.isnan(deltas)):
        raise ValueError('bbox_deltas nan')                         # :62-63
    im_info = np.asarray(im_info, dtype=np.float32)
    H, W = scores.shape[-2:]
    all_anchors = all_anchors_for_level(anchors, H, W, 1. / spatial_scale)
    rois = np.empty((0, 5), dtype=np.float32)
    probs = np.empty((0, 1), dtype=np.float32)
    for i in range(scores.shape[0]):
        b, p = proposals_for_one_image(im_info[i], all_anchors, deltas[i], scores[i],
                                       pre_nms_topN, post_nms_topN, nms_thresh, min_size)
        bi = i * np.ones((b.shape[0], 1), dtype=np.float32)
        rois = np.append(rois, np.hstack((bi, b)), axis=0)
        probs = np.append(probs, p, axis=0)
    return rois, probs


# (a8) collect -- lib/modeling/collect_and_distribute_fpn_rpn_proposals.py:91-106
def collect(roi_inputs, score_inputs, post_nms_topN):
    rois = np.concatenate(roi_inputs)
    scores = np.concatenate(score_inputs).reshape(-1)
    inds = np.argsort(-scores, kind='stable')[:post_nms_topN]
    return rois[inds, :]


# (a9) map_rois_to_fpn_levels -- lib/utils/fpn.py:11-28 (+ boxes_area, boxes.py:58-69)
def map_rois_to_fpn_levels(rois, k_min=2, k_max=5, s0=224, lvl0=4):
    w = rois[:, 2] - rois[:, 0] + 1
    h = rois[:, 3] - rois[:, 1] + 1
    areas = w * h
    areas[areas < 0] = 0
    s = np.sqrt(areas)
    lv = np.floor(lvl0 + np.log2(s / s0 + 1e-6))
    return np.clip(lv, k_min, k_max)


# (a10) distribute -- collect_and_distribute_fpn_rpn_proposals.py:109-138
def distribute(rois, k_min=2, k_max=5, prefix='rois'):
    lvls = map_rois_to_fpn_levels(rois[:, 1:5], k_min, k_max)
    out = {prefix: rois}
    order = np.empty((0,))
    for lvl in range(k_min, k_max + 1):
        idx = np.where(lvls == lvl)[0]
        out['%s_fpn%d' % (prefix, lvl)] = rois[idx, :]
        order = np.concatenate((order, idx))
    out[prefix + '_idx_restore_int32'] = np.argsort(order).astype(np.int32)
    return out


# --------------------------------------------------------------------------- #
# (f1) box-head post-processing -- lib/core/test.py:166-181 (decode) and :733-797
# --------------------------------------------------------------------------- #
def box_decode(boxes, deltas, weights=(10.0, 10.0, 5.0, 5.0), im_shape=None):
    """pred_boxes = bbox_transform(boxes, box_deltas, BBOX_REG_WEIGHTS); clip_tiled_boxes(pred_boxes, im.shape)
    (test.py:178-179)."""
    pred = bbox_transform(boxes, deltas, weights)
    return clip_tiled_boxes(pred, im_shape) if im_shape is not None else pred


def box_results_with_nms_and_limit(scores, boxes, num_classes, score_thresh=0.05, nms_thresh=0.3,
                                   detections_per_im=100, num_det_per_class=0):
    """test.py:733-797 without the (default-off) Soft-NMS / box-voting branches."""
    cls_boxes = [[] for _ in range(num_classes)]
    for j in range(1, num_classes):
        inds = np.where(scores[:, j] >= score_thresh)[0]                       # :747
        dets_j = np.hstack((boxes[inds, j * 4:(j + 1) * 4], scores[inds, j][:, np.newaxis])).astype(np.float32, copy=False)
        keep = nms(dets_j, nms_thresh)                                         # :761
        cls_boxes[j] = dets_j[keep, :]
    if detections_per_im > 0:                                                  # :775-784
        image_scores = np.hstack([cls_boxes[j][:, -1] for j in range(1, num_classes)])
        if len(image_scores) > detections_per_im:
            image_thresh = np.sort(image_scores)[-detections_per_im]
            for j in range(1, num_classes):
                keep = np.where(cls_boxes[j][:, -1] >= image_thresh)[0]
                cls_boxes[j] = cls_boxes[j][keep, :]
    if num_det_per_class > 0:                                                  # :785-788
        for j in range(1, num_classes):
            keep = np.argsort(-cls_boxes[j][:, -1])[:num_det_per_class]
            cls_boxes[j] = cls_boxes[j][keep, :]
    im_results = np.vstack([cls_boxes[j] for j in range(1, num_classes)])
    return im_results[:, -1], im_results[:, :-1], cls_boxes


# --------------------------------------------------------------------------- #
# (a14/a15) RoIAlign -- roi_xfrom/roi_align/src/roi_align_kernel.cu:16-121,150-270
# --------------------------------------------------------------------------- #
def roi_align_forward(features, rois, ph, pw, spatial_scale, sampling_ratio, nthreads=0):
    f = np.ascontiguousarray(features, dtype=np.float32)
    r = np.ascontiguousarray(rois, dtype=np.float32)
    N, C, H, W = f.shape
    out = np.zeros((r.shape[0], C, ph, pw), dtype=np.float32)
    if r.shape[0]:
        lib().orc_roialign_fwd(_fp(f), N, C, H, W, _fp(r), r.shape[0], ph, pw,
                               np.float32(spatial_scale), sampling_ratio, _fp(out), nthreads)
    return out


def roi_align_backward(top_grad, rois, feat_shape, ph, pw, spatial_scale, sampling_ratio, nthreads=0):
    t = np.ascontiguousarray(top_grad, dtype=np.float32)
    r = np.ascontiguousarray(rois, dtype=np.float32)
    N, C, H, W = feat_shape
    g = np.zeros((N, C, H, W), dtype=np.float32)
    if r.shape[0]:
        lib().orc_roialign_bwd(_fp(t), N, C, H, W, _fp(r), r.shape[0], ph, pw,
                               np.float32(spatial_scale), sampling_ratio, _fp(g), nthreads)
    return g


# --------------------------------------------------------------------------- #
# (f4) FlowAlign -- lib_vos/vos_model/flow_align/src/flow_align_cuda_kernel.cu:15-117
# --------------------------------------------------------------------------- #
def flow_align_forward(features, flows, nthreads=0):
    f = np.ascontiguousarray(features, dtype=np.float32)
    fl = np.ascontiguousarray(flows, dtype=np.float32)
    N, C, H, W = f.shape
    assert fl.shape == (N, 2, H, W)
    out = np.zeros_like(f)
    if f.size:
        lib().orc_flow_align_fwd(_fp(f), _fp(fl), N, C, H, W, _fp(out), nthreads)
    return out


def flow_align_backward(top_grad, features, flows, nthreads=0):
    """-> (grad_feature (N,C,H,W), grad_flow (N,2,H,W)), functions/flow_align.py:33-50."""
    t = np.ascontiguousarray(top_grad, dtype=np.float32)
    f = np.ascontiguousarray(features, dtype=np.float32)
    fl = np.ascontiguousarray(flows, dtype=np.float32)
    N, C, H, W = f.shape
    gf = np.zeros_like(f)
    gfl = np.zeros((N, 2, H, W), dtype=np.float32)
    if f.size:
        lib().orc_flow_align_bwd(_fp(t), _fp(f), _fp(fl), N, C, H, W, _fp(gf), _fp(gfl), nthreads)
    return gf, gfl


def roi_feature_transform(blobs_in, rpn_ret, blob_rois, resolution, spatial_scales, sampling_ratio,
                          k_min=2, k_max=5, nthreads=0):
    """lib/modeling/model_builder.py:252-303 (FPN branch, method='RoIAlign');
    ``blobs_in`` / ``spatial_scales`` are coarsest-first like the reference."""
    outs = []
    for lvl in range(k_min, k_max + 1):
        r = rpn_ret['%s_fpn%d' % (blob_rois, lvl)]
        if len(r):
            outs.append(roi_align_forward(blobs_in[k_max - lvl], r, resolution, resolution,
                                          spatial_scales[k_max - lvl], sampling_ratio, nthreads))
    shuffled = np.concatenate(outs, axis=0)
    return shuffled[rpn_ret[blob_rois + '_idx_restore_int32'].astype(np.int64)]


# --------------------------------------------------------------------------- #
# (a17/a18) mask paste -- lib/core/test.py:801-855, lib/utils/boxes.py:242-258
# --------------------------------------------------------------------------- #
def expand_boxes(boxes, scale):
    """fp32 arithmetic, stored to a float64 array (boxes.py:242-258)."""
    w_half = (boxes[:, 2] - boxes[:, 0]) * .5
    h_half = (boxes[:, 3] - boxes[:, 1]) * .5
    x_c = (boxes[:, 2] + boxes[:, 0]) * .5
    y_c = (boxes[:, 3] + boxes[:, 1]) * .5
    w_half *= scale
    h_half *= scale
    out = np.zeros(boxes.shape)
    out[:, 0] = x_c - w_half
    out[:, 2] = x_c + w_half
    out[:, 1] = y_c - h_half
    out[:, 3] = y_c + h_half
    return out


def resize_linear(src, dw, dh):
    s = np.ascontiguousarray(src, dtype=np.float32)
    dst = np.empty((dh, dw), dtype=np.float32)
    lib().orc_resize_linear(_fp(s), s.shape[0], s.shape[1], _fp(dst), dh, dw)
    return dst


def paste_masks(masks, cls, ref_boxes, im_h, im_w, thresh=0.5, cls_specific=True,
                want_prob=False, resize=None):
    """Device-layout restatement of segm_results (test.py:801-855) without the RLE
    step: ``masks`` (R,K,M,M) fp32, ``cls`` (R,) class of each detection in the
    reference's class-major running order, ``ref_boxes`` (R,4) fp32.
    Returns (R, im_h, im_w) uint8 [and the fp32 probabilities the reference
    thresholds at :832, pasted the same way, when ``want_prob``]."""
    resize = resize or resize_linear
    R, K, M, _ = masks.shape
    scale = (M + 2.0) / M
    boxes = expand_boxes(ref_boxes, scale).astype(np.int32)           # truncation toward zero
    out = np.zeros((R, im_h, im_w), dtype=np.uint8)
    prob = np.zeros((R, im_h, im_w), dtype=np.float32) if want_prob else None
    padded = np.zeros((M + 2, M + 2), dtype=np.float32)
    for i in range(R):
        padded[1:-1, 1:-1] = masks[i, int(cls[i]) if cls_specific else 0]
        x0b, y0b, x1b, y1b = (int(v) for v in boxes[i])
        w = max(x1b - x0b + 1, 1)
        h = max(y1b - y0b + 1, 1)
        m = resize(padded, w, h)
        x_0, x_1 = max(x0b, 0), min(x1b + 1, im_w)
        y_0, y_1 = max(y0b, 0), min(y1b + 1, im_h)
        if x_1 <= x_0 or y_1 <= y_0:
            continue
        sub = m[(y_0 - y0b):(y_1 - y0b), (x_0 - x0b):(x_1 - x0b)]
        out[i, y_0:y_1, x_0:x_1] = (sub > thresh).astype(np.uint8)
        if want_prob:
            prob[i, y_0:y_1, x_0:x_1] = sub
    return (out, prob) if want_prob else out


# --------------------------------------------------------------------------- #
# COCO RLE (pycocotools): the last line of segm_results' loop body,
#   rle = mask_util.encode(np.array(im_mask[:, :, np.newaxis], order='F'))[0]   (lib/core/test.py:843-846)
# pycocotools is a third-party dependency the reference does not vendor or pin (README.md:63-74) and it is
# absent from this image, so this is a restatement of its published algorithm, cocoapi common/maskApi.c:
# rleEncode (runs of the column-major pixel sequence, the first run counts zeros), rleToString / rleFrString
# (per run, for i > 2 the difference to the run two back; 5-bit groups, bit 5 = continuation, sign-extended
# from bit 4 of the last group; + 48 to land in printable ASCII).  PARITY UNPINNED for the string form: no
# golden vector is available offline; the tests pin it by round trip (rle_from_string . rle_to_string == id,
# rle_decode . rle_counts == id) and against the dense paste, which IS pinned to the reference.
# --------------------------------------------------------------------------- #
def rle_counts(mask):
    """maskApi.c rleEncode for one (h, w) uint8 mask: list of run lengths, Fortran order, zeros first."""
    flat = np.asarray(mask, dtype=np.uint8).ravel(order='F')
    counts, p, c = [], 0, 0
    for v in flat.tolist():
        if v != p:
            counts.append(c)
            c = 0
            p = v
        c += 1
    counts.append(c)
    return counts


def rle_counts_fast(mask):
    """Vectorised rle_counts (same result; the loop above is the restatement, this is for full-size frames)."""
    flat = np.asarray(mask, dtype=np.uint8).ravel(order='F')
    change = np.flatnonzero(flat[1:] != flat[:-1]) + 1
    bounds = np.concatenate(([0], change, [flat.size]))
    counts = np.diff(bounds).tolist()
    if flat.size and flat[0] != 0:
        counts = [0] + counts
    return counts


def rle_to_string(counts):
    """maskApi.c rleToString."""
    out = []
    for i, x in enumerate(counts):
        x = int(x)
        if i > 2:
            x -= int(counts[i - 2])
        more = True
        while more:
            c = x & 0x1f
            x >>= 5
            more = (x != -1) if (c & 0x10) else (x != 0)
            if more:
                c |= 0x20
            out.append(chr(c + 48))
    return ''.join(out)


def rle_from_string(s):
    """maskApi.c rleFrString."""
    counts, p = [], 0
    while p < len(s):
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


def rle_decode(counts, h, w):
    """maskApi.c rleDecode: (h, w) uint8 mask from Fortran-order run lengths."""
    flat = np.zeros(h * w, dtype=np.uint8)
    p, v = 0, 0
    for c in counts:
        if v:
            flat[p:p + c] = 1
        p += c
        v ^= 1
    assert p == h * w
    return flat.reshape((h, w), order='F')


def rle_encode(mask):
    h, w = mask.shape
    return {'size': [int(h), int(w)], 'counts': rle_to_string(rle_counts_fast(mask))}



# --------------------------------------------------------------------------- #
# (f1/f2 extras) lib_vos/tools/vos_test.py: box_results_with_nms_and_limit :748-865,
# bb_intersection_over_union :961-982, nms_with_mask_iou :985-1029, iou_half_numpy :953-959
# --------------------------------------------------------------------------- #
def bb_intersection_over_union(boxA, boxB):
    xA, yA = max(boxA[0], boxB[0]), max(boxA[1], boxB[1])
    xB, yB = min(boxA[2], boxB[2]), min(boxA[3], boxB[3])
    interArea = max(0, xB - xA + 1) * max(0, yB - yA + 1)
    boxAArea = (boxA[2] - boxA[0] + 1) * (boxA[3] - boxA[1] + 1)
    boxBArea = (boxB[2] - boxB[0] + 1) * (boxB[3] - boxB[1] + 1)
    return interArea / float(boxAArea + boxBArea - interArea)


def vos_box_results(scores, boxes, num_classes, score_thresh, nms_thresh, detections_per_im, nms_cross_class=0.,
                    num_det_per_class_pre=0, nms_small_box_iou=0., nms_small_box_score_threshold=0.,
                    prev_cls_boxes=None):
    _, _, cls_boxes = box_results_with_nms_and_limit(scores, boxes, num_classes, score_thresh, nms_thresh,
                                                     detections_per_im)
    K = num_classes
    if nms_cross_class > 0.:                                                    # :816-837
        all_dets = np.vstack([cls_boxes[j] for j in range(1, K)])
        class_ids = np.vstack([np.ones(shape=(len(cls_boxes[j]), 1)) * j for j in range(1, K)])
        keep = nms(all_dets, nms_cross_class)
        all_dets, class_ids = all_dets[keep, :], class_ids[keep, :]
        for j in range(1, K):
            cls_boxes[j] = all_dets[np.where(class_ids == j)[0], :]
    if num_det_per_class_pre > 0:                                               # :839-843
        for j in range(1, K):
            cls_boxes[j] = cls_boxes[j][np.argsort(-cls_boxes[j][:, -1])[:num_det_per_class_pre], :]
    if nms_small_box_iou > 0 and prev_cls_boxes is not None:                    # :845-860
        for j in range(1, K):
            if len(prev_cls_boxes[j]) == 1 and not prev_cls_boxes[j][0][-1] < nms_small_box_score_threshold:
                prev = prev_cls_boxes[j][0][:-1]
                rm = [i for i in range(len(cls_boxes[j]) - 1, -1, -1)
                      if bb_intersection_over_union(prev, cls_boxes[j][i][:-1]) < nms_small_box_iou]
                cls_boxes[j] = np.delete(cls_boxes[j], rm, 0)
    im_results = np.vstack([cls_boxes[j] for j in range(1, K)])
    return im_results[:, -1], im_results[:, :-1], cls_boxes


def mask_iou_greedy(masks_sorted, iou_th):
    """masks_sorted: sequence of equal-shape uint8 masks in descending score order -> removed flags per position
    (vos_test.py:1000-1010 with iou_half_numpy :953-959, float64 arithmetic)."""
    n = len(masks_sorted)
    flat = [np.asarray(m, dtype=np.uint8).ravel() for m in masks_sorted]
    area = [int(f.sum()) for f in flat]
    removed = np.zeros(n, dtype=np.int32)
    for i in range(n):
        if removed[i]:
            continue
        for j in range(i + 1, n):
            inter = int(np.sum(flat[i] & flat[j]))
            if inter / (area[i] + 1e-6) > iou_th or inter / (area[j] + 1e-6) > iou_th:
                removed[j] = 1
    return removed


def nms_with_mask_iou(cls_boxes, cls_segms, num_classes, iou_th=0.9, max_per_class=1):
    box_list = [np.asarray(b).reshape(-1, 5) for b in cls_boxes if len(b) > 0]
    if not box_list:
        return cls_boxes, cls_segms
    boxes = np.concatenate(box_list)
    segms = [s for sl in cls_segms for s in sl]
    classes = np.array([j for j in range(len(cls_boxes)) for _ in range(len(cls_boxes[j]))])
    order = np.argsort(-boxes[:, -1])
    masks = [rle_decode(rle_from_string(segms[k]['counts']), *segms[k]['size']) for k in order]
    keep = np.flatnonzero(mask_iou_greedy(masks, iou_th) == 0)
    out_b = [[] for _ in range(num_classes)]
    out_s = [[] for _ in range(num_classes)]
    for k in keep:
        c = int(classes[order[k]])
        if len(out_b[c]) < max_per_class:
            out_b[c].append(boxes[order[k], :])
            out_s[c].append(rle_encode(masks[k]))
    return out_b, out_s



# --------------------------------------------------------------------------- #
# (f3) label assignment pieces: bbox_overlaps is the reference's own cython_bbox (oracle/_ref);
# bbox_transform_inv lib/utils/boxes.py:208-239, _compute_targets / _expand_bbox_targets
# lib/roi_data/fast_rcnn.py:216-260, outside weights :206-208
# --------------------------------------------------------------------------- #
def bbox_transform_inv(boxes, gt_boxes, weights=(1.0, 1.0, 1.0, 1.0)):
    ex_widths = boxes[:, 2] - boxes[:, 0] + 1.0
    ex_heights = boxes[:, 3] - boxes[:, 1] + 1.0
    ex_ctr_x = boxes[:, 0] + 0.5 * ex_widths
    ex_ctr_y = boxes[:, 1] + 0.5 * ex_heights
    gt_widths = gt_boxes[:, 2] - gt_boxes[:, 0] + 1.0
    gt_heights = gt_boxes[:, 3] - gt_boxes[:, 1] + 1.0
    gt_ctr_x = gt_boxes[:, 0] + 0.5 * gt_widths
    gt_ctr_y = gt_boxes[:, 1] + 0.5 * gt_heights
    wx, wy, ww, wh = weights
    return np.vstack((wx * (gt_ctr_x - ex_ctr_x) / ex_widths, wy * (gt_ctr_y - ex_ctr_y) / ex_heights,
                      ww * np.log(gt_widths / ex_widths), wh * np.log(gt_heights / ex_heights))).transpose()


def bbox_targets(ex_rois, gt_rois, labels, num_classes, weights=(10., 10., 5., 5.), class_agnostic=False):
    t = bbox_transform_inv(ex_rois, gt_rois, weights).astype(np.float32, copy=False)
    clss = np.asarray(labels).copy()
    K = num_classes
    if class_agnostic:
        K = 2
        clss = clss.clip(max=1)
    targets = np.zeros((clss.size, 4 * K), dtype=np.float32)
    inside = np.zeros_like(targets)
    for i in np.where(clss > 0)[0]:
        c = int(clss[i])
        targets[i, 4 * c:4 * c + 4] = t[i]
        inside[i, 4 * c:4 * c + 4] = 1.0
    return targets, inside, np.array(inside > 0, dtype=np.float32)


# --------------------------------------------------------------------------- #
# (f3 / a11) label assignment: add_proposals (datasets/json_dataset.py:413-427 -> _merge_proposal_boxes_into_roidb
# :429-490, _add_class_assignments :513-532), _sample_rois (roi_data/fast_rcnn.py:132-213), add_fast_rcnn_blobs
# (:108-129), _add_multilevel_rois (:262-290), the mask_rois / roi_has_mask part of add_mask_rcnn_blobs
# (roi_data/mask_rcnn.py:34-102).  RNG contract: one uniform key per box; "npr.choice(inds, size, replace=False)" =
# the `size` candidates with the smallest keys in ascending key order (ties: lower index first).
# --------------------------------------------------------------------------- #
def bbox_overlaps(boxes, query):
    """utils.cython_bbox.bbox_overlaps: the reference's own compiled .pyx when oracle/_ref has it, else its restatement
    (float64 areas / union, float32 products -- see csrc/overlaps.cu)."""
    boxes = np.ascontiguousarray(boxes, np.float32)
    query = np.ascontiguousarray(query, np.float32)
    try:
        import ref_harness
        mod = ref_harness._load_ext("cython_bbox", "cython_bbox")
        return mod.bbox_overlaps(boxes, query)
    except Exception:  # noqa: BLE001
        pass
    N, K = boxes.shape[0], query.shape[0]
    out = np.zeros((N, K), np.float32)
    for k in range(K):
        qa = np.float32((query[k, 2] - query[k, 0] + 1) * (query[k, 3] - query[k, 1] + 1))
        for n in range(N):
            iw = np.float32(min(boxes[n, 2], query[k, 2]) - max(boxes[n, 0], query[k, 0]) + 1)
            if iw > 0:
                ih = np.float32(min(boxes[n, 3], query[k, 3]) - max(boxes[n, 1], query[k, 1]) + 1)
                if ih > 0:
                    ua = np.float32(np.float32((boxes[n, 2] - boxes[n, 0] + 1) * (boxes[n, 3] - boxes[n, 1] + 1)) + qa - iw * ih)
                    out[n, k] = iw * ih / ua
    return out


def add_proposals(gt_boxes, gt_classes, rois, im_scale, batch_idx):
    """One image of add_proposals: returns (boxes, max_overlaps, max_classes, box_to_gt_ind_map) of the roidb entry
    after the merge: the G ground-truth rows first, then the image's proposals in original-image coordinates."""
    gt_boxes = np.asarray(gt_boxes, np.float32)
    gt_classes = np.asarray(gt_classes, np.int32)
    G = gt_boxes.shape[0]
    inv = np.float32(1.) / np.float32(im_scale)
    props = (rois[rois[:, 0] == batch_idx, 1:] * inv).astype(np.float32)
    n = props.shape[0]
    mx = np.zeros(n, np.float32)
    cls = np.zeros(n, np.int32)
    b2g = -np.ones(n, np.int32)
    if G > 0 and n > 0:
        ov = bbox_overlaps(props, gt_boxes)
        am, m = ov.argmax(axis=1), ov.max(axis=1)
        pos = m > 0
        mx[pos], cls[pos], b2g[pos] = m[pos], gt_classes[am[pos]], am[pos].astype(np.int32)
    return (np.concatenate([gt_boxes, props]), np.concatenate([np.ones(G, np.float32), mx]),
            np.concatenate([gt_classes, cls]), np.concatenate([np.arange(G, dtype=np.int32), b2g]))


def choice_by_keys(inds, size, keys):
    inds = np.asarray(inds)
    return inds[np.argsort(keys[inds], kind='stable')[:int(size)]]


def sample_rois(boxes, max_overlaps, max_classes, box_to_gt, gt_boxes, keys, im_scale, batch_idx, num_classes,
                rois_per_image=512, fg_fraction=0.25, fg_thresh=0.5, bg_hi=0.5, bg_lo=0.0,
                weights=(10., 10., 5., 5.)):
    """_sample_rois for one image -> dict(labels_int32, rois, bbox_targets, bbox_inside_weights, bbox_outside_weights,
    keep_inds, num_fg)."""
    fg_per = int(np.round(fg_fraction * rois_per_image))
    fg = np.where(max_overlaps >= fg_thresh)[0]
    nfg = min(fg_per, fg.size)
    if fg.size > 0:
        fg = choice_by_keys(fg, nfg, keys)
    bg = np.where((max_overlaps < bg_hi) & (max_overlaps >= bg_lo))[0]
    nbg = min(rois_per_image - nfg, bg.size)
    if bg.size > 0:
        bg = choice_by_keys(bg, nbg, keys)
    keep = np.append(fg, bg).astype(np.int64)
    labels = max_classes[keep].copy()
    labels[nfg:] = 0
    sb = boxes[keep]
    gt_assign = box_to_gt[keep]
    t, iw, ow = bbox_targets(sb, np.asarray(gt_boxes, np.float32)[gt_assign, :], labels, num_classes, weights)
    rois = np.hstack((batch_idx * np.ones((sb.shape[0], 1), np.float32), sb * np.float32(im_scale))).astype(np.float32)
    return {"labels_int32": labels.astype(np.int32), "rois": rois, "bbox_targets": t, "bbox_inside_weights": iw,
            "bbox_outside_weights": ow, "keep_inds": keep, "num_fg": nfg, "sampled_boxes": sb}


def mask_rois_of(sample, im_scale, batch_idx):
    """mask_rois / roi_has_mask_int32 of add_mask_rcnn_blobs (the masks_int32 rasterisation is pycocotools': not here)."""
    labels = sample["labels_int32"]
    fg = np.where(labels > 0)[0]
    has = (labels > 0).astype(np.int32)
    if fg.size > 0:
        rf = sample["sampled_boxes"][fg].copy()
    else:
        rf = sample["sampled_boxes"][np.where(labels == 0)[0][0]].reshape((1, -1)).copy()
        has[0] = 1
    rf = rf * np.float32(im_scale)
    return np.hstack((batch_idx * np.ones((rf.shape[0], 1), np.float32), rf)).astype(np.float32), has


def add_fast_rcnn_blobs(samples, k_min=2, k_max=5, mask_rois=None):
    """Concatenation over the minibatch + _add_multilevel_rois."""
    blobs = {k: np.concatenate([s[k] for s in samples]) for k in
             ("labels_int32", "rois", "bbox_targets", "bbox_inside_weights", "bbox_outside_weights")}
    blobs.update({k: v for k, v in distribute(blobs["rois"], k_min, k_max, 'rois').items() if k != 'rois'})
    if mask_rois is not None:
        blobs["mask_rois"] = np.concatenate([m[0] for m in mask_rois])
        blobs["roi_has_mask_int32"] = np.concatenate([m[1] for m in mask_rois])
        blobs.update({k: v for k, v in distribute(blobs["mask_rois"], k_min, k_max, 'mask_rois').items() if k != 'mask_rois'})
    return blobs


# --------------------------------------------------------------------------- #
# (f3) RPN label assignment: get_field_of_anchors (roi_data/data_utils.py:50-102), _get_rpn_blobs
# (roi_data/rpn.py:143-270).  RNG contract: npr.choice(fg_inds, size, replace=False) = the `size` candidates with the
# smallest per-anchor keys (keys indexed by the anchor's position in the whole field); npr.randint(n, size=k) =
# floor(u[:k] * n) of the image's uniforms u (float64 product).
# --------------------------------------------------------------------------- #
def field_of_anchors(stride, sizes, aspect_ratios, train_max_size, coarsest_stride=32):
    """-> ((field*field*A, 4) float32 in (y, x, anchor) order, A, field)."""
    cell = generate_anchors(stride=stride, sizes=sizes, aspect_ratios=aspect_ratios)
    A = cell.shape[0]
    field = int(np.ceil(coarsest_stride * np.ceil(train_max_size / float(coarsest_stride)) / float(stride)))
    out = np.zeros((field, field, A, 4), np.float64)
    for y in range(field):
        for x in range(field):
            out[y, x] = cell + np.array([x * stride, y * stride, x * stride, y * stride], np.float64)
    return out.reshape(-1, 4).astype(np.float32), A, field


def rpn_labels(im_height, im_width, all_anchors, gt_boxes, keys, rand_bg, batch=256, fg_fraction=0.5, pos=0.7,
               neg=0.3, straddle=0):
    """_get_rpn_blobs before the per-field split -> (labels (T) int32, bbox_targets, inside, outside weights (T,4))."""
    T = all_anchors.shape[0]
    if straddle >= 0:
        ins = np.where((all_anchors[:, 0] >= -straddle) & (all_anchors[:, 1] >= -straddle)
                       & (all_anchors[:, 2] < im_width + straddle) & (all_anchors[:, 3] < im_height + straddle))[0]
    else:
        ins = np.arange(T)
    anchors = all_anchors[ins]
    n = len(ins)
    labels = np.full(n, -1, np.int32)
    a2g_max = None
    gt_boxes = np.asarray(gt_boxes, np.float32).reshape(-1, 4)
    if len(gt_boxes) > 0:
        ov = bbox_overlaps(anchors, gt_boxes)
        a2g_arg = ov.argmax(axis=1)
        a2g_max = ov[np.arange(n), a2g_arg]
        g2a_max = ov.max(axis=0)
        labels[np.where(ov == g2a_max)[0]] = 1
        labels[a2g_max >= pos] = 1
    num_fg = int(fg_fraction * batch)
    fg = np.where(labels == 1)[0]
    if len(fg) > num_fg:
        drop = np.argsort(keys[ins[fg]], kind='stable')[:len(fg) - num_fg]
        labels[fg[drop]] = -1
    fg = np.where(labels == 1)[0]
    num_bg = batch - int(np.sum(labels == 1))
    bg = np.where(a2g_max < neg)[0] if a2g_max is not None else np.arange(n)
    if len(bg) > num_bg:
        posn = np.floor(np.asarray(rand_bg[:num_bg], np.float64) * len(bg)).astype(np.int64)
        labels[bg[posn]] = 0
    targets = np.zeros((n, 4), np.float32)
    if len(fg) > 0:
        targets[fg] = bbox_transform_inv(anchors[fg], gt_boxes[a2g_arg[fg]]).astype(np.float32, copy=False)
    inside = np.zeros((n, 4), np.float32)
    inside[labels == 1] = 1.0
    outside = np.zeros((n, 4), np.float32)
    n_ex = np.sum(labels >= 0)
    if n_ex > 0:
        outside[labels >= 0] = 1.0 / n_ex

    def unmap(d, fill):
        full = np.full((T,) + d.shape[1:], fill, d.dtype)
        full[ins] = d
        return full
    return unmap(labels, -1), unmap(targets, 0), unmap(inside, 0), unmap(outside, 0)


def rpn_blobs_split(fields, labels, targets, inside, outside):
    """The per-field reshape of rpn.py:237-270; fields = [(A, field_size), ...]."""
    out, s = [], 0
    for A, F in fields:
        e = s + F * F * A
        out.append({"rpn_labels_int32_wide": labels[s:e].reshape(1, F, F, A).transpose(0, 3, 1, 2),
                    "rpn_bbox_targets_wide": targets[s:e].reshape(1, F, F, 4 * A).transpose(0, 3, 1, 2),
                    "rpn_bbox_inside_weights_wide": inside[s:e].reshape(1, F, F, 4 * A).transpose(0, 3, 1, 2),
                    "rpn_bbox_outside_weights_wide": outside[s:e].reshape(1, F, F, 4 * A).transpose(0, 3, 1, 2)})
        s = e
    return out


def num_threads():
    return lib().orc_num_threads()
