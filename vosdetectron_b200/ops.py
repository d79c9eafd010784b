"""Tensor-in / tensor-out host wrappers over the C ABI (the ``*_cuda`` fast variants of
SURVEY.md section 8b).  Everything stays on the device; nothing here synchronises.

PyTorch is used for device memory (caching allocator) and the current stream only.
CPU tensors raise NotImplementedError, like the reference's RoIAlignFunction
(lib/modeling/roi_xfrom/roi_align/functions/roi_align.py:29-30): there is no CPU path.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from ._lib import RpnLevel, c_float_p, c_int_p, vp


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(t, name, dtype=torch.float32):
    if not isinstance(t, torch.Tensor):
        raise TypeError("%s must be a torch.Tensor" % name)
    if not t.is_cuda:
        raise NotImplementedError("%s is a CPU tensor: vosdetectron_b200 has no CPU path" % name)
    if t.dtype != dtype:
        raise TypeError("%s must be %s, got %s" % (name, dtype, t.dtype))
    return t.contiguous()


def _on(t):
    """Context of a library call: the tensor's device is current (for torch AND for the library's CUDA runtime: both
    read cudaGetDevice), so `_stream()` is the current stream OF THAT DEVICE, and the caller's device is restored on
    exit.  Tensors on a non-current GPU therefore launch on their own device behind their own producer."""
    return torch.cuda.device(t.device)


# ----------------------------------------------------------------------------- RoIAlign
def _is_channels_last(t):
    """(N,C,H,W) tensor whose memory order is N,H,W,C (torch.channels_last) and not also plain contiguous."""
    return (isinstance(t, torch.Tensor) and t.is_cuda and t.dim() == 4 and t.dtype == torch.float32
            and t.is_contiguous(memory_format=torch.channels_last) and not t.is_contiguous())


def _nhwc_supported(C, aligned_width, sampling_ratio, levels=()):
    """What vosd_roialign_ml_fwd_nhwc takes (it answers VOSD_ERR_UNSUPPORTED otherwise, include/vosd_b200.h): the
    reference's heads, at most four levels, 16-byte aligned bases (the tensor maps need them)."""
    return (int(sampling_ratio) == 2 and int(aligned_width) in (7, 14, 28) and C % 32 == 0 and len(levels) <= 4
            and all(f.data_ptr() % 16 == 0 for f in levels))


def roi_align_forward(features, rois, aligned_height, aligned_width, spatial_scale, sampling_ratio):
    """(N,C,H,W) x (R,5) -> (R,C,ph,pw); roi_align_kernel.cu:65-121 semantics.  Channels-last features take the
    TMA-fed kernel (no layout conversion) where it applies."""
    if _is_channels_last(features) and _nhwc_supported(features.shape[1], aligned_width, sampling_ratio, [features]):
        r = _need_cuda(rois, "rois")
        if r.dim() != 2 or r.size(1) != 5:
            raise ValueError("rois must be (R,5)")
        return roi_align_ml_forward([features], [spatial_scale], r, None, aligned_height, aligned_width, sampling_ratio)
    f = _need_cuda(features, "features")
    r = _need_cuda(rois, "rois")
    if r.dim() != 2 or r.size(1) != 5:
        raise ValueError("rois must be (R,5)")          # reference shim returns 0 here (roi_align_cuda.c:15-18)
    return roi_align_ml_forward([f], [spatial_scale], r, None, aligned_height, aligned_width, sampling_ratio)


def roi_align_backward(grad_output, rois, feature_size, aligned_height, aligned_width, spatial_scale,
                       sampling_ratio, channels_last=False):
    """grad (R,C,ph,pw) -> (N,C,H,W); roi_align_kernel.cu:195-270 semantics.  The gradient map
    is cleared on the stream by the library (zero_init=1), no separate fill kernel.  ``channels_last``: the gradient
    comes back as a torch.channels_last tensor (see roi_align_ml_backward)."""
    if channels_last:
        return roi_align_ml_backward(grad_output, [feature_size], [spatial_scale], rois, None, aligned_height,
                                     aligned_width, sampling_ratio, channels_last=True)[0]
    g = _need_cuda(grad_output, "grad_output")
    r = _need_cuda(rois, "rois")
    N, C, H, W = (int(v) for v in feature_size)
    gi = torch.empty((N, C, H, W), dtype=torch.float32, device=g.device)
    with _on(g):
        _lib.call("vosd_roialign_bwd", _ptr(g), float(spatial_scale), N, r.size(0), H, W, C,
                  int(aligned_height), int(aligned_width), int(sampling_ratio), _ptr(r), _ptr(gi), 1, _stream())
        return gi


# Scratch of the workspace entry points: one grow-only buffer per (device, stream), so that two calls in flight on
# different streams never share one and a CUDA graph captured after the first call sees a fixed address.
_WORKSPACES = {}


def _workspace(device, nbytes):
    key = (device.index if device.index is not None else torch.cuda.current_device(),
           torch.cuda.current_stream(device).cuda_stream)
    buf = _WORKSPACES.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(int(nbytes * 1.25) + 4096, dtype=torch.uint8, device=device)
        _WORKSPACES[key] = buf
    return buf


def _level_arrays(tensors, scales):
    L = len(tensors)
    ptrs = (vp * L)(*[t.data_ptr() for t in tensors])
    hs = (ctypes.c_int * L)(*[int(t.shape[2]) for t in tensors])
    ws = (ctypes.c_int * L)(*[int(t.shape[3]) for t in tensors])
    sc = (ctypes.c_float * L)(*[float(s) for s in scales])
    return ptrs, hs, ws, sc


def roi_align_ml_forward(level_features, level_scales, rois, roi_level, aligned_height, aligned_width,
                         sampling_ratio, out_index=None):
    """All FPN levels in one launch.  ``level_features[k]`` is the map ``roi_level == k`` reads.
    If every level is a channels-last tensor (memory order N,H,W,C) and the head is one of the reference's
    (sampling_ratio 2, pooled width 7/14/28, C % 32 == 0) the TMA-fed channels-last kernel runs on the maps as
    they are; otherwise the maps are used in (or converted to) the reference's NCHW order."""
    r = _need_cuda(rois, "rois")
    lv = None if roi_level is None else _need_cuda(roi_level, "roi_level", torch.int32)
    oi = None if out_index is None else _need_cuda(out_index, "out_index", torch.int32)
    R = r.size(0)
    if (len(level_features) > 0 and all(_is_channels_last(f) for f in level_features)
            and _nhwc_supported(level_features[0].shape[1], aligned_width, sampling_ratio, level_features)):
        N, C = (int(v) for v in level_features[0].shape[:2])
        if any(tuple(f.shape[:2]) != (N, C) for f in level_features):
            raise ValueError("all levels must share batch size and channel count")
        out = torch.empty((R, C, aligned_height, aligned_width), dtype=torch.float32, device=r.device)
        ptrs, hs, ws, sc = _level_arrays(level_features, level_scales)
        with _on(r):
            try:
                _lib.call("vosd_roialign_ml_fwd_nhwc", ptrs, hs, ws, sc, len(level_features), N, C,
                          int(aligned_height), int(aligned_width), int(sampling_ratio), R, _ptr(r), _ptr(lv),
                          _ptr(oi), _ptr(out), _stream())
                return out
            except _lib.VosdError as e:
                # -3: a shape this kernel does not take (documented in the header: the caller uses the NCHW entry
                # point, still on the GPU); anything else is an error
                if e.status != -3:
                    raise
    feats = [_need_cuda(f, "level_features[%d]" % i) for i, f in enumerate(level_features)]
    N, C = (int(v) for v in feats[0].shape[:2])
    if any(tuple(f.shape[:2]) != (N, C) for f in feats):
        raise ValueError("all levels must share batch size and channel count")
    out = torch.empty((R, C, aligned_height, aligned_width), dtype=torch.float32, device=r.device)
    ptrs, hs, ws, sc = _level_arrays(feats, level_scales)
    with _on(r):
        if len(feats) <= 4:
            # default: plan kernel + persistent TMA-fed kernel; the scratch comes from the caching allocator
            need = int(_lib.load().vosd_roialign_fwd_workspace_bytes(hs, ws, len(feats), N, C, int(aligned_height),
                                                                     int(aligned_width), R))
            wsp = _workspace(r.device, max(need, 256))
            _lib.call("vosd_roialign_ml_fwd_ws", ptrs, hs, ws, sc, len(feats), N, C, int(aligned_height),
                      int(aligned_width), int(sampling_ratio), R, _ptr(r), _ptr(lv), _ptr(oi), _ptr(out), _ptr(wsp),
                      wsp.numel(), _stream())
            return out
        _lib.call("vosd_roialign_ml_fwd", ptrs, hs, ws, sc, len(feats), C, int(aligned_height), int(aligned_width),
                  int(sampling_ratio), R, _ptr(r), _ptr(lv), _ptr(oi), _ptr(out), _stream())
        return out


def roi_align_ml_backward(grad_output, level_shapes, level_scales, rois, roi_level, aligned_height,
                          aligned_width, sampling_ratio, out_index=None, channels_last=False):
    """grad (R,C,ph,pw) -> one gradient map per level.  ``channels_last=True`` (the maps of the forward were
    torch.channels_last tensors): the gradients are accumulated directly in that memory order by the channels-last
    kernel (heads of the reference: sampling_ratio 2, pooled sizes multiples of 7, C % 32 == 0); other heads are
    accumulated in NCHW order and converted."""
    g = _need_cuda(grad_output, "grad_output")
    r = _need_cuda(rois, "rois")
    lv = None if roi_level is None else _need_cuda(roi_level, "roi_level", torch.int32)
    oi = None if out_index is None else _need_cuda(out_index, "out_index", torch.int32)
    shapes = [tuple(int(v) for v in s) for s in level_shapes]
    if channels_last:
        N, C = shapes[0][:2]
        if (int(sampling_ratio) == 2 and int(aligned_height) % 7 == 0 and int(aligned_width) % 7 == 0 and C % 32 == 0
                and len(shapes) <= _lib.MAX_LEVELS):
            grads = [torch.empty(s, dtype=torch.float32, device=g.device, memory_format=torch.channels_last) for s in shapes]
            ptrs, hs, ws, sc = _level_arrays(grads, level_scales)
            with _on(g):
                _lib.call("vosd_roialign_ml_bwd_nhwc", _ptr(g), ptrs, hs, ws, sc, len(grads), N, C, int(aligned_height),
                          int(aligned_width), int(sampling_ratio), r.size(0), _ptr(r), _ptr(lv), _ptr(oi), 1, _stream())
                return grads
        return [x.contiguous(memory_format=torch.channels_last)
                for x in roi_align_ml_backward(g, shapes, level_scales, r, lv, aligned_height, aligned_width, sampling_ratio, oi)]
    grads = [torch.empty(s, dtype=torch.float32, device=g.device) for s in shapes]
    N, C = grads[0].shape[:2]
    ptrs, hs, ws, sc = _level_arrays(grads, level_scales)
    with _on(g):
        _lib.call("vosd_roialign_ml_bwd", _ptr(g), ptrs, hs, ws, sc, len(grads), N, C, int(aligned_height),
                  int(aligned_width), int(sampling_ratio), r.size(0), _ptr(r), _ptr(lv), _ptr(oi), 1, _stream())
        return grads


# ----------------------------------------------------------------------------- FlowAlign
def _flow_pair(features, flows):
    f = _need_cuda(features, "features")
    fl = _need_cuda(flows, "flows")
    if f.dim() != 4 or fl.dim() != 4:
        raise ValueError("features must be (N,C,H,W) and flows (N,2,H,W)")
    N, C, H, W = f.shape
    if tuple(fl.shape) != (N, 2, H, W):
        # the reference only asserts this (functions/flow_align.py:23-25) and would read out of bounds
        raise ValueError("flows %s does not match features %s" % (tuple(fl.shape), tuple(f.shape)))
    return f, fl


def flow_align_forward(features, flows):
    """(N,C,H,W) x (N,2,H,W) -> (N,C,H,W); flow_align_cuda_kernel.cu:15-55 semantics, bit-identical."""
    f, fl = _flow_pair(features, flows)
    N, C, H, W = f.shape
    out = torch.empty_like(f)
    with _on(f):
        _lib.call("vosd_flow_align_fwd", N, H, W, C, _ptr(f), _ptr(fl), _ptr(out), _stream())
        return out


def flow_align_backward(grad_output, features, flows, want_flow_grad=True):
    """-> (grad_feature, grad_flow or None); flow_align_cuda_kernel.cu:57-117.  The gradients are cleared on the
    stream by the library (zero_init=1).  ``want_flow_grad=False`` skips the flow gradient (and the tap loads and
    arithmetic that only feed it)."""
    f, fl = _flow_pair(features, flows)
    g = _need_cuda(grad_output, "grad_output")
    if g.shape != f.shape:
        raise ValueError("grad_output %s does not match features %s" % (tuple(g.shape), tuple(f.shape)))
    N, C, H, W = f.shape
    gf = torch.empty_like(f)
    gfl = torch.empty_like(fl) if want_flow_grad else None
    with _on(f):
        _lib.call("vosd_flow_align_bwd", N, H, W, C, _ptr(g), _ptr(f), _ptr(fl), _ptr(gf), _ptr(gfl), 1, _stream())
        return gf, gfl


def _ptr_table(tensors):
    return (vp * len(tensors))(*[t.data_ptr() for t in tensors])


def flow_align_ml_forward(level_features, level_flows):
    """Every FPN level in one launch (the loop of vos_model_builder.py:329-335); level k warps
    ``level_features[k]`` by ``level_flows[k]`` (already at that level's resolution)."""
    pairs = [_flow_pair(f, fl) for f, fl in zip(level_features, level_flows)]
    if not pairs:
        return []
    N, C = pairs[0][0].shape[:2]
    if any(f.shape[:2] != (N, C) for f, _ in pairs):
        raise ValueError("all levels must share batch size and channel count")
    outs = [torch.empty_like(f) for f, _ in pairs]
    L = len(pairs)
    hs = (ctypes.c_int * L)(*[int(f.shape[2]) for f, _ in pairs])
    ws = (ctypes.c_int * L)(*[int(f.shape[3]) for f, _ in pairs])
    with _on(pairs[0][0]):
        _lib.call("vosd_flow_align_ml_fwd", L, N, C, hs, ws, _ptr_table([f for f, _ in pairs]),
                  _ptr_table([fl for _, fl in pairs]), _ptr_table(outs), _stream())
        return outs


def flow_align_ml_backward(level_grads, level_features, level_flows, want_flow_grad=True):
    pairs = [_flow_pair(f, fl) for f, fl in zip(level_features, level_flows)]
    if not pairs:
        return [], []
    grads = [_need_cuda(g, "grad_output") for g in level_grads]
    N, C = pairs[0][0].shape[:2]
    if any(f.shape[:2] != (N, C) for f, _ in pairs):
        raise ValueError("all levels must share batch size and channel count")
    if len(grads) != len(pairs) or len(level_features) != len(level_flows):
        raise ValueError("one grad_output and one flow per level")
    for k, (g, (f, _)) in enumerate(zip(grads, pairs)):
        if g.shape != f.shape:
            raise ValueError("grad_output[%d] %s does not match features %s" % (k, tuple(g.shape), tuple(f.shape)))
    gfs = [torch.empty_like(f) for f, _ in pairs]
    gfls = [torch.empty_like(fl) for _, fl in pairs] if want_flow_grad else None
    L = len(pairs)
    hs = (ctypes.c_int * L)(*[int(f.shape[2]) for f, _ in pairs])
    ws = (ctypes.c_int * L)(*[int(f.shape[3]) for f, _ in pairs])
    with _on(pairs[0][0]):
        _lib.call("vosd_flow_align_ml_bwd", L, N, C, hs, ws, _ptr_table(grads), _ptr_table([f for f, _ in pairs]),
                  _ptr_table([fl for _, fl in pairs]), _ptr_table(gfs), None if gfls is None else _ptr_table(gfls), 1,
                  _stream())
        return gfs, gfls


# ----------------------------------------------------------------------------- proposals
def make_rpn_levels(level_inputs):
    """level_inputs: list of (scores (N,A,H,W), deltas (N,4A,H,W), anchors ndarray (A,4) float64,
    feat_stride).  Returns (ctypes array of vosd_rpn_level, list keeping the tensors alive)."""
    L = len(level_inputs)
    arr = (RpnLevel * L)()
    keep = []
    for i, (sc, dl, anchors, stride) in enumerate(level_inputs):
        sc = _need_cuda(sc, "scores")
        dl = _need_cuda(dl, "deltas")
        a = np.ascontiguousarray(anchors, dtype=np.float64)
        N, A, H, W = sc.shape
        if a.shape != (A, 4) or tuple(dl.shape) != (N, 4 * A, H, W):
            raise ValueError("level %d: scores %s, deltas %s, anchors %s are inconsistent"
                             % (i, tuple(sc.shape), tuple(dl.shape), a.shape))
        arr[i].scores = sc.data_ptr()
        arr[i].deltas = dl.data_ptr()
        arr[i].height, arr[i].width, arr[i].num_anchors = H, W, A
        arr[i].feat_stride = float(stride)
        for k, v in enumerate(a.ravel()):
            arr[i].anchors[k] = v
        keep += [sc, dl]
    return arr, keep


def generate_proposals_cuda(level_inputs, im_info, pre_nms_topN, post_nms_topN, nms_thresh, min_size,
                            workspace=None, zero_fill=True):
    """All (level, image) segments in 3 launches.  Returns device tensors
    rois (L,N,cap,5) [image,x1,y1,x2,y2], probs (L,N,cap), count (L,N) int32 (rows >= count are 0; with
    ``zero_fill=False`` they are uninitialised: two fill kernels less for a consumer that honours ``count``, like
    collect_distribute_cuda)."""
    arr, keep = make_rpn_levels(level_inputs)
    L = len(level_inputs)
    N = keep[0].shape[0]
    info = _need_cuda(im_info, "im_info")
    if nms_thresh <= 0:
        post_nms_topN = 0
    lib = _lib.load()
    cap = lib.vosd_proposals_capacity(arr, L, int(pre_nms_topN), int(post_nms_topN))
    if cap < 0:
        _lib.check("vosd_proposals_capacity", cap)
    nbytes = lib.vosd_generate_proposals_workspace_bytes(arr, L, N, int(pre_nms_topN), int(post_nms_topN))
    dev = info.device
    if workspace is None or workspace.numel() < nbytes:
        workspace = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    alloc = torch.zeros if zero_fill else torch.empty
    rois = alloc((L, N, cap, 5), dtype=torch.float32, device=dev)
    probs = alloc((L, N, cap), dtype=torch.float32, device=dev)
    count = torch.empty((L, N), dtype=torch.int32, device=dev)
    with _on(info):
        _lib.call("vosd_generate_proposals", arr, L, N, _ptr(info), int(pre_nms_topN), int(post_nms_topN),
                  float(nms_thresh), float(min_size), _ptr(rois), _ptr(probs), _ptr(count),
                  _ptr(workspace), workspace.numel(), _stream())
        return rois, probs, count


def decode_anchors_cuda(deltas, anchors, feat_stride, im_info):
    """Streaming decode+clip of every anchor: (N,4A,H,W) -> (N, H*W*A, 4)."""
    dl = _need_cuda(deltas, "deltas")
    info = _need_cuda(im_info, "im_info")
    N, A4, H, W = dl.shape
    A = A4 // 4
    lvl = (RpnLevel * 1)()
    lvl[0].scores = None
    lvl[0].deltas = dl.data_ptr()
    lvl[0].height, lvl[0].width, lvl[0].num_anchors = H, W, A
    lvl[0].feat_stride = float(feat_stride)
    for k, v in enumerate(np.ascontiguousarray(anchors, dtype=np.float64).ravel()):
        lvl[0].anchors[k] = v
    out = torch.empty((N, H * W * A, 4), dtype=torch.float32, device=dl.device)
    with _on(dl):
        _lib.call("vosd_decode_anchors", lvl, N, _ptr(info), _ptr(out), _stream())
        return out


def any_nan_cuda(t):
    """Device int32 flag tensor (1 element): 1 if any NaN (generate_proposals.py:62-63)."""
    x = _need_cuda(t, "tensor")
    flag = torch.zeros(1, dtype=torch.int32, device=x.device)
    with _on(x):
        _lib.call("vosd_any_nan", _ptr(x), x.numel(), _ptr(flag), _stream())
        return flag


# ----------------------------------------------------------------------------- NMS
def nms_cuda(dets, thresh):
    """dets (n,5) [x1,y1,x2,y2,score] any order -> (keep int64 (n), num_keep int32 (1)) device
    tensors; keep[:num_keep] are ascending indices (cython_nms.pyx:87)."""
    d = _need_cuda(dets, "dets")
    if d.dim() != 2 or d.size(1) != 5:
        raise ValueError("dets must be (n,5)")
    n = d.size(0)
    keep = torch.empty(max(n, 1), dtype=torch.int64, device=d.device)
    num = torch.empty(1, dtype=torch.int32, device=d.device)
    nbytes = _lib.load().vosd_nms_workspace_bytes(n)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=d.device)
    with _on(d):
        _lib.call("vosd_nms", _ptr(d), n, float(np.float32(thresh)), _ptr(keep), _ptr(num), _ptr(ws), nbytes, _stream())
        return keep, num


# ----------------------------------------------------------------------------- collect / distribute
def collect_distribute_cuda(rois, probs, count, post_nms_topN, images_per_group=1, k_min=2, k_max=5,
                            canonical_scale=224.0, canonical_level=4):
    """Outputs of generate_proposals_cuda -> dict of device tensors (G = N / images_per_group):
    rois (G,post,5), count (G), level (G,post), level_count (G,nl), order (G,post), restore (G,post)."""
    r = _need_cuda(rois, "rois")
    p = _need_cuda(probs, "probs")
    c = _need_cuda(count, "count", torch.int32)
    L, N, cap, _ = r.shape
    G = N // images_per_group
    dev = r.device
    post = int(post_nms_topN)
    out = {
        # (the kernel writes every row: zeros beyond the group's count)
        "rois": torch.empty((G, post, 5), dtype=torch.float32, device=dev),
        "count": torch.empty((G,), dtype=torch.int32, device=dev),
        "level": torch.empty((G, post), dtype=torch.int32, device=dev),
        "level_count": torch.empty((G, k_max - k_min + 1), dtype=torch.int32, device=dev),
        "order": torch.empty((G, post), dtype=torch.int32, device=dev),
        "restore": torch.empty((G, post), dtype=torch.int32, device=dev),
    }
    with _on(r):
        _lib.call("vosd_collect_distribute", _ptr(r), _ptr(p), _ptr(c), L, N, cap, int(images_per_group), post,
                  int(k_min), int(k_max), float(canonical_scale), int(canonical_level),
                  _ptr(out["rois"]), _ptr(out["count"]), _ptr(out["level"]), _ptr(out["level_count"]),
                  _ptr(out["order"]), _ptr(out["restore"]), None, 0, _stream())
        return out


def distribute_cuda(rois, k_min=2, k_max=5, canonical_scale=224.0, canonical_level=4):
    """rois (R,5) -> (level (R), level_count (nl), order (R), restore (R)) int32 device tensors."""
    r = _need_cuda(rois, "rois")
    R = r.size(0)
    dev = r.device
    level = torch.empty(max(R, 1), dtype=torch.int32, device=dev)
    lc = torch.empty(k_max - k_min + 1, dtype=torch.int32, device=dev)
    order = torch.empty(max(R, 1), dtype=torch.int32, device=dev)
    restore = torch.empty(max(R, 1), dtype=torch.int32, device=dev)
    with _on(r):
        _lib.call("vosd_distribute", _ptr(r), R, int(k_min), int(k_max), float(canonical_scale), int(canonical_level),
                  _ptr(level), _ptr(lc), _ptr(order), _ptr(restore), _stream())
        return level[:R], lc, order[:R], restore[:R]


# ----------------------------------------------------------------------------- box head post-processing
def bbox_transform_cuda(boxes, deltas, weights=(1.0, 1.0, 1.0, 1.0), clip_hw=None):
    """boxes (n,4), deltas (n,4k) -> pred_boxes (n,4k): box_utils.bbox_transform (boxes.py:156-205) and, with
    clip_hw = (h, w), clip_tiled_boxes (:138-153) -- the decode of im_detect_bbox (core/test.py:178-179)."""
    b = _need_cuda(boxes, "boxes")
    d = _need_cuda(deltas, "deltas")
    if b.dim() != 2 or b.size(1) != 4 or d.dim() != 2 or d.size(0) != b.size(0) or d.size(1) % 4:
        raise ValueError("boxes must be (n,4) and deltas (n,4k)")
    out = torch.empty_like(d)
    w = (ctypes.c_float * 4)(*[float(x) for x in weights])
    ch, cw = (float(clip_hw[0]), float(clip_hw[1])) if clip_hw is not None else (-1.0, -1.0)
    with _on(b):
        _lib.call("vosd_bbox_transform", _ptr(b), _ptr(d), b.size(0), d.size(1) // 4, w, ch, cw, _ptr(out), _stream())
        return out


def box_results_cuda(scores, boxes, score_thresh=0.05, nms_thresh=0.3, max_per_image=100, rows=None, cap=None):
    """scores (N,R,K), boxes (N,R,4K) -> (dets (N,cap,6) [x1,y1,x2,y2,score,class], count (N,), cls_count (N,K)):
    box_results_with_nms_and_limit (core/test.py:733-797) for N images in four launches, nothing leaves the device.
    ``count`` can exceed ``cap`` (rows beyond it are dropped): pass cap=R*(K-1) for the lossless worst case."""
    s = _need_cuda(scores, "scores")
    b = _need_cuda(boxes, "boxes")
    if s.dim() != 3 or b.dim() != 3 or b.shape[:2] != s.shape[:2] or b.size(2) != 4 * s.size(2):
        raise ValueError("scores must be (N,R,K) and boxes (N,R,4K)")
    N, R, K = (int(v) for v in s.shape)
    if cap is None:
        cap = min(R * (K - 1), 4 * max_per_image) if max_per_image > 0 else R * (K - 1)
    cap = max(int(cap), 1)
    dev = s.device
    dets = torch.zeros((N, cap, 6), dtype=torch.float32, device=dev)
    count = torch.empty((N,), dtype=torch.int32, device=dev)
    cls_count = torch.empty((N, K), dtype=torch.int32, device=dev)
    r = None if rows is None else _need_cuda(rows, "rows", torch.int32)
    with _on(s):
        nbytes = int(_lib.load().vosd_box_results_workspace_bytes(N, R, K))
        ws = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
        _lib.call("vosd_box_results", _ptr(s), _ptr(b), _ptr(r), N, R, K, float(score_thresh), float(nms_thresh),
                  int(max_per_image), cap, _ptr(dets), _ptr(count), _ptr(cls_count), _ptr(ws), nbytes, _stream())
        return dets, count, cls_count


# ----------------------------------------------------------------------------- paste
def paste_masks_cuda(masks, cls, ref_boxes, im_h, im_w, thresh=0.5, want_prob=False):
    """masks (R,K,M,M), cls (R) int32 or None, ref_boxes (R,4) -> uint8 (R,im_h,im_w)
    [, fp32 probabilities (R,im_h,im_w)]."""
    m = _need_cuda(masks, "masks")
    b = _need_cuda(ref_boxes, "ref_boxes")
    c = None if cls is None else _need_cuda(cls, "cls", torch.int32)
    R, K, M, M2 = m.shape
    if M != M2 or b.shape != (R, 4):
        raise ValueError("masks must be (R,K,M,M) and ref_boxes (R,4)")
    out = torch.empty((R, im_h, im_w), dtype=torch.uint8, device=m.device)
    prob = torch.empty((R, im_h, im_w), dtype=torch.float32, device=m.device) if want_prob else None
    with _on(m):
        _lib.call("vosd_paste_masks", _ptr(m), _ptr(c), _ptr(b), R, K, M, int(im_h), int(im_w), float(thresh),
                  _ptr(out), _ptr(prob), _stream())
        return (out, prob) if want_prob else out


def paste_masks_packed_cuda(masks, cls, ref_boxes, im_h, im_w, thresh=0.5, want_dense=True):
    """Paste with the bit-packed copy written by the same kernel: returns (dense uint8 (R,im_h,im_w) or None,
    packed uint8 (R, ceil(im_h*im_w/8)), 8 pixels per byte LSB first)."""
    m = _need_cuda(masks, "masks")
    b = _need_cuda(ref_boxes, "ref_boxes")
    c = None if cls is None else _need_cuda(cls, "cls", torch.int32)
    R, K, M, M2 = m.shape
    if M != M2 or b.shape != (R, 4):
        raise ValueError("masks must be (R,K,M,M) and ref_boxes (R,4)")
    out = torch.empty((R, im_h, im_w), dtype=torch.uint8, device=m.device) if want_dense else None
    packed = torch.empty((R, (im_h * im_w + 7) // 8), dtype=torch.uint8, device=m.device)
    with _on(m):
        _lib.call("vosd_paste_masks_packed", _ptr(m), _ptr(c), _ptr(b), R, K, M, int(im_h), int(im_w), float(thresh),
                  _ptr(out), _ptr(packed), _stream())
        return out, packed


def paste_rle_cuda(masks, cls, ref_boxes, im_h, im_w, thresh=0.5, run_capacity=None, str_capacity=None):
    """Fused paste -> COCO RLE on the device (vosd_paste_rle).  Returns a dict of device tensors:
    runs (arena, uint32 stored as int32), run_offset (R) int64, run_count (R) int32, chars (arena, uint8),
    str_offset (R) int64, str_len (R) int32, status (R) int32, cursors (2) int64 [runs used, chars used].
    Default capacities hold ~8 runs per box column per detection on average; a detection that does not fit
    gets status != 0 (see ``rle_results`` for the retry)."""
    m = _need_cuda(masks, "masks")
    b = _need_cuda(ref_boxes, "ref_boxes")
    c = None if cls is None else _need_cuda(cls, "cls", torch.int32)
    R, K, M, M2 = m.shape
    if M != M2 or b.shape != (R, 4):
        raise ValueError("masks must be (R,K,M,M) and ref_boxes (R,4)")
    dev = m.device
    if run_capacity is None:
        run_capacity = max(1024, R * (8 * min(int(im_w), 512) + 2))
    if str_capacity is None:
        str_capacity = 3 * run_capacity
    out = {
        "runs": torch.empty(int(run_capacity), dtype=torch.int32, device=dev),
        "chars": torch.empty(int(str_capacity), dtype=torch.uint8, device=dev),
        "cursors": torch.empty(2, dtype=torch.int64, device=dev),
        "run_offset": torch.empty(R, dtype=torch.int64, device=dev),
        "run_count": torch.empty(R, dtype=torch.int32, device=dev),
        "str_offset": torch.empty(R, dtype=torch.int64, device=dev),
        "str_len": torch.empty(R, dtype=torch.int32, device=dev),
        "status": torch.empty(R, dtype=torch.int32, device=dev),
    }
    with _on(m):
        _lib.call("vosd_paste_rle", _ptr(m), _ptr(c), _ptr(b), R, K, M, int(im_h), int(im_w), float(thresh),
                  _ptr(out["runs"]), int(run_capacity), _ptr(out["chars"]), int(str_capacity), _ptr(out["cursors"]),
                  _ptr(out["run_offset"]), _ptr(out["run_count"]), _ptr(out["str_offset"]), _ptr(out["str_len"]),
                  _ptr(out["status"]), _stream())
        return out


def rle_results(masks, cls, ref_boxes, im_h, im_w, thresh=0.5):
    """Host view of paste_rle_cuda: list of R COCO RLE dicts {'size': [h, w], 'counts': str} (what
    mask_util.encode returns per mask, counts decoded to str as segm_results does, test.py:847).  One retry with the
    exact capacities if the default arenas were too small."""
    R = int(masks.shape[0])
    if R == 0:
        return []
    out = paste_rle_cuda(masks, cls, ref_boxes, im_h, im_w, thresh)
    used = out["cursors"].cpu().numpy()
    if int(out["status"].max().item()) != 0:
        out = paste_rle_cuda(masks, cls, ref_boxes, im_h, im_w, thresh, run_capacity=int(used[0]),
                             str_capacity=max(int(used[1]), 6 * int(used[0])))
        used = out["cursors"].cpu().numpy()
        if int(out["status"].max().item()) != 0:
            raise RuntimeError("vosd_paste_rle: arena overflow after retry")
    chars = out["chars"][:int(used[1])].cpu().numpy().tobytes()
    so, sl = out["str_offset"].cpu().numpy(), out["str_len"].cpu().numpy()
    return [{'size': [int(im_h), int(im_w)], 'counts': chars[int(so[i]):int(so[i]) + int(sl[i])].decode('ascii')}
            for i in range(R)]


def pack_mask_bits_cuda(masks_u8):
    """(..., H, W) uint8 {0,1} on the device -> (..., ceil(H*W/8)) uint8, 8 pixels per byte (LSB first)."""
    m = _need_cuda(masks_u8, "masks", torch.uint8)
    lead = m.shape[:-2]
    pixels = int(m.shape[-2]) * int(m.shape[-1])
    n = 1
    for v in lead:
        n *= int(v)
    out = torch.empty(tuple(lead) + ((pixels + 7) // 8,), dtype=torch.uint8, device=m.device)
    with _on(m):
        _lib.call("vosd_pack_mask_bits", _ptr(m), n, pixels, _ptr(out), _stream())
        return out


# ----------------------------------------------------------------------------- mask-IoU suppression
def rle_to_bits_cuda(run_lists, pixels, device=None):
    """COCO run lengths of R masks (list of int sequences, alternating 0-runs / 1-runs) -> packed
    (R, ceil(pixels/32)*4) uint8 on the device, 1 bit per pixel in the RLE's column-major order."""
    R = len(run_lists)
    dev = device or torch.device("cuda", torch.cuda.current_device())
    words = (int(pixels) + 31) // 32
    out = torch.empty((R, words * 4), dtype=torch.uint8, device=dev)
    if R == 0 or pixels == 0:
        return out
    counts = np.array([len(r) for r in run_lists], dtype=np.int32)
    offsets = np.concatenate(([0], np.cumsum(counts[:-1], dtype=np.int64))).astype(np.int64)
    flat = np.concatenate([np.asarray(r, dtype=np.int64) for r in run_lists]) if counts.sum() else np.zeros(1, np.int64)
    if flat.min() < 0 or any(int(np.sum(r)) != int(pixels) for r in run_lists):
        raise ValueError("run lengths must be non-negative and sum to the pixel count")
    runs = torch.from_numpy(flat.astype(np.uint32).view(np.int32)).to(dev)
    offs = torch.from_numpy(offsets).to(dev)        # named: a temporary would be recycled before the launch
    cnts = torch.from_numpy(counts).to(dev)
    with _on(out):
        _lib.call("vosd_rle_to_bits", _ptr(runs), _ptr(offs), _ptr(cnts), R, int(pixels), _ptr(out), int(counts.max()),
                  _stream())
        return out


def mask_iou_nms_cuda(packed, order, iou_th):
    """packed (R, bytes) uint8 bit masks, order (R) int32 (mask index per descending-score position, or None) ->
    (removed (R) int32 per POSITION, num_keep (1) int32); vos_test.py:1001-1010 semantics."""
    p = _need_cuda(packed, "packed", torch.uint8)
    if p.dim() != 2:
        raise ValueError("packed must be (R, bytes_per_mask)")
    R, nb = int(p.shape[0]), int(p.shape[1])
    if nb % 4:
        p = torch.nn.functional.pad(p, (0, 4 - nb % 4)).contiguous()       # whole 32-bit words, zero padding bits
        nb = int(p.shape[1])
    o = None if order is None else _need_cuda(order, "order", torch.int32)
    removed = torch.empty((max(R, 1),), dtype=torch.int32, device=p.device)
    num = torch.empty((1,), dtype=torch.int32, device=p.device)
    nbytes = _lib.load().vosd_mask_iou_nms_workspace_bytes(R)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=p.device)
    with _on(p):
        _lib.call("vosd_mask_iou_nms", _ptr(p), R, nb, _ptr(o), float(iou_th), _ptr(removed), _ptr(num), _ptr(ws), nbytes,
                  _stream())
        return removed[:R], num


# ----------------------------------------------------------------------------- bbox_overlaps
def bbox_overlaps_cuda(boxes, query_boxes, want_matrix=True):
    """boxes (N,4), query_boxes (K,4) on the device -> (overlaps (N,K) or None, row_max (N), row_argmax (N) int32);
    cython_bbox.pyx:32-73 arithmetic, np.argmax tie rule."""
    b = _need_cuda(boxes, "boxes")
    q = _need_cuda(query_boxes, "query_boxes")
    if b.dim() != 2 or b.size(1) != 4 or q.dim() != 2 or q.size(1) != 4:
        raise ValueError("boxes must be (N,4) and query_boxes (K,4)")
    N, K = int(b.size(0)), int(q.size(0))
    ov = torch.zeros((N, K), dtype=torch.float32, device=b.device) if want_matrix else None
    mx = torch.zeros((N,), dtype=torch.float32, device=b.device)
    am = torch.zeros((N,), dtype=torch.int32, device=b.device)
    with _on(b):
        _lib.call("vosd_bbox_overlaps", _ptr(b), N, _ptr(q), K, _ptr(ov), _ptr(mx), _ptr(am), _stream())
        return ov, mx, am


def bbox_targets_cuda(ex_rois, gt_rois, labels, num_classes, weights=(10.0, 10.0, 5.0, 5.0), class_agnostic=False):
    """Sampled RoIs (n,4), their assigned gt boxes (n,4), labels (n) int32 -> (bbox_targets, bbox_inside_weights,
    bbox_outside_weights), each (n, 4*K): fast_rcnn.py:216-260 + :206-208 in one launch."""
    e = _need_cuda(ex_rois, "ex_rois")
    g = _need_cuda(gt_rois, "gt_rois")
    lb = _need_cuda(labels, "labels", torch.int32)
    if e.dim() != 2 or e.size(1) != 4 or g.shape != e.shape or lb.shape != (e.size(0),):
        raise ValueError("ex_rois / gt_rois must be (n,4) and labels (n)")
    n = int(e.size(0))
    K = 2 if class_agnostic else int(num_classes)
    outs = [torch.empty((n, 4 * K), dtype=torch.float32, device=e.device) for _ in range(3)]
    w = (ctypes.c_float * 4)(*[float(v) for v in weights])
    with _on(e):
        _lib.call("vosd_bbox_targets", _ptr(e), _ptr(g), _ptr(lb), n, int(num_classes), int(bool(class_agnostic)), w,
                  _ptr(outs[0]), _ptr(outs[1]), _ptr(outs[2]), _stream())
        return tuple(outs)


# ----------------------------------------------------------------------------- label assignment (training)
def label_proposals_cuda(proposals, gt_boxes, gt_classes):
    """One image of add_proposals (datasets/json_dataset.py:413-490 + _add_class_assignments :513-532) on the device:
    proposals (n,4) in ORIGINAL-image coordinates, gt_boxes (G,4), gt_classes (G) int32 (> 0) ->
    (max_overlaps (n) fp32, max_classes (n) int32, box_to_gt_ind_map (n) int32; 0 / 0 / -1 where nothing overlaps)."""
    p = _need_cuda(proposals, "proposals")
    n = int(p.size(0))
    if gt_boxes is None or int(gt_boxes.size(0)) == 0 or n == 0:
        return (torch.zeros(n, dtype=torch.float32, device=p.device), torch.zeros(n, dtype=torch.int32, device=p.device),
                torch.full((n,), -1, dtype=torch.int32, device=p.device))
    g = _need_cuda(gt_boxes, "gt_boxes")
    c = _need_cuda(gt_classes, "gt_classes", torch.int32)
    _, mx, am = bbox_overlaps_cuda(p, g, want_matrix=False)
    pos = mx > 0
    cls = torch.where(pos, c[am.long()], torch.zeros_like(am))
    b2g = torch.where(pos, am, torch.full_like(am, -1))
    return torch.where(pos, mx, torch.zeros_like(mx)), cls, b2g


def sample_rois_cuda(max_overlaps, keys, num_boxes, rois_per_image, fg_rois_per_image, fg_thresh, bg_thresh_hi,
                     bg_thresh_lo):
    """Index selection of _sample_rois (fast_rcnn.py:132-160) for a minibatch in one launch; see vosd_sample_rois for the
    key-based RNG contract.  max_overlaps, keys (B,N) fp32, num_boxes (B) int32 ->
    (keep_inds (B, rois_per_image) int32 [-1 padded], num_fg (B), num_keep (B))."""
    ov = _need_cuda(max_overlaps, "max_overlaps")
    k = _need_cuda(keys, "keys")
    nb = _need_cuda(num_boxes, "num_boxes", torch.int32)
    if ov.dim() != 2 or k.shape != ov.shape or nb.shape != (ov.size(0),):
        raise ValueError("max_overlaps / keys must be (B,N) and num_boxes (B)")
    B, N = int(ov.size(0)), int(ov.size(1))
    keep = torch.empty((B, int(rois_per_image)), dtype=torch.int32, device=ov.device)
    nfg = torch.empty((B,), dtype=torch.int32, device=ov.device)
    nkeep = torch.empty((B,), dtype=torch.int32, device=ov.device)
    with _on(ov):
        _lib.call("vosd_sample_rois", _ptr(ov), _ptr(k), _ptr(nb), B, N, int(rois_per_image), int(fg_rois_per_image),
                  float(fg_thresh), float(bg_thresh_hi), float(bg_thresh_lo), _ptr(keep), _ptr(nfg), _ptr(nkeep), _stream())
        return keep, nfg, nkeep
