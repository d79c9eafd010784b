"""CUDA-backed subset of lib/utils/boxes.py with the reference's ndarray signatures.

``bbox_transform(boxes, deltas, weights)`` replaces boxes.py:156-205 (one thread per (row, class)).

``nms(dets, thresh)`` replaces boxes.py:329-333 -> cython_nms.pyx:37-87: ndarray (n,5) float32
in, int64 ndarray of kept indices (ascending) out, ``[]`` for empty input.  One H2D copy in,
one D2H copy out; the work is done by the bitmask kernels of csrc/proposals.cu.
``nms_cuda`` is the tensor-in / tensor-out variant that stays on the device.

``bbox_overlaps(boxes, query_boxes)`` replaces cython_bbox.bbox_overlaps (boxes.py:55); ``ops.bbox_overlaps_cuda``
also returns the row maxima / arg-maxima the label assignment takes next (json_dataset.py:453-456).
"""
import numpy as np
import torch

from .. import ops

nms_cuda = ops.nms_cuda


def nms(dets, thresh):
    if dets.shape[0] == 0:
        return []
    d = torch.from_numpy(np.ascontiguousarray(dets, dtype=np.float32)).cuda()
    keep, num = ops.nms_cuda(d, thresh)
    n = int(num.item())
    return keep[:n].cpu().numpy()


def bbox_overlaps(boxes, query_boxes):
    """cython_bbox.bbox_overlaps (cython_bbox.pyx:32-73; boxes.py:55) with the reference's ndarray signature:
    (N,4) x (K,4) float32 -> (N,K) float32."""
    if boxes.dtype != np.float32 or query_boxes.dtype != np.float32:
        raise ValueError("Buffer dtype mismatch, expected 'float32'")        # what the typed Cython signature raises
    N, K = boxes.shape[0], query_boxes.shape[0]
    if N == 0 or K == 0:
        return np.zeros((N, K), dtype=np.float32)
    b = torch.from_numpy(np.ascontiguousarray(boxes)).cuda()
    q = torch.from_numpy(np.ascontiguousarray(query_boxes)).cuda()
    return ops.bbox_overlaps_cuda(b, q)[0].cpu().numpy()


def bbox_transform(boxes, deltas, weights=(1.0, 1.0, 1.0, 1.0)):
    """boxes.py:156-205 with the reference's ndarray signature: (n,4) x (n,4k) -> (n,4k) float32."""
    if boxes.shape[0] == 0:
        return np.zeros((0, deltas.shape[1]), dtype=deltas.dtype)
    b = torch.from_numpy(np.ascontiguousarray(boxes, dtype=np.float32)).cuda()
    d = torch.from_numpy(np.ascontiguousarray(deltas, dtype=np.float32)).cuda()
    return ops.bbox_transform_cuda(b, d, weights).cpu().numpy()


def expand_boxes(boxes, scale):
    """boxes.py:242-258 (host helper kept for callers that need the expanded reference boxes;
    the paste kernel applies the same arithmetic on the device)."""
    w_half = (boxes[:, 2] - boxes[:, 0]) * .5
    h_half = (boxes[:, 3] - boxes[:, 1]) * .5
    x_c = (boxes[:, 2] + boxes[:, 0]) * .5
    y_c = (boxes[:, 3] + boxes[:, 1]) * .5
    w_half *= scale
    h_half *= scale
    out = np.zeros(boxes.shape)
    out[:, 0], out[:, 2] = x_c - w_half, x_c + w_half
    out[:, 1], out[:, 3] = y_c - h_half, y_c + h_half
    return out
