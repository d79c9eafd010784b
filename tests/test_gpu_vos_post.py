"""GPU parity: lib_vos detection post-processing (SURVEY 8f rank 1 extras + rank 2 second half) through the C ABI:
  * vos_test.box_results_with_nms_and_limit mirror against the reference's own outputs (tests/golden/vos_post.npz);
  * vosd_rle_to_bits against the oracle's rleDecode; vosd_mask_iou_nms / nms_with_mask_iou mirror against the
    reference's outputs and against the oracle on random masks.  All bit-exact (keep decisions, orders, strings)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cfg(K, **kw):
    from vosdetectron_b200.config import RegionConfig
    return RegionConfig(num_classes=K, **kw)


def _vos_prev(g, K):
    prev, start = [[] for _ in range(K)], 0
    for j in range(1, K):
        n = int(g["prev_count"][j])
        prev[j] = g["prev_boxes"][start:start + n]
        start += n
    return prev


def test_vos_box_results_against_reference_outputs(golden):
    from vosdetectron_b200.core import vos_test
    g, b = golden("vos_post"), golden("box_results")
    K = int(b["num_classes"])
    prev = _vos_prev(g, K)
    for tag in "xyzw":
        cross, pre, small, small_th = g["set_" + tag]
        cfg = _cfg(K, test_score_thresh=float(b["score_thresh"]), test_nms=0.5, test_detections_per_im=100,
                   test_nms_cross_class=float(cross), test_num_det_per_class_pre=int(pre),
                   test_nms_small_box_iou=float(small), test_nms_small_box_score_threshold=float(small_th))
        s, bx, cls_boxes = vos_test.box_results_with_nms_and_limit(b["scores"], b["pred_boxes"], prev_cls_boxes=prev, cfg=cfg)
        assert np.array_equal(s, g["out_scores_" + tag]) and np.array_equal(bx, g["out_boxes_" + tag]), tag
        assert [len(c) for c in cls_boxes[1:]] == g["cls_count_" + tag][1:].tolist()


def _mask_case(g):
    Km = int(g["m_cls"].max()) + 1
    cls_boxes = [[] for _ in range(Km)]
    cls_segms = [[] for _ in range(Km)]
    h, w = (int(v) for v in g["m_frame"])
    for i, c in enumerate(g["m_cls"]):
        cls_segms[int(c)].append({'size': [h, w], 'counts': str(g["m_counts"][i])})
    for j in range(1, Km):
        cls_boxes[j] = g["m_boxes"][g["m_cls"] == j]
    return Km, cls_boxes, cls_segms, h, w


def test_segm_results_then_nms_with_mask_iou_against_reference_outputs(golden):
    from vosdetectron_b200.core import vos_test
    g = golden("vos_post")
    Km, cls_boxes, cls_segms, h, w = _mask_case(g)
    cfg = _cfg(Km)
    # the RLE strings the reference's segm_results produced come out of the fused paste -> RLE kernel
    mine = vos_test.segm_results(cls_boxes, g["m_masks"], g["m_boxes"][:, :4], h, w, cfg=cfg)
    assert [s['counts'] for sl in mine for s in sl] == [str(c) for c in g["m_counts"]]
    for tag in "pqr":
        th, per = g["mset_" + tag]
        ob, os_ = vos_test.nms_with_mask_iou(cls_boxes, mine, iou_th=float(th), max_per_class=int(per), cfg=cfg)
        assert [len(c) for c in ob] == g["mout_count_" + tag].tolist(), tag
        assert np.array_equal(np.vstack([np.vstack(c) for c in ob if len(c)]), g["mout_boxes_" + tag])
        assert [s['counts'] for sl in os_ for s in sl] == [str(c) for c in g["mout_counts_" + tag]]


def _blobs(rs, n, h, w, groups):
    """n masks in `groups` clusters of near-identical blobs (so suppression really happens) + a few empty ones."""
    ys, xs = np.mgrid[0:h, 0:w]
    out = np.zeros((n, h, w), dtype=np.uint8)
    centres = rs.uniform(0.2, 0.8, (groups, 2)) * (h, w)
    radii = rs.uniform(0.08, 0.3, groups) * min(h, w)
    for i in range(n):
        k = rs.randint(groups)
        cy, cx = centres[k] + rs.uniform(-2, 2, 2)
        r = radii[k] * rs.uniform(0.85, 1.15)
        out[i] = ((ys - cy) ** 2 + (xs - cx) ** 2 <= r * r)
    out[rs.randint(n, size=max(1, n // 15))] = 0
    return out


@pytest.mark.parametrize("n,h,w", [(1, 8, 8), (2, 5, 7), (37, 60, 81), (100, 120, 168), (130, 33, 47)])
@pytest.mark.parametrize("th", [0.5, 0.9])
def test_mask_iou_nms_vs_oracle(orc, n, h, w, th):
    from vosdetectron_b200 import ops
    rs = np.random.RandomState(n * 7 + h)
    masks = _blobs(rs, n, h, w, groups=max(1, n // 6))
    scores = rs.permutation(n).astype(np.float32)
    order = np.argsort(-scores)
    want = orc.mask_iou_greedy([masks[k] for k in order], th)
    # (i) row-major bits straight from the dense masks (the layout vosd_paste_masks_packed writes)
    packed = ops.pack_mask_bits_cuda(torch.from_numpy(masks).cuda())
    removed, num = ops.mask_iou_nms_cuda(packed, torch.from_numpy(order.astype(np.int32)).cuda(), th)
    assert np.array_equal(removed.cpu().numpy(), want) and int(num.item()) == int((want == 0).sum())
    # (ii) column-major bits expanded from the RLE runs
    runs = [orc.rle_counts_fast(m) for m in masks]
    bits = ops.rle_to_bits_cuda(runs, h * w)
    ref_bits = np.stack([np.packbits(m.ravel(order='F'), bitorder='little') for m in masks])
    got = bits.cpu().numpy()[:, :ref_bits.shape[1]]
    assert np.array_equal(got, ref_bits) and not bits.cpu().numpy()[:, ref_bits.shape[1]:].any()
    removed2, _ = ops.mask_iou_nms_cuda(bits, torch.from_numpy(order.astype(np.int32)).cuda(), th)
    assert np.array_equal(removed2.cpu().numpy(), want)


def test_mask_iou_nms_device_pipeline_and_edges(orc, synth):
    """Paste (bit-packed) -> mask-IoU NMS without leaving the device, identity order, empty input."""
    from vosdetectron_b200 import ops
    from vosdetectron_b200.core import vos_test
    boxes, cls, masks = synth.detections(77, 40, (96, 128), 28, 5)
    boxes = np.concatenate([boxes, boxes[:10] + 0.5]).astype(np.float32)         # 10 near-duplicates
    masks = np.concatenate([masks, masks[:10]])
    cls = np.concatenate([cls, cls[:10]]).astype(np.int32)
    scores = np.random.RandomState(3).permutation(len(boxes)).astype(np.float32)
    dense, packed = ops.paste_masks_packed_cuda(torch.from_numpy(masks).cuda(), torch.from_numpy(cls).cuda(),
                                                torch.from_numpy(boxes).cuda(), 96, 128, 0.5)
    order, removed, num = vos_test.nms_with_mask_iou_cuda(torch.from_numpy(scores).cuda(), packed, 0.8)
    want = orc.mask_iou_greedy([dense[k].cpu().numpy() for k in order.cpu().numpy()], 0.8)
    assert np.array_equal(removed.cpu().numpy(), want) and 0 < int(num.item()) < len(boxes)
    removed_id, _ = ops.mask_iou_nms_cuda(packed, None, 0.8)
    assert np.array_equal(removed_id.cpu().numpy(), orc.mask_iou_greedy(list(dense.cpu().numpy()), 0.8))
    empty, num = ops.mask_iou_nms_cuda(torch.empty((0, 64), dtype=torch.uint8, device="cuda"), None, 0.5)
    assert empty.numel() == 0 and int(num.item()) == 0
    assert vos_test.nms_with_mask_iou([[], []], [[], []], cfg=_cfg(2)) == ([[], []], [[], []])
    with pytest.raises(ValueError):
        ops.rle_to_bits_cuda([[3, 2]], 6)
