"""GPU parity: box-head post-processing through the C ABI --
vosd_bbox_transform (lib/core/test.py:178-179) and vosd_box_results (box_results_with_nms_and_limit,
lib/core/test.py:733-797) -- against the golden outputs of the reference itself and against the oracle.
Bars: kept detections, their order, per-class counts and scores bit-exact (tie-free inputs); decoded boxes
rtol 1e-5 (bit-identical in practice)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_bbox_transform_golden(golden):
    from vosdetectron_b200 import ops
    from vosdetectron_b200.utils import boxes as box_utils
    g = golden("box_results")
    pred = ops.bbox_transform_cuda(cu(g["props"]), cu(g["deltas"]), tuple(g["weights"]), clip_hw=(192, 256))
    assert np.allclose(pred.cpu().numpy(), g["pred_boxes"], rtol=1e-5, atol=0)
    assert np.array_equal(pred.cpu().numpy(), g["pred_boxes"])
    b = golden("boxes")
    out = box_utils.bbox_transform(b["boxes"], b["deltas"], (10., 10., 5., 5.))        # reference ndarray signature
    assert out.dtype == np.float32 and np.allclose(out, b["transformed"], rtol=1e-5, atol=0)
    assert box_utils.bbox_transform(np.zeros((0, 4), np.float32), np.zeros((0, 8), np.float32)).shape == (0, 8)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_box_results_golden(golden, tag):
    from vosdetectron_b200 import ops
    g = golden("box_results")
    K = int(g["num_classes"])
    R = g["scores"].shape[0]
    dets, count, cls_count = ops.box_results_cuda(cu(g["scores"])[None], cu(g["pred_boxes"])[None],
                                                  float(g["score_thresh"]), float(g["nms_" + tag]),
                                                  int(g["per_im_" + tag]), cap=R * (K - 1))
    n = int(count[0])
    d = dets[0, :n].cpu().numpy()
    assert n == len(g["out_scores_" + tag])
    assert np.array_equal(d[:, 4], g["out_scores_" + tag])
    assert np.array_equal(d[:, :4], g["out_boxes_" + tag])
    assert np.array_equal(cls_count[0].cpu().numpy(), g["cls_count_" + tag])
    assert np.array_equal(d[:, 5].astype(np.int32), np.repeat(np.arange(K), g["cls_count_" + tag]))


def test_reference_signature_full_size(orc, synth):
    """R = 1000 proposals x 81 classes (BASELINE config 3 sizes) through the drop-in with the reference's
    ndarray signature, against the oracle."""
    from vosdetectron_b200.core.test import box_results_with_nms_and_limit
    from vosdetectron_b200.config import clone_cfg
    props, scores, deltas = synth.box_head_outputs(31, 1000, 81)
    pred = orc.box_decode(props, deltas, (10., 10., 5., 5.), np.array(synth.COCO_BLOB, dtype=np.float32))
    for nms_t, per_im in ((0.5, 100), (0.3, 100), (0.5, 0)):
        cfg = clone_cfg()
        cfg.test_nms, cfg.test_detections_per_im = nms_t, per_im
        s, b, cls_boxes = box_results_with_nms_and_limit(scores, pred, cfg)
        so, bo, co = orc.box_results_with_nms_and_limit(scores, pred, 81, 0.05, nms_t, per_im)
        assert s.dtype == np.float32 and np.array_equal(s, so) and np.array_equal(b, bo)
        assert all(np.array_equal(x, y) for x, y in zip(cls_boxes[1:], co[1:]))
    cfg = clone_cfg()
    cfg.test_num_det_per_class = 3
    s, b, cls_boxes = box_results_with_nms_and_limit(scores, pred, cfg)
    so, bo, co = orc.box_results_with_nms_and_limit(scores, pred, 81, 0.05, 0.3, 100, 3)
    assert np.array_equal(s, so) and np.array_equal(b, bo)
    cfg.test_soft_nms = True
    with pytest.raises(NotImplementedError):
        box_results_with_nms_and_limit(scores, pred, cfg)


def test_batch_rows_ties_and_capacity(orc, synth):
    from vosdetectron_b200 import ops
    K, R = 6, 300
    imgs = [synth.box_head_outputs(100 + i, R, K, (192, 256)) for i in range(3)]
    pred = [orc.box_decode(p, d, (10., 10., 5., 5.), np.array([192, 256], np.float32)) for p, s, d in imgs]
    scores = np.stack([s for p, s, d in imgs])
    # image 1: ties at the image threshold -- 10 distinct scores above 30 identical ones, all on disjoint boxes:
    # the 20th largest score is the tied value, so all 40 detections must be kept (test.py:780-783 uses >=)
    scores[1] *= 0.01
    grid = np.array([[2 + 31 * (i % 8), 2 + 37 * (i // 8), 28 + 31 * (i % 8), 34 + 37 * (i // 8)] for i in range(40)],
                    np.float32)
    scores[1, :10, 1] = 0.9 + 0.005 * np.arange(10, dtype=np.float32)
    scores[1, 10:40, 2] = 0.5
    pred[1][:10, 4:8] = grid[:10]
    pred[1][10:40, 8:12] = grid[10:]
    rows = np.array([R, R, 120], dtype=np.int32)                       # image 2 has only 120 valid proposals
    dets, count, cls_count = ops.box_results_cuda(cu(scores), cu(np.stack(pred)), 0.05, 0.5, 20, rows=cu(rows),
                                                  cap=R * (K - 1))
    for i in range(3):
        so, bo, co = orc.box_results_with_nms_and_limit(scores[i, :rows[i]], pred[i][:rows[i]], K, 0.05, 0.5, 20)
        n = int(count[i])
        assert n == len(so), i
        d = dets[i, :n].cpu().numpy()
        assert np.array_equal(d[:, 4], so) and np.array_equal(d[:, :4], bo), i
        assert cls_count[i].cpu().numpy().tolist() == [len(c) for c in co], i
    assert int(count[1]) == 40                                         # the tie group pushed it over the limit
    # a capacity smaller than the result: the true count is still reported, the first `cap` rows are exact
    d2, c2, _ = ops.box_results_cuda(cu(scores), cu(np.stack(pred)), 0.05, 0.5, 20, rows=cu(rows), cap=7)
    assert torch.equal(c2, count) and torch.equal(d2[:, :7], dets[:, :7])
    # nothing above the threshold
    z = torch.zeros((1, 50, K), device="cuda")
    d3, c3, cc3 = ops.box_results_cuda(z, torch.zeros((1, 50, 4 * K), device="cuda"), 0.05, 0.5, 20)
    assert int(c3[0]) == 0 and int(cc3.sum()) == 0


def test_bbox_overlaps_bit_identical_to_cython_bbox(synth):
    """vosd_bbox_overlaps against the reference's own cython_bbox (oracle/_ref, compiled from
    lib/utils/cython_bbox.pyx): bit-identical matrix, np.argmax / np.max of it for the labels."""
    import importlib.util
    import os
    from vosdetectron_b200 import ops
    from vosdetectron_b200.utils import boxes as box_utils
    ref_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")
    so = [f for f in os.listdir(ref_dir) if f.startswith("cython_bbox")] if os.path.isdir(ref_dir) else []
    if not so:
        pytest.skip("oracle/_ref/cython_bbox*.so not built")
    spec = importlib.util.spec_from_file_location("cython_bbox", os.path.join(ref_dir, so[0]))
    cb = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cb)
    rs = np.random.RandomState(12)
    for N, K in [(1, 1), (2000, 17), (513, 1500), (300, 0), (0, 5)]:
        boxes = synth.random_rois(100 + N, max(N, 1), (800, 1344))[:N, 1:5].copy()
        gts = synth.random_rois(200 + K, max(K, 1), (800, 1344), smin=30)[:K, 1:5].copy()
        if N > 10 and K > 3:
            boxes[:5] = gts[:1]                         # exact matches (overlap 1.0) and ties across duplicated gts
            gts[2] = gts[1]
            boxes[7, 2] = boxes[7, 0] - 3.0             # negative width
        want = cb.bbox_overlaps(boxes, gts)
        got = box_utils.bbox_overlaps(boxes, gts)
        assert got.dtype == np.float32 and got.shape == (N, K) and np.array_equal(got, want)
        if N and K:
            _, mx, am = ops.bbox_overlaps_cuda(torch.from_numpy(boxes).cuda(), torch.from_numpy(gts).cuda(), want_matrix=False)
            assert np.array_equal(mx.cpu().numpy(), want.max(axis=1))
            assert np.array_equal(am.cpu().numpy(), want.argmax(axis=1))
    with pytest.raises(ValueError):
        box_utils.bbox_overlaps(np.zeros((1, 4)), np.zeros((1, 4), np.float32))


def test_bbox_targets_against_reference_outputs(golden):
    """vosd_bbox_targets against lib/roi_data/fast_rcnn.py's own outputs: zeros / weights / class columns exact,
    dx, dy bit-identical (fp32 NumPy order), dw, dh within 2 ulp-ish (logf vs NumPy's float32 log): rtol 1e-6 + 1e-7."""
    from vosdetectron_b200 import ops
    g = golden("bbox_targets")
    ex, gt = torch.from_numpy(g["ex"]).cuda(), torch.from_numpy(g["gt"]).cuda()
    lb = torch.from_numpy(g["labels"]).cuda()
    for tag, agn in (("k", False), ("a", True)):
        t, w, o = (x.cpu().numpy() for x in ops.bbox_targets_cuda(ex, gt, lb, int(g["num_classes"]), class_agnostic=agn))
        rt = g["targets_" + tag]
        assert t.shape == rt.shape and np.array_equal(w, g["inside_" + tag]) and np.array_equal(o, g["outside_" + tag])
        assert np.array_equal(t == 0, rt == 0)
        hit = g["inside_" + tag] > 0
        cols = np.arange(rt.shape[1])[None, :].repeat(rt.shape[0], 0) % 4
        xy = hit & (cols < 2)
        assert np.array_equal(t[xy], rt[xy])
        assert np.allclose(t[hit], rt[hit], rtol=1e-6, atol=1e-7)
    e = ops.bbox_targets_cuda(ex[:0], gt[:0], lb[:0], 9)
    assert e[0].shape == (0, 36)
