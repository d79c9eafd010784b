"""CPU: the oracle restatement against the golden vectors produced by the reference itself
(tests/golden/make_golden.py) and against the third-party functions it restates."""
import numpy as np
import pytest


def test_anchor_known_answer(orc, golden):
    # the 9 stride-16 anchors quoted in lib/modeling/generate_anchors.py:26-51
    kat = np.array([[-84, -40, 99, 55], [-176, -88, 191, 103], [-360, -184, 375, 199],
                    [-56, -56, 71, 71], [-120, -120, 135, 135], [-248, -248, 263, 263],
                    [-36, -80, 51, 95], [-80, -168, 95, 183], [-168, -344, 183, 359]], dtype=np.float64)
    a = orc.generate_anchors(16, (128, 256, 512), (0.5, 1, 2))
    assert a.dtype == np.float64 and np.array_equal(a, kat)
    g = golden("anchors")
    assert np.array_equal(g["kat_stride16"], kat)
    for lvl in range(2, 7):
        assert np.array_equal(orc.fpn_anchors(lvl), g["fpn%d" % lvl])


def test_generate_proposals_golden(orc, golden):
    g = golden("proposals")
    for lvl in range(2, 7):
        rois, probs = orc.generate_proposals(g["scores%d" % lvl], g["deltas%d" % lvl], g["im_info"],
                                             orc.fpn_anchors(lvl), 1. / 2 ** lvl, int(g["pre"]), int(g["post"]),
                                             float(g["thresh"]), float(g["min_size"]))
        assert rois.dtype == np.float32 and probs.dtype == np.float32
        assert np.array_equal(rois, g["rois%d" % lvl]), lvl
        assert np.array_equal(probs, g["probs%d" % lvl]), lvl
    rois, probs = orc.generate_proposals(g["scores3"], g["deltas3"], g["im_info"], orc.fpn_anchors(3), 1. / 8,
                                         int(g["pre"]), int(g["post"]), float(g["thresh"]), 16)
    assert np.array_equal(rois, g["rois3_min16"]) and np.array_equal(probs, g["probs3_min16"])
    assert not np.array_equal(rois, g["rois3"])      # the filter really removed boxes


def test_nms_golden(orc, golden):
    g = golden("nms")
    for t in (0.3, 0.5, 0.7):
        for kind in ("unsorted", "sorted"):
            keep = orc.nms(g["dets_" + kind], t)
            ref = g["keep_%s_%02d" % (kind, int(t * 10))]
            assert keep.dtype == np.int64 and np.array_equal(keep, ref), (t, kind)
    assert orc.nms(np.zeros((0, 5), np.float32), 0.5) == []


def test_collect_distribute_golden(orc, golden):
    g = golden("collect_distribute")
    rois = orc.collect([g["in_rois%d" % l] for l in range(2, 7)], [g["in_probs%d" % l] for l in range(2, 7)],
                       int(g["post"]))
    assert np.array_equal(rois, g["rois"])
    assert np.array_equal(orc.map_rois_to_fpn_levels(rois[:, 1:5]), g["levels"])
    d = orc.distribute(rois)
    for k in ("rois_fpn2", "rois_fpn3", "rois_fpn4", "rois_fpn5", "rois_idx_restore_int32"):
        assert d[k].dtype == g[k].dtype and np.array_equal(d[k], g[k]), k
    wide = g["wide_rois"]
    assert np.array_equal(orc.map_rois_to_fpn_levels(wide[:, 1:5]), g["wide_levels"])
    d = orc.distribute(wide)
    for k in ("rois_fpn2", "rois_fpn3", "rois_fpn4", "rois_fpn5", "rois_idx_restore_int32"):
        assert np.array_equal(d[k], g["wide_" + k]), k
    assert {int(v) for v in g["wide_levels"]} == {2, 3, 4, 5}


def test_box_helpers_golden(orc, golden):
    g = golden("boxes")
    xf = orc.bbox_transform(g["boxes"], g["deltas"], (10., 10., 5., 5.))
    assert np.array_equal(xf, g["transformed"])
    cl = orc.clip_tiled_boxes(xf.copy(), np.array([192, 256], dtype=np.float32))
    assert np.array_equal(cl, g["clipped"])
    ex = orc.expand_boxes(g["boxes"], 30.0 / 28.0)
    assert ex.dtype == np.float64 and np.array_equal(ex, g["expanded"])


def test_paste_golden(orc, golden):
    g = golden("paste")
    fh, fw = (int(v) for v in g["frame_hw"])
    ref = np.unpackbits(g["packed"], axis=-1)[..., :fw]
    out, prob = orc.paste_masks(g["masks"], g["cls"], g["boxes"], fh, fw, want_prob=True)
    # the C restatement of cv2.resize is within 4e-6 of cv2, so a pixel may only flip when the
    # probability sits on the 0.5 threshold
    diff = out != ref
    assert np.all(np.abs(prob[diff] - 0.5) < 1e-5)
    assert diff.sum() <= 2
    assert ref.sum() > 1000


def test_resize_restatement_vs_cv2(orc):
    cv2 = pytest.importorskip("cv2")
    rs = np.random.RandomState(3)
    worst = 0.0
    for _ in range(150):
        M = rs.choice([14, 28, 56])
        src = np.zeros((M + 2, M + 2), np.float32)
        src[1:-1, 1:-1] = rs.uniform(size=(M, M)).astype(np.float32)
        w, h = int(rs.randint(1, 300)), int(rs.randint(1, 300))
        a = orc.resize_linear(src, w, h)
        b = cv2.resize(src, (w, h))
        worst = max(worst, float(np.abs(a - b).max()))
    a = orc.resize_linear(src, (M + 2) // 2, (M + 2) // 2)        # the INTER_AREA special case
    worst = max(worst, float(np.abs(a - cv2.resize(src, ((M + 2) // 2, (M + 2) // 2))).max()))
    assert worst <= 4e-6, worst


def test_roialign_forward_vs_torchvision_golden(orc, golden):
    g = golden("roialign_tv")
    for lvl in (2, 3, 4, 5):
        for res in (7, 14):
            out = orc.roi_align_forward(g["feat%d" % lvl], g["rois"], res, res, 1. / 2 ** lvl, 2)
            ref = g["tv_fwd_l%d_r%d" % (lvl, res)]
            # two fp32 evaluations of the same formula that round the sample coordinates
            # differently: loose anchor only
            assert np.abs(out - ref).max() < 1e-4, (lvl, res, np.abs(out - ref).max())


def test_roialign_backward_is_adjoint_of_forward(orc, synth):
    rs = np.random.RandomState(9)
    feats = rs.standard_normal((2, 3, 20, 24)).astype(np.float32)
    rois = synth.random_rois(10, 12, (160, 192), num_images=2, smin=8, smax=150)
    rois = np.concatenate([rois, synth.edge_rois((160, 192))])
    for sr in (2, 0):
        out = orc.roi_align_forward(feats, rois, 7, 7, 0.125, sr)
        gout = rs.standard_normal(out.shape).astype(np.float32)
        gin = orc.roi_align_backward(gout, rois, feats.shape, 7, 7, 0.125, sr)
        lhs = float((out.astype(np.float64) * gout).sum())
        rhs = float((feats.astype(np.float64) * gin).sum())
        assert abs(lhs - rhs) <= 1e-4 * max(1.0, abs(lhs)), (sr, lhs, rhs)


def test_roialign_oracle_threads_agree(orc, synth):
    feats = np.random.RandomState(1).standard_normal((1, 4, 30, 40)).astype(np.float32)
    rois = synth.random_rois(2, 20, (240, 320))
    a = orc.roi_align_forward(feats, rois, 7, 7, 0.125, 2, nthreads=1)
    b = orc.roi_align_forward(feats, rois, 7, 7, 0.125, 2, nthreads=4)
    assert np.array_equal(a, b)


def test_box_results_golden(orc, golden):
    """Box-head decode + box_results_with_nms_and_limit (lib/core/test.py:178-179, 733-797) against the
    reference's own outputs."""
    g = golden("box_results")
    K = int(g["num_classes"])
    pred = orc.box_decode(g["props"], g["deltas"], tuple(g["weights"]), np.array([192, 256], dtype=np.float32))
    assert np.array_equal(pred, g["pred_boxes"])
    for tag in "abc":
        s, b, cls_boxes = orc.box_results_with_nms_and_limit(g["scores"], g["pred_boxes"], K, float(g["score_thresh"]),
                                                            float(g["nms_" + tag]), int(g["per_im_" + tag]))
        assert np.array_equal(s, g["out_scores_" + tag]) and np.array_equal(b, g["out_boxes_" + tag]), tag
        assert [len(c) for c in cls_boxes[1:]] == g["cls_count_" + tag][1:].tolist()


def test_rle_restatement_round_trips_and_hand_vectors(orc):
    """maskApi.c restatement: rleFrString . rleToString == id, rleDecode . rleEncode == id, the loop and the
    vectorised run extraction agree; vectors worked by hand from rleToString (6 -> '6'; 20 = 0b10100 has bit 4
    set in its last group, so a continuation group follows: 'd0')."""
    rs = np.random.RandomState(5)
    for _ in range(100):
        h, w = rs.randint(1, 24), rs.randint(1, 24)
        m = (rs.rand(h, w) > rs.rand()).astype(np.uint8)
        c = orc.rle_counts(m)
        assert c == orc.rle_counts_fast(m) and sum(c) == h * w
        assert orc.rle_from_string(orc.rle_to_string(c)) == c
        assert np.array_equal(orc.rle_decode(c, h, w), m)
    assert orc.rle_counts(np.zeros((3, 4), np.uint8)) == [12]
    assert orc.rle_counts(np.ones((3, 4), np.uint8)) == [0, 12]
    assert orc.rle_counts(np.array([[0, 1], [1, 1]], np.uint8)) == [1, 3]          # column-major: 0,1 | 1,1
    assert orc.rle_to_string([6]) == '6' and orc.rle_to_string([20]) == 'd0'
    assert orc.rle_from_string('d0') == [20]
    # negative differences (run i smaller than run i-2) survive the sign extension
    c = [5, 1000, 3, 2, 70000, 1]
    assert orc.rle_from_string(orc.rle_to_string(c)) == c


def test_flow_align_oracle_vs_grid_sample_anchor(orc, synth):
    """FlowAlign restatement (flow_align_cuda_kernel.cu:15-117) against a loose CPU anchor: inside the valid
    range it is torch's bilinear grid_sample (align_corners=True), forward and both gradients; outside
    [0,H-1) x [0,W-1) the reference writes 0 and propagates nothing.  The pin proper is the reference kernel on
    the GPU (tests/test_gpu_flowalign.py)."""
    import torch
    N, C, H, W = 2, 5, 14, 19
    rs = np.random.RandomState(3)
    f = rs.standard_normal((N, C, H, W)).astype(np.float32)
    fl = synth.flow_field(4, N, H, W, "smooth", 2.5)
    out = orc.flow_align_forward(f, fl)
    ys, xs = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    gx, gy = xs + fl[:, 0].astype(np.float64), ys + fl[:, 1].astype(np.float64)
    inb = (gx >= 0) & (gx < W - 1) & (gy >= 0) & (gy < H - 1)
    assert 0.3 < inb.mean() < 1.0
    tf = torch.from_numpy(f).double().requires_grad_(True)
    tfl = torch.from_numpy(fl).double().requires_grad_(True)
    grid = torch.stack([(torch.from_numpy(xs) + tfl[:, 0]) / (W - 1) * 2 - 1,
                        (torch.from_numpy(ys) + tfl[:, 1]) / (H - 1) * 2 - 1], -1)
    ref = torch.nn.functional.grid_sample(tf, grid, mode="bilinear", padding_mode="zeros", align_corners=True)
    m = torch.from_numpy(inb)[:, None].expand(N, C, H, W)
    assert np.all(out[~m.numpy()] == 0)
    assert np.abs(out - ref.detach().numpy())[m.numpy()].max() < 1e-5
    g = rs.standard_normal((N, C, H, W)).astype(np.float32)
    (ref * torch.from_numpy(g).double() * m).sum().backward()
    gf, gfl = orc.flow_align_backward(g, f, fl)
    assert np.abs(gf - tf.grad.numpy()).max() < 1e-4
    assert np.abs(gfl - tfl.grad.numpy()).max() < 1e-4 * max(1.0, float(np.abs(gfl).max()))
    # thread count does not change a bit
    assert np.array_equal(out, orc.flow_align_forward(f, fl, nthreads=1))
    # identity and integer shift are exact
    z = orc.flow_align_forward(f, synth.flow_field(0, N, H, W, "zero"))
    assert np.array_equal(z[:, :, :H - 1, :W - 1], f[:, :, :H - 1, :W - 1]) and not z[:, :, H - 1].any()
    sh = orc.flow_align_forward(f, synth.flow_field(0, N, H, W, "shift"))
    assert np.array_equal(sh[:, :, 1:H, :W - 3], f[:, :, 0:H - 1, 2:W - 1])


def _vos_prev(g, K):
    prev, start = [[] for _ in range(K)], 0
    for j in range(1, K):
        n = int(g["prev_count"][j])
        prev[j] = g["prev_boxes"][start:start + n]
        start += n
    return prev


def test_vos_box_results_golden(orc, golden):
    """lib_vos box_results_with_nms_and_limit (vos_test.py:748-865: cross-class NMS, per-class top-k, previous-box
    filter) against the reference's own outputs."""
    g, b = golden("vos_post"), golden("box_results")
    K = int(b["num_classes"])
    prev = _vos_prev(g, K)
    for tag in "xyzw":
        cross, pre, small, small_th = g["set_" + tag]
        s, bx, cls_boxes = orc.vos_box_results(b["scores"], b["pred_boxes"], K, float(b["score_thresh"]), 0.5, 100,
                                               float(cross), int(pre), float(small), float(small_th), prev)
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
    return Km, cls_boxes, cls_segms


def test_nms_with_mask_iou_golden(orc, golden):
    """vos_test.py:985-1029 against the reference's own outputs (its mask codec bound to the maskApi.c restatement);
    the RLE strings the reference produced from ITS segm_results are reproduced by the oracle's paste + encode."""
    g = golden("vos_post")
    Km, cls_boxes, cls_segms = _mask_case(g)
    h, w = (int(v) for v in g["m_frame"])
    pasted = orc.paste_masks(g["m_masks"], g["m_cls"], g["m_boxes"][:, :4], h, w)
    assert [orc.rle_encode(m)['counts'] for m in pasted] == [str(c) for c in g["m_counts"]]
    for tag in "pqr":
        th, per = g["mset_" + tag]
        ob, os_ = orc.nms_with_mask_iou(cls_boxes, cls_segms, 6, float(th), int(per))
        assert [len(c) for c in ob] == g["mout_count_" + tag].tolist(), tag
        got = np.vstack([np.vstack(c) for c in ob if len(c)])
        assert np.array_equal(got, g["mout_boxes_" + tag])
        assert [s['counts'] for sl in os_ for s in sl] == [str(c) for c in g["mout_counts_" + tag]]


def test_bbox_targets_golden(orc, golden):
    """_compute_targets + _expand_bbox_targets (lib/roi_data/fast_rcnn.py:216-260) against the reference's outputs."""
    g = golden("bbox_targets")
    for tag, agn in (("k", False), ("a", True)):
        t, w, o = orc.bbox_targets(g["ex"], g["gt"], g["labels"], int(g["num_classes"]), class_agnostic=agn)
        assert np.array_equal(t, g["targets_" + tag]) and np.array_equal(w, g["inside_" + tag])
        assert np.array_equal(o, g["outside_" + tag])


def test_label_assignment_oracle_reproduces_the_reference(orc, golden):
    """add_proposals + _sample_rois + add_fast_rcnn_blobs (+ mask_rois) restated in oracle/region_oracle.py against
    tests/golden/labels.npz, produced by the unmodified reference under the key-based RNG contract."""
    g = golden("labels")
    K, scales = int(g["num_classes"]), g["im_scales"]
    samples, mrois = [], []
    for i in range(2):
        boxes, mo, mc, b2g = orc.add_proposals(g["gt_boxes%d" % i], g["gt_classes%d" % i], g["rpn_rois"], scales[i], i)
        assert np.array_equal(boxes, g["boxes%d" % i]) and np.array_equal(mo, g["max_overlaps%d" % i])
        assert np.array_equal(mc, g["max_classes%d" % i]) and np.array_equal(b2g, g["box_to_gt%d" % i])
        s = orc.sample_rois(boxes, mo, mc, b2g, g["gt_boxes%d" % i], g["keys%d" % i], scales[i], i, K)
        samples.append(s)
        mrois.append(orc.mask_rois_of(s, scales[i], i))
    blobs = orc.add_fast_rcnn_blobs(samples, mask_rois=mrois)
    for k in ("labels_int32", "rois", "bbox_targets", "bbox_inside_weights", "bbox_outside_weights", "rois_fpn2", "rois_fpn3",
              "rois_fpn4", "rois_fpn5", "rois_idx_restore_int32"):
        assert np.array_equal(blobs[k], g[k]), k
        assert np.array_equal(blobs[k], g["m_" + k]), k
    for k in ("mask_rois", "roi_has_mask_int32", "mask_rois_fpn2", "mask_rois_fpn3", "mask_rois_fpn4", "mask_rois_fpn5",
              "mask_rois_idx_restore_int32"):
        assert np.array_equal(blobs[k], g["m_" + k]), k


def _rpn_case(orc, g, tag, fields_spec):
    """Re-run one case of tests/golden/rpn_labels.npz through the restatement -> dict of concatenated blobs."""
    parts = [orc.field_of_anchors(st, sz, (0.5, 1, 2), 384) for st, sz in fields_spec]
    anchors = np.concatenate([p[0] for p in parts])
    fields = [(p[1], p[2]) for p in parts]
    blobs = {}
    for i in range(int(g[tag + "n"])):
        s = g[tag + "im_scales"][i]
        h, w = g["%shw%d" % (tag, i)]
        keep = np.where((g["%sgt_classes%d" % (tag, i)] > 0) & (g["%sis_crowd%d" % (tag, i)] == 0))[0]
        gt = g["%sboxes%d" % (tag, i)][keep] * s
        lab = orc.rpn_labels(np.round(h * s), np.round(w * s), anchors, gt, g["%skeys%d" % (tag, i)], g["%subg%d" % (tag, i)])
        for j, d in enumerate(orc.rpn_blobs_split(fields, *lab)):
            for k, v in d.items():
                blobs.setdefault(k + ("_fpn%d" % (j + 2) if len(fields) > 1 else ""), []).append(v)
    return {k: np.concatenate(v) for k, v in blobs.items()}


FPN_FIELDS = [(2. ** l, (32 * 2. ** (l - 2),)) for l in range(2, 7)]


@pytest.mark.parametrize("tag,spec", [("fpn_", FPN_FIELDS), ("nogt_", FPN_FIELDS), ("quirk_", FPN_FIELDS),
                                      ("single_", [(16, (32, 64, 128, 256))])])
def test_rpn_label_oracle_reproduces_the_reference(orc, golden, tag, spec):
    """_get_rpn_blobs restated in oracle/region_oracle.py against tests/golden/rpn_labels.npz (unmodified reference under
    the RNG contract of vosdetectron_b200/roi_data/rpn.py): every blob bit-exact."""
    g = golden("rpn_labels")
    blobs = _rpn_case(orc, g, tag, spec)
    assert blobs
    for k, v in blobs.items():
        ref = g[tag + k]
        assert v.dtype == ref.dtype and v.shape == ref.shape, k
        assert np.array_equal(v, ref), k
