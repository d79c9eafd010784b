"""GPU parity of the training label assignment (SURVEY 8f rank 3, row a11's training branch) against
tests/golden/labels.npz, produced by the UNMODIFIED reference (json_dataset.add_proposals, roi_data.fast_rcnn.
add_fast_rcnn_blobs / _sample_rois, roi_data.mask_rcnn.add_mask_rcnn_blobs) under the key-based RNG contract
(tests/golden/make_golden_labels.py).  Index / label / level blobs bit-exact; regression targets rtol 1e-6 (logf vs
NumPy's float32 log, <= 1 ulp), weights bit-exact."""
import numpy as np
import pytest
import scipy.sparse
import torch

pytestmark = pytest.mark.gpu


def _entries(g):
    K = int(g["num_classes"])
    out = []
    for i in range(2):
        gb, gc = g["gt_boxes%d" % i], g["gt_classes%d" % i]
        G = len(gc)
        ov = np.zeros((G, K), np.float32)
        ov[np.arange(G), gc] = 1.0
        out.append({"boxes": gb.copy(), "gt_classes": gc.copy(), "seg_areas": np.zeros(G, np.float32),
                    "gt_overlaps": scipy.sparse.csr_matrix(ov), "is_crowd": np.zeros(G, dtype=bool),
                    "box_to_gt_ind_map": np.arange(G, dtype=np.int32)})
    return out


def _check_blobs(blobs, g, tag):
    for k in ("labels_int32", "rois", "bbox_inside_weights", "bbox_outside_weights", "rois_fpn2", "rois_fpn3", "rois_fpn4",
              "rois_fpn5", "rois_idx_restore_int32"):
        assert blobs[k].dtype == g[tag + k].dtype and np.array_equal(blobs[k], g[tag + k]), k
    t, ref = blobs["bbox_targets"], g[tag + "bbox_targets"]
    assert t.shape == ref.shape and np.allclose(t, ref, rtol=1e-6, atol=1e-7)


def test_add_proposals_matches_the_reference(golden):
    from vosdetectron_b200.datasets import json_dataset
    g = golden("labels")
    roidb = _entries(g)
    json_dataset.add_proposals(roidb, g["rpn_rois"], g["im_scales"], crowd_thresh=0)
    for i, e in enumerate(roidb):
        assert np.array_equal(e["boxes"], g["boxes%d" % i])
        assert np.array_equal(e["max_overlaps"], g["max_overlaps%d" % i])
        assert np.array_equal(e["max_classes"], g["max_classes%d" % i])
        assert np.array_equal(e["box_to_gt_ind_map"], g["box_to_gt%d" % i])
        assert e["gt_classes"].shape[0] == e["boxes"].shape[0] and e["is_crowd"].shape[0] == e["boxes"].shape[0]


def test_sample_rois_kernel_against_the_oracle(orc, golden):
    """vosd_sample_rois alone: fg picks then bg picks in ascending key order, counts, -1 padding; several fg / bg mixes
    including an image without any foreground and one with fewer candidates than rois_per_image."""
    from vosdetectron_b200 import ops
    g = golden("labels")
    rs = np.random.RandomState(3)
    cases = [(g["max_overlaps0"], g["keys0"]), (g["max_overlaps1"], g["keys1"]),
             (np.clip(g["max_overlaps1"], 0, 0.4).astype(np.float32), g["keys1"]),                # no foreground
             (rs.uniform(0, 1, 100).astype(np.float32), rs.uniform(size=100).astype(np.float32)),   # fewer boxes than 512
             (np.full(30, 0.7, np.float32), rs.uniform(size=30).astype(np.float32))]                # only foreground
    N = max(len(c[0]) for c in cases)
    ov = np.full((len(cases), N), -1.0, np.float32)
    ky = np.ones((len(cases), N), np.float32)
    for i, (o, k) in enumerate(cases):
        ov[i, :len(o)], ky[i, :len(k)] = o, k
    nb = torch.tensor([len(c[0]) for c in cases], dtype=torch.int32).cuda()
    keep, nfg, nkeep = ops.sample_rois_cuda(torch.from_numpy(ov).cuda(), torch.from_numpy(ky).cuda(), nb, 512, 128, 0.5, 0.5, 0.0)
    keep, nfg, nkeep = keep.cpu().numpy(), nfg.cpu().numpy(), nkeep.cpu().numpy()
    for i, (o, k) in enumerate(cases):
        fg = np.where(o >= 0.5)[0]
        n1 = min(128, fg.size)
        fg = orc.choice_by_keys(fg, n1, k) if fg.size else fg
        bg = np.where((o < 0.5) & (o >= 0.0))[0]
        n2 = min(512 - n1, bg.size)
        bg = orc.choice_by_keys(bg, n2, k) if bg.size else bg
        want = np.append(fg, bg)
        assert nfg[i] == n1 and nkeep[i] == n1 + n2
        assert np.array_equal(keep[i, :n1 + n2], want) and (keep[i, n1 + n2:] == -1).all()


@pytest.mark.parametrize("mask_on", [False, True])
def test_add_fast_rcnn_blobs_matches_the_reference(golden, mask_on):
    from vosdetectron_b200.config import RegionConfig
    from vosdetectron_b200.datasets import json_dataset
    from vosdetectron_b200.roi_data import fast_rcnn
    g = golden("labels")
    cfg = RegionConfig(num_classes=int(g["num_classes"]), mask_on=mask_on)
    roidb = _entries(g)
    json_dataset.add_proposals(roidb, g["rpn_rois"], g["im_scales"], crowd_thresh=0)
    blobs = {k: [] for k in fast_rcnn.get_fast_rcnn_blob_names(True, cfg)}
    assert fast_rcnn.add_fast_rcnn_blobs(blobs, g["im_scales"], roidb, [g["keys0"], g["keys1"]], cfg)
    tag = "m_" if mask_on else ""
    _check_blobs(blobs, g, tag)
    if mask_on:
        for k in ("mask_rois", "roi_has_mask_int32", "mask_rois_fpn2", "mask_rois_fpn3", "mask_rois_fpn4", "mask_rois_fpn5",
                  "mask_rois_idx_restore_int32"):
            assert np.array_equal(blobs[k], g["m_" + k]), k
    want = set(fast_rcnn.get_fast_rcnn_blob_names(True, cfg)) - {"masks_int32"}     # pycocotools rasteriser: not here
    assert set(blobs) == want


def test_training_branch_of_the_collect_op(golden):
    """CollectAndDistributeFpnRpnProposalsOp.forward in training mode: collect (scores decreasing along the fixture's
    proposal rows, so the collected order is the fixture's) -> add_proposals -> add_fast_rcnn_blobs."""
    from vosdetectron_b200.config import RegionConfig
    from vosdetectron_b200.modeling.collect_and_distribute_fpn_rpn_proposals import CollectAndDistributeFpnRpnProposalsOp
    g = golden("labels")
    cfg = RegionConfig(num_classes=int(g["num_classes"]))
    rois = g["rpn_rois"]
    n = len(rois)
    probs = np.linspace(0.99, 0.01, n).astype(np.float32)
    cut = [0, 500, 900, 1300, 1600, n]                       # any split over the 5 RPN levels
    inputs = [rois[cut[i]:cut[i + 1]] for i in range(5)] + [probs[cut[i]:cut[i + 1]].reshape(-1, 1) for i in range(5)]
    op = CollectAndDistributeFpnRpnProposalsOp(cfg)
    op.train()
    im_info = torch.tensor([[800., 1333., float(g["im_scales"][0])], [800., 640., float(g["im_scales"][1])]])
    blobs = op(inputs, _entries(g), im_info, rand_keys=[g["keys0"], g["keys1"]])
    _check_blobs(blobs, g, "")
    op.eval()
    ev = op(inputs, None, im_info)
    assert np.array_equal(ev["rois"], rois[:1000])           # TEST post_nms_topN = 1000 rows, score order


# ----------------------------------------------------------------------------- RPN labels (roi_data/rpn.py)
def _rpn_cfg(fpn, max_size=384):
    from vosdetectron_b200.config import RegionConfig
    c = RegionConfig()
    c.fpn_on = c.multilevel_rpn = fpn
    c.train_max_size = max_size
    c.rpn_sizes = (32, 64, 128, 256)
    return c


def _rpn_roidb(g, tag):
    n = int(g[tag + "n"])
    roidb = [{"height": int(g["%shw%d" % (tag, i)][0]), "width": int(g["%shw%d" % (tag, i)][1]),
              "boxes": g["%sboxes%d" % (tag, i)], "gt_classes": g["%sgt_classes%d" % (tag, i)],
              "is_crowd": g["%sis_crowd%d" % (tag, i)]} for i in range(n)]
    return (roidb, [float(s) for s in g[tag + "im_scales"]], [g["%skeys%d" % (tag, i)] for i in range(n)],
            [g["%subg%d" % (tag, i)] for i in range(n)])


@pytest.mark.parametrize("tag,fpn", [("fpn_", True), ("nogt_", True), ("quirk_", True), ("single_", False)])
def test_add_rpn_blobs_matches_the_reference(golden, tag, fpn):
    """add_rpn_blobs / _get_rpn_blobs against tests/golden/rpn_labels.npz (unmodified reference, RNG contract of
    roi_data/rpn.py): labels, weights and im_info bit-exact; regression targets rtol 1e-6 (logf vs NumPy's log)."""
    from vosdetectron_b200.roi_data import rpn
    g = golden("rpn_labels")
    c = _rpn_cfg(fpn)
    roidb, scales, keys, ubg = _rpn_roidb(g, tag)
    names = rpn.get_rpn_blob_names(cfg=c)
    blobs = {k: [] for k in names}
    assert rpn.add_rpn_blobs(blobs, scales, roidb, rand_keys=keys, rand_bg=ubg, cfg=c)
    assert set(names) == set(k[len(tag):] for k in g.files if k.startswith(tag + "rpn_") or k == tag + "im_info") | {"roidb"}
    nfg = 0
    for k in names:
        if k == "roidb":
            assert len(blobs[k]) == len(roidb) and all("boxes" in e and "height" not in e for e in blobs[k])
            continue
        ref = g[tag + k]
        assert blobs[k].dtype == ref.dtype and blobs[k].shape == ref.shape, k
        if "targets" in k:
            assert np.allclose(blobs[k], ref, rtol=1e-6, atol=1e-7), k
            assert np.array_equal(blobs[k] != 0, ref != 0), k
        else:
            assert np.array_equal(blobs[k], ref), k
        if "labels" in k:
            nfg += int((ref == 1).sum())
    assert (nfg > 0) == (tag != "nogt_")


def test_rpn_labels_full_field_against_the_oracle(orc, synth):
    """The full training field (TRAIN.MAX_SIZE 1333: 451 143 anchors over five levels) against the restatement on
    random gt boxes, plus the size-independent properties: at most RPN_BATCH_SIZE_PER_IM labelled anchors, at most
    num_fg foreground, outside weights sum to 4 (one per coordinate), every labelled anchor inside the image."""
    from vosdetectron_b200.roi_data import rpn
    c = _rpn_cfg(True, 1333)
    foas = rpn._fields(c)
    anchors = np.concatenate([f.field_of_anchors for f in foas])
    assert anchors.shape[0] == 451143
    ref_anchors = np.concatenate([orc.field_of_anchors(2. ** l, (32 * 2. ** (l - 2),), (0.5, 1, 2), 1333)[0] for l in range(2, 7)])
    assert np.array_equal(anchors, ref_anchors)
    rs = np.random.RandomState(11)
    h, w = 800., 1216.
    for G in (1, 12, 60):
        gt = synth.random_rois(100 + G, G, (int(h), int(w)), 1, smin=20, smax=500)[:, 1:5].astype(np.float32)
        keys = rs.uniform(size=len(anchors)).astype(np.float32)
        ubg = rs.uniform(size=256).astype(np.float32)
        out = rpn._get_rpn_blobs(h, w, foas, anchors, gt, keys, ubg, cfg=c)
        lab, tg, iw, ow = orc.rpn_labels(h, w, anchors, gt, keys, ubg)
        ref = orc.rpn_blobs_split([(f.num_cell_anchors, f.field_size) for f in foas], lab, tg, iw, ow)
        n_lab = n_fg = 0
        osum = 0.0
        for mine, r in zip(out, ref):
            assert np.array_equal(mine["rpn_labels_int32_wide"], r["rpn_labels_int32_wide"])
            assert np.array_equal(mine["rpn_bbox_inside_weights_wide"], r["rpn_bbox_inside_weights_wide"])
            assert np.array_equal(mine["rpn_bbox_outside_weights_wide"], r["rpn_bbox_outside_weights_wide"])
            assert np.allclose(mine["rpn_bbox_targets_wide"], r["rpn_bbox_targets_wide"], rtol=1e-6, atol=1e-7)
            n_lab += int((mine["rpn_labels_int32_wide"] >= 0).sum())
            n_fg += int((mine["rpn_labels_int32_wide"] == 1).sum())
            osum += float(mine["rpn_bbox_outside_weights_wide"].astype(np.float64).sum())
        assert 0 < n_lab <= 256 and n_fg <= 128
        assert abs(osum - 4.0) < 1e-4
