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
