"""Generates tests/golden/labels.npz by running the UNMODIFIED label-assignment code of the reference in the build
container (through oracle/ref_harness.py):

    datasets/json_dataset.py: add_proposals -> _merge_proposal_boxes_into_roidb (:413-490), _add_class_assignments (:513-)
    roi_data/fast_rcnn.py:    add_fast_rcnn_blobs (:108-129) -> _sample_rois (:132-213) -> _compute_targets /
                              _expand_bbox_targets, _add_multilevel_rois (:262-290)
    roi_data/mask_rcnn.py:    add_mask_rcnn_blobs (:34-102) for mask_rois / roi_has_mask_int32 (the polygon rasteriser is
                              pycocotools', absent from the image: masks_int32 is not part of the fixture)

RNG contract of this repository (vosdetectron_b200/roi_data/fast_rcnn.py): the sampler consumes one uniform key per box;
"npr.choice(inds, size, replace=False)" = the `size` candidates with the smallest keys, in ascending key order.  The
reference draws a permutation from numpy's global generator instead; to compare the two on identical randomness the
generator below installs exactly that contract as `npr.choice` (the function bodies of the reference stay untouched).

    python oracle/build_ref.py && python tests/golden/make_golden_labels.py
"""
import importlib.util
import os
import sys
import types

import numpy as np
import scipy.sparse

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)

import ref_harness as rh  # noqa: E402
import region_oracle as orc  # noqa: E402
from vosdetectron_b200 import synth  # noqa: E402


def make_entry(rs, K, G, im_hw):
    """roidb entry with G ground-truth boxes (no proposals yet), as JsonDataset._add_gt_annotations leaves it."""
    h, w = im_hw
    cx, cy = rs.uniform(0.15 * w, 0.85 * w, G), rs.uniform(0.15 * h, 0.85 * h, G)
    bw, bh = rs.uniform(30, 0.4 * w, G), rs.uniform(30, 0.4 * h, G)
    boxes = np.stack([np.clip(cx - bw / 2, 0, w - 1), np.clip(cy - bh / 2, 0, h - 1),
                      np.clip(cx + bw / 2, 0, w - 1), np.clip(cy + bh / 2, 0, h - 1)], 1).astype(np.float32)
    cls = rs.randint(1, K, G).astype(np.int32)
    ov = np.zeros((G, K), np.float32)
    ov[np.arange(G), cls] = 1.0
    # one rectangle polygon per instance (only polys_to_boxes reads them here)
    segms = [[[b[0], b[1], b[2], b[1], b[2], b[3], b[0], b[3]]] for b in boxes.astype(np.float64)]
    return {"boxes": boxes, "segms": segms, "gt_classes": cls, "seg_areas": ((boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1])),
            "gt_overlaps": scipy.sparse.csr_matrix(ov), "is_crowd": np.zeros(G, dtype=bool),
            "box_to_gt_ind_map": np.arange(G, dtype=np.int32)}


def main():
    r = rh.ref()
    cfg = r.cfg
    spec = importlib.util.spec_from_file_location("datasets._real_json_dataset", os.path.join(rh.REF_ROOT, "lib/datasets/json_dataset.py"))
    jd = importlib.util.module_from_spec(spec)
    jd.__package__ = "datasets"
    spec.loader.exec_module(jd)
    import roi_data.fast_rcnn as f
    import roi_data.mask_rcnn as mr

    K = 81
    cfg.MODEL.NUM_CLASSES = K
    cfg.MODEL.BBOX_REG_WEIGHTS = (10., 10., 5., 5.)
    cfg.MODEL.CLS_AGNOSTIC_BBOX_REG = False
    cfg.MODEL.KEYPOINTS_ON = False
    cfg.MODEL.IDENTITY_TRAINING = False
    cfg.FPN.FPN_ON = True
    cfg.FPN.MULTILEVEL_ROIS = True
    cfg.FPN.ROI_MIN_LEVEL, cfg.FPN.ROI_MAX_LEVEL = 2, 5
    cfg.TRAIN.BATCH_SIZE_PER_IM = 512
    cfg.TRAIN.FG_FRACTION = 0.25
    cfg.TRAIN.FG_THRESH, cfg.TRAIN.BG_THRESH_HI, cfg.TRAIN.BG_THRESH_LO = 0.5, 0.5, 0.0
    cfg.MRCNN.RESOLUTION = 28
    cfg.MRCNN.CLS_SPECIFIC_MASK = True

    rs = np.random.RandomState(77)
    B = 2
    im_sizes = [(480, 800), (640, 512)]                        # original image sizes
    im_scales = np.asarray([1.6667, 1.25], np.float32)         # im_info[:, 2] (train-style per-image scale)
    G = [7, 4]
    entries = [make_entry(rs, K, G[i], im_sizes[i]) for i in range(B)]
    g = {"num_classes": K, "im_scales": im_scales}
    for i, e in enumerate(entries):
        g["gt_boxes%d" % i], g["gt_classes%d" % i] = e["boxes"].copy(), e["gt_classes"].copy()

    # proposals in blob coordinates: jittered copies of the gt boxes (foreground), near misses and random boxes
    rows = []
    for i, e in enumerate(entries):
        s = float(im_scales[i])
        h, w = im_sizes[i]
        n_fg, n_rand = 260 if i == 0 else 40, 700
        src = e["boxes"][rs.randint(0, G[i], n_fg)]
        jit = src + rs.normal(0, 6.0 + 14.0 * rs.uniform(size=(n_fg, 1)), src.shape)
        rnd = synth.random_rois(900 + i, n_rand, (h, w), 1, smin=12, smax=0.7 * min(h, w))[:, 1:5]
        bx = np.concatenate([jit, rnd]).astype(np.float32)
        bx[:, 2:] = np.maximum(bx[:, 2:], bx[:, :2] + 1)
        bx = np.clip(bx, 0, [w - 1, h - 1, w - 1, h - 1]).astype(np.float32) * np.float32(s)
        rows.append(np.hstack([np.full((len(bx), 1), i, np.float32), bx]))
    rois = np.concatenate(rows).astype(np.float32)
    rois = rois[rs.permutation(len(rois))]                     # collect() interleaves the images (score order)
    g["rpn_rois"] = rois

    # ---- add_proposals (unmodified) ----
    jd.add_proposals(entries, rois, im_scales, crowd_thresh=0)
    for i, e in enumerate(entries):
        g["boxes%d" % i] = e["boxes"]
        g["max_overlaps%d" % i] = e["max_overlaps"]
        g["max_classes%d" % i] = e["max_classes"].astype(np.int32)
        g["box_to_gt%d" % i] = e["box_to_gt_ind_map"].astype(np.int32)

    # ---- the RNG contract as npr.choice ----
    keys = [rs.uniform(size=e["boxes"].shape[0]).astype(np.float32) for e in entries]
    for i in range(B):
        g["keys%d" % i] = keys[i]
    cur = {"i": 0}

    def choice(a, size=None, replace=True):
        assert replace is False
        a = np.asarray(a)
        k = keys[cur["i"]][a]
        return a[np.argsort(k, kind="stable")[:int(size)]]
    f.npr = types.SimpleNamespace(choice=choice)
    orig_sample = f._sample_rois

    def sample(entry, im_scale, batch_idx):                    # harness wrapper: which image's keys are current
        cur["i"] = batch_idx
        return orig_sample(entry, im_scale, batch_idx)
    f._sample_rois = sample
    mr.segm_utils.polys_to_mask_wrt_box = lambda poly, box, M: np.zeros((M, M), np.int32)   # pycocotools rasteriser absent

    for tag, mask_on in (("", False), ("m_", True)):
        cfg.MODEL.MASK_ON = mask_on
        names = f.get_fast_rcnn_blob_names()
        blobs = {k: [] for k in names}
        assert f.add_fast_rcnn_blobs(blobs, im_scales, entries)
        for k, v in blobs.items():
            if k == "masks_int32":
                continue
            g[tag + k] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, "labels.npz"), **g)
    print("labels.npz", {k: np.asarray(v).shape for k, v in g.items() if k in ("rpn_rois", "rois", "labels_int32", "bbox_targets", "m_mask_rois")})
    print("fg per image:", [(int((g["labels_int32"][:512] > 0).sum())), int((g["labels_int32"][512:] > 0).sum())])


if __name__ == "__main__":
    main()
