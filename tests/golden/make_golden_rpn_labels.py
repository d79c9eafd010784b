"""Generates tests/golden/rpn_labels.npz by running the UNMODIFIED RPN label assignment of the reference in the build
container (through oracle/ref_harness.py):

    roi_data/rpn.py:        add_rpn_blobs (:64-140) -> _get_rpn_blobs (:143-270)
    roi_data/data_utils.py: get_field_of_anchors (:50-102), compute_targets, unmap

RNG contract of this repository (vosdetectron_b200/roi_data/rpn.py): `npr.choice(fg_inds, size, replace=False)` = the
`size` candidates with the smallest per-anchor keys (keys indexed by the anchor's position in the whole field of
anchors, ties: lower index first); `npr.randint(n, size=k)` = floor(u[:k] * n) of the image's uniforms.  The generator
installs exactly that as `rpn.npr` (the function bodies of the reference stay untouched).

Cases: "quirk" (see main), "fpn" two images on a five-level field (TRAIN.MAX_SIZE 384), one of them with a gt box that lies outside every
inside anchor's reach plus many foreground anchors (both subsamples fire); "nogt" an image without gt boxes; "single"
the classical single-level RPN (RPN.STRIDE 16, four sizes).

    python oracle/build_ref.py && python tests/golden/make_golden_rpn_labels.py
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)

import ref_harness as rh  # noqa: E402


def entry(h, w, boxes, classes, crowd=None):
    boxes = np.asarray(boxes, np.float32).reshape(-1, 4)
    return {"height": h, "width": w, "boxes": boxes, "gt_classes": np.asarray(classes, np.int32),
            "is_crowd": np.zeros(len(boxes), bool) if crowd is None else np.asarray(crowd, bool)}


def run_case(rpn, cfg, tag, entries, im_scales, rs, g):
    names = rpn.get_rpn_blob_names()
    blobs = {k: [] for k in names}
    # the same field the reference builds
    import roi_data.data_utils as du
    if cfg.FPN.FPN_ON and cfg.FPN.MULTILEVEL_RPN:
        foas = [du.get_field_of_anchors(2. ** l, (cfg.FPN.RPN_ANCHOR_START_SIZE * 2. ** (l - cfg.FPN.RPN_MIN_LEVEL),),
                                        cfg.FPN.RPN_ASPECT_RATIOS)
                for l in range(cfg.FPN.RPN_MIN_LEVEL, cfg.FPN.RPN_MAX_LEVEL + 1)]
    else:
        foas = [du.get_field_of_anchors(cfg.RPN.STRIDE, cfg.RPN.SIZES, cfg.RPN.ASPECT_RATIOS)]
    all_anchors = np.concatenate([f.field_of_anchors for f in foas])
    T = all_anchors.shape[0]
    keys = [rs.uniform(size=T).astype(np.float32) for _ in entries]
    ubg = [rs.uniform(size=cfg.TRAIN.RPN_BATCH_SIZE_PER_IM).astype(np.float32) for _ in entries]
    cur = {"i": -1, "inside": None}
    orig = rpn._get_rpn_blobs

    def wrapped(im_height, im_width, foas_, all_anchors_, gt_boxes):       # harness: which image's randomness is current
        cur["i"] += 1
        st = cfg.TRAIN.RPN_STRADDLE_THRESH
        cur["inside"] = np.where((all_anchors_[:, 0] >= -st) & (all_anchors_[:, 1] >= -st)
                                 & (all_anchors_[:, 2] < im_width + st) & (all_anchors_[:, 3] < im_height + st))[0]
        return orig(im_height, im_width, foas_, all_anchors_, gt_boxes)

    def choice(a, size=None, replace=True):
        assert replace is False
        a = np.asarray(a)
        k = keys[cur["i"]][cur["inside"][a]]
        return a[np.argsort(k, kind="stable")[:int(size)]]

    def randint(n, size=None):
        return np.floor(ubg[cur["i"]][:int(size)].astype(np.float64) * n).astype(np.int64)

    rpn._get_rpn_blobs = wrapped
    rpn.npr = types.SimpleNamespace(choice=choice, randint=randint)
    try:
        assert rpn.add_rpn_blobs(blobs, im_scales, entries)
    finally:
        rpn._get_rpn_blobs = orig
    g[tag + "n"] = np.int32(len(entries))
    g[tag + "im_scales"] = np.asarray(im_scales, np.float64)
    for i, e in enumerate(entries):
        g["%shw%d" % (tag, i)] = np.asarray([e["height"], e["width"]], np.int32)
        g["%sboxes%d" % (tag, i)] = e["boxes"]
        g["%sgt_classes%d" % (tag, i)] = e["gt_classes"]
        g["%sis_crowd%d" % (tag, i)] = e["is_crowd"]
        g["%skeys%d" % (tag, i)] = keys[i]
        g["%subg%d" % (tag, i)] = ubg[i]
    stats = {}
    for k, v in blobs.items():
        if k == "roidb":
            continue
        g[tag + k] = np.asarray(v)
        if "labels" in k:
            stats[k] = (int((v == 1).sum()), int((v == 0).sum()))
    print(tag, "anchors", T, "fg/bg per level", stats)


def main():
    r = rh.ref()
    cfg = r.cfg
    import roi_data.rpn as rpn
    cfg.MODEL.IDENTITY_TRAINING = False
    cfg.TRAIN.MAX_SIZE = 384
    cfg.FPN.COARSEST_STRIDE = 32
    cfg.FPN.RPN_MIN_LEVEL, cfg.FPN.RPN_MAX_LEVEL = 2, 6
    cfg.FPN.RPN_ANCHOR_START_SIZE = 32
    cfg.FPN.RPN_ASPECT_RATIOS = (0.5, 1, 2)
    cfg.TRAIN.RPN_POSITIVE_OVERLAP, cfg.TRAIN.RPN_NEGATIVE_OVERLAP = 0.7, 0.3
    cfg.TRAIN.RPN_FG_FRACTION, cfg.TRAIN.RPN_BATCH_SIZE_PER_IM = 0.5, 256
    cfg.TRAIN.RPN_STRADDLE_THRESH = 0
    rs = np.random.RandomState(2024)
    g = {}

    # ---- five-level FPN field ----
    cfg.FPN.FPN_ON, cfg.FPN.MULTILEVEL_RPN = True, True
    h0, w0 = 200, 300
    # image 0: 40 small boxes that coincide with level-2 anchors (many foreground anchors: the fg subsample fires), a
    # crowd box and a background-class box (both skipped), a box hugging the border
    small = []
    for _ in range(40):
        cx, cy = rs.uniform(30, w0 * 1.25 - 30), rs.uniform(30, h0 * 1.25 - 30)
        s = rs.uniform(26, 40)
        small.append([cx - s / 2, cy - s / 2, cx + s / 2, cy + s / 2])
    small = (np.asarray(small) / 1.25).astype(np.float32)
    b0 = np.concatenate([small, [[10, 20, 150, 160], [0, 0, 299, 199], [100, 50, 200, 120], [5, 5, 60, 40]]]).astype(np.float32)
    c0 = np.concatenate([rs.randint(1, 81, 40), [3, 7, 0, 9]]).astype(np.int32)
    crowd0 = np.zeros(len(b0), bool)
    crowd0[-1] = True
    # image 1: few boxes (fewer foreground anchors than num_fg), one of them 2 px wide (no anchor reaches 0.7)
    b1 = np.asarray([[40, 30, 180, 170], [200, 100, 203, 230], [10, 10, 120, 60]], np.float32)
    entries = [entry(h0, w0, b0, c0, crowd0), entry(240, 220, b1, [1, 2, 3])]
    run_case(rpn, cfg, "fpn_", entries, [1.25, 1.5], rs, g)

    # ---- no gt boxes: every inside anchor is a background candidate ----
    run_case(rpn, cfg, "nogt_", [entry(180, 260, np.zeros((0, 4), np.float32), np.zeros(0, np.int32))], [1.4], rs, g)

    # ---- the reference's quirks: (a) a gt box outside the image overlaps no inside anchor, so EVERY inside anchor with
    #      zero overlap to it becomes foreground (`anchor_by_gt_overlap == gt_to_anchor_max` with max 0) and the
    #      foreground subsample decides; (b) a tiny image has fewer background candidates than num_bg: none is labelled
    run_case(rpn, cfg, "quirk_", [entry(200, 300, [[40, 40, 120, 100], [900, 900, 960, 950]], [5, 6]),
                                  entry(48, 64, [[4, 4, 40, 40]], [2])], [1.25, 1.0], rs, g)

    # ---- classical single-level RPN ----
    cfg.FPN.FPN_ON, cfg.FPN.MULTILEVEL_RPN = False, False
    cfg.RPN.STRIDE, cfg.RPN.SIZES, cfg.RPN.ASPECT_RATIOS = 16, (32, 64, 128, 256), (0.5, 1, 2)
    run_case(rpn, cfg, "single_", [entry(h0, w0, b0[30:], c0[30:], crowd0[30:]), entry(240, 220, b1, [1, 2, 3])],
             [1.25, 1.5], rs, g)

    out = os.path.join(HERE, "rpn_labels.npz")
    np.savez_compressed(out, **g)
    print("wrote", out, os.path.getsize(out), "bytes,", len(g), "arrays")


if __name__ == "__main__":
    main()
