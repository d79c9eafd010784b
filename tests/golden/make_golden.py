"""Generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, via oracle/ref_harness.py) in the build container.

    python oracle/build_ref.py && python tests/golden/make_golden.py

The fixtures are small (a down-sized blob) so they can live in git; inputs are
stored next to the outputs so that the tests do not depend on the synthetic
generator staying byte-stable.  Environment that produced the committed files:
Python 3.12.3, NumPy 2.3.5, OpenCV 4.13.0, torchvision 0.26.0, Cython 3.3.0, gcc 13.3.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)

import ref_harness as rh  # noqa: E402
from vosdetectron_b200 import synth  # noqa: E402

BLOB = (192, 256)          # levels 2..6: 48x64, 24x32, 12x16, 6x8, 3x4
PRE, POST, THRESH = 300, 100, 0.7


def main():
    import torch
    r = rh.ref()
    r.merge_cfg_from_file(r.yaml_r50)
    cfg = r.cfg
    cfg.FPN.MULTILEVEL_ROIS = True

    # ---- anchors (generate_anchors.py:54-123) + the KAT quoted at :26-51 ----
    g = {"kat_stride16": r.generate_anchors(16, (128, 256, 512), (0.5, 1, 2))}
    for lvl in synth.FPN_LEVELS:
        g["fpn%d" % lvl] = r.generate_anchors(2. ** lvl, (32 * 2. ** (lvl - 2),), (0.5, 1, 2))
    np.savez_compressed(os.path.join(HERE, "anchors.npz"), **g)

    # ---- GenerateProposalsOp.forward, eval mode, 2 images -------------------
    cfg.TEST.RPN_PRE_NMS_TOP_N, cfg.TEST.RPN_POST_NMS_TOP_N, cfg.TEST.RPN_NMS_THRESH = PRE, POST, THRESH
    rpn = synth.rpn_outputs(4242, BLOB, num_images=2)
    im_info = np.array([[192, 256, 1.5], [160, 250, 1.25]], dtype=np.float32)
    g = {"im_info": im_info, "pre": PRE, "post": POST, "thresh": THRESH, "min_size": 0}
    rois_l, probs_l = [], []
    for lvl in synth.FPN_LEVELS:
        sc, d = rpn[lvl]
        op = r.GenerateProposalsOp(r.generate_anchors(2. ** lvl, (32 * 2. ** (lvl - 2),), (0.5, 1, 2)), 1. / 2 ** lvl)
        op.eval()
        rois, probs = op.forward(torch.from_numpy(sc), torch.from_numpy(d), torch.from_numpy(im_info))
        g["scores%d" % lvl], g["deltas%d" % lvl] = sc, d
        g["rois%d" % lvl], g["probs%d" % lvl] = rois, probs
        rois_l.append(rois)
        probs_l.append(probs)
    # min_size > 0 variant on level 3 (exercises _filter_boxes, :171-182)
    cfg.TEST.RPN_MIN_SIZE = 16
    sc, d = rpn[3]
    op = r.GenerateProposalsOp(r.generate_anchors(8., (64.,), (0.5, 1, 2)), 1. / 8)
    op.eval()
    g["rois3_min16"], g["probs3_min16"] = op.forward(torch.from_numpy(sc), torch.from_numpy(d), torch.from_numpy(im_info))
    cfg.TEST.RPN_MIN_SIZE = 0
    np.savez_compressed(os.path.join(HERE, "proposals.npz"), **g)

    # ---- collect + distribute (collect_and_distribute...py:91-138) ----------
    cfg.TEST.RPN_POST_NMS_TOP_N = 150
    allp = np.concatenate(probs_l).ravel()
    assert np.unique(allp).size == allp.size, 'collect fixture must be tie-free'
    rois = r.collect(rois_l + probs_l, False)
    blobs = r.distribute(rois, None)
    g = {"post": 150}
    for i, lvl in enumerate(synth.FPN_LEVELS):
        g["in_rois%d" % lvl], g["in_probs%d" % lvl] = rois_l[i], probs_l[i]
    g.update(blobs)
    g["levels"] = r.fpn_utils.map_rois_to_fpn_levels(rois[:, 1:5], 2, 5)
    # distribute() alone on RoIs that span all four levels (incl. degenerate ones)
    wide = np.concatenate([synth.random_rois(314, 300, synth.COCO_BLOB, num_images=2), synth.edge_rois()])
    for k, v in r.distribute(wide, None).items():
        g["wide_" + k] = v
    g["wide_levels"] = r.fpn_utils.map_rois_to_fpn_levels(wide[:, 1:5], 2, 5)
    np.savez_compressed(os.path.join(HERE, "collect_distribute.npz"), **g)

    # ---- box_utils.nms -> cython_nms.nms (cython_nms.pyx:37-87) -------------
    d_uns = synth.clustered_dets(11, 700, BLOB, n_centres=12)
    d_srt = d_uns[np.argsort(-d_uns[:, 4], kind="stable")]
    g = {"dets_unsorted": d_uns, "dets_sorted": d_srt}
    for t in (0.3, 0.5, 0.7):
        g["keep_unsorted_%02d" % int(t * 10)] = r.box_utils.nms(d_uns, t)
        g["keep_sorted_%02d" % int(t * 10)] = r.box_utils.nms(d_srt, t)
    np.savez_compressed(os.path.join(HERE, "nms.npz"), **g)

    # ---- bbox_transform / clip / expand (boxes.py:138-258) ------------------
    rs = np.random.RandomState(77)
    boxes = synth.random_rois(78, 64, BLOB)[:, 1:5]
    deltas = rs.normal(0, 1.0, (64, 4 * 5)).astype(np.float32)
    deltas[3, 2] = 40.0   # hits BBOX_XFORM_CLIP
    xf = r.box_utils.bbox_transform(boxes, deltas, (10., 10., 5., 5.))
    cl = r.box_utils.clip_tiled_boxes(xf.copy(), np.array([192, 256], dtype=np.float32))
    ex = r.box_utils.expand_boxes(boxes, 30.0 / 28.0)
    np.savez_compressed(os.path.join(HERE, "boxes.npz"), boxes=boxes, deltas=deltas, transformed=xf,
                        clipped=cl, expanded=ex)

    # ---- segm_results (core/test.py:801-855), RLE step captured -------------
    fh, fw, R, M, K = 120, 160, 14, 28, 5
    bx, cls, masks = synth.detections(909, R, (fh, fw), M, K)
    bx[0] = [-6.5, -4.2, 30.3, 25.8]            # sticks out top-left
    bx[1] = [140.2, 100.7, 171.9, 133.1]        # sticks out bottom-right
    bx[2] = [50.0, 40.0, 50.4, 40.3]            # tiny
    bx[3] = [20.0, 20.0, 20.0 + 14 * 28 / 30.0 - 1, 20.0 + 14 * 28 / 30.0 - 1]   # ~15x15: cv2 2x INTER_AREA path
    cfg.MODEL.NUM_CLASSES, cfg.MRCNN.RESOLUTION = K, M
    cls_boxes = [np.zeros((int((cls == j).sum()), 5), np.float32) for j in range(K)]
    cap = np.stack(rh.segm_results_capture(r, cls_boxes, masks, bx, fh, fw))
    np.savez_compressed(os.path.join(HERE, "paste.npz"), boxes=bx, cls=cls, masks=masks,
                        frame_hw=np.array([fh, fw]), packed=np.packbits(cap, axis=-1))

    # ---- box head: decode (core/test.py:178-179) + box_results_with_nms_and_limit (:733-797) ----
    # lib/core/test.py:785 reads cfg.TEST.NUM_DET_PER_CLASS, a key lib/core/config.py never defines (it has
    # NUM_DET_PER_CLASS_PRE/_POST, :948-949): the unmodified function raises AttributeError there.  The fixture
    # adds the key (= 0, i.e. the branch is skipped) to the cfg OBJECT; nothing under /root/reference changes.
    Kd, Rd = 9, 400
    cfg.MODEL.NUM_CLASSES = Kd
    cfg.TEST.NUM_DET_PER_CLASS = 0
    props, sc, dl = synth.box_head_outputs(4711, Rd, Kd, BLOB)
    pred = r.box_utils.bbox_transform(props, dl, cfg.MODEL.BBOX_REG_WEIGHTS)
    pred = r.box_utils.clip_tiled_boxes(pred, np.array([BLOB[0], BLOB[1]], dtype=np.float32))
    g = {"props": props, "scores": sc, "deltas": dl, "pred_boxes": pred, "num_classes": Kd,
         "score_thresh": cfg.TEST.SCORE_THRESH, "weights": np.asarray(cfg.MODEL.BBOX_REG_WEIGHTS, np.float32)}
    for tag, nms_t, per_im in (("a", 0.5, 100), ("b", 0.3, 30), ("c", 0.5, 0)):
        cfg.TEST.NMS, cfg.TEST.DETECTIONS_PER_IM = nms_t, per_im
        s_out, b_out, cls_boxes = r.core_test.box_results_with_nms_and_limit(sc, pred)
        g["nms_" + tag], g["per_im_" + tag] = nms_t, per_im
        g["out_scores_" + tag], g["out_boxes_" + tag] = s_out, b_out
        g["cls_count_" + tag] = np.array([len(cls_boxes[j]) for j in range(Kd)], dtype=np.int32)
    assert len(g["out_scores_a"]) >= 100 and len(g["out_scores_c"]) > len(g["out_scores_a"])
    np.savez_compressed(os.path.join(HERE, "box_results.npz"), **g)
    cfg.MODEL.NUM_CLASSES = K

    # ---- RoIAlign: the reference is a CUDA kernel (not runnable here).  Loose
    # CPU anchor: torchvision aligned=False implements the same Caffe2 formula.
    import torchvision
    feats = synth.fpn_features(5150, BLOB, num_images=2, C=8)
    rois = synth.random_rois(5151, 40, BLOB, num_images=2, smin=8, smax=200)
    g = {"rois": rois}
    for lvl in synth.ROI_LEVELS:
        f = torch.from_numpy(feats[lvl]).requires_grad_(True)
        for res in (7, 14):
            out = torchvision.ops.roi_align(f, torch.from_numpy(rois), (res, res), 1. / 2 ** lvl,
                                            sampling_ratio=2, aligned=False)
            g["tv_fwd_l%d_r%d" % (lvl, res)] = out.detach().numpy()
        g["feat%d" % lvl] = feats[lvl]
    np.savez_compressed(os.path.join(HERE, "roialign_tv.npz"), **g)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
