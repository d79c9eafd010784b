"""Generates tests/golden/vos_post.npz by running the UNMODIFIED lib_vos/tools/vos_test.py
(box_results_with_nms_and_limit :748-865, segm_results :867-921, nms_with_mask_iou :985-1029) in the build
container, through oracle/ref_harness.py.

    python oracle/build_ref.py && python tests/golden/make_golden_vos.py

pycocotools is absent from the image (third party, unpinned upstream), so `mask_util.encode/decode` are bound to
the oracle's restatement of maskApi.c (oracle/region_oracle.py: rle_encode / rle_decode); everything else is the
reference's own code.  Inputs are stored next to the outputs.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, ROOT)

import ref_harness as rh  # noqa: E402
import region_oracle as orc  # noqa: E402
from vosdetectron_b200 import synth  # noqa: E402


def main():
    r = rh.ref()
    r.merge_cfg_from_file(r.yaml_r50)
    cfg = r.cfg
    sys.path.insert(0, os.path.join(rh.REF_ROOT, "lib_vos", "tools"))
    sys.path.insert(0, os.path.join(rh.REF_ROOT, "lib_vos"))
    import vos_test

    class NumpyCompat(object):
        """vos_test.py:1000-1011 builds `discard` as a float64 array and passes it to np.delete, which NumPy < 1.19
        cast to integers (with a DeprecationWarning) and NumPy 2 rejects.  Same kind of shim as np.int / np.float in
        ref_harness: the reference file is untouched, its module-level `np` gets the old casting back."""
        def __getattr__(self, name):
            return getattr(np, name)

        @staticmethod
        def delete(arr, obj, axis=None):
            return np.delete(arr, np.asarray(obj).astype(np.intp), axis)
    vos_test.np = NumpyCompat()
    g = {}

    # ---- box_results_with_nms_and_limit with the VOS extras -------------------------------------------
    box = np.load(os.path.join(HERE, "box_results.npz"))
    K = int(box["num_classes"])
    sc, pred = box["scores"], box["pred_boxes"]
    cfg.MODEL.NUM_CLASSES = K
    cfg.TEST.SCORE_THRESH = float(box["score_thresh"])
    cfg.TEST.NMS, cfg.TEST.DETECTIONS_PER_IM = 0.5, 100
    cfg.TEST.NMS_CROSS_CLASS, cfg.TEST.NUM_DET_PER_CLASS_PRE, cfg.TEST.NMS_SMALL_BOX_IOU = 0., 0, 0.
    _, _, base = vos_test.box_results_with_nms_and_limit(sc, pred)
    # previous-frame boxes: one per class at most; even classes confident (filter applies), class 3 unsure (skipped),
    # class 5 none
    prev = [[] for _ in range(K)]
    rs = np.random.RandomState(77)
    for j in range(1, K):
        if j == 5 or len(base[j]) == 0:
            prev[j] = np.zeros((0, 5), dtype=np.float32)
            continue
        b = base[j][int(np.argmax(base[j][:, -1]))].copy()
        b[:4] += rs.uniform(-6, 6, 4).astype(np.float32)
        b[4] = 0.1 if j == 3 else 0.9
        prev[j] = b[None].astype(np.float32)
    g["prev_boxes"] = np.vstack([p for p in prev[1:]])
    g["prev_count"] = np.array([len(p) for p in prev], dtype=np.int32)
    settings = {"x": (0.5, 0, 0.0), "y": (0.5, 3, 0.0), "z": (0.6, 50, 0.3), "w": (0.0, 2, 0.25)}
    for tag, (cross, pre, small) in settings.items():
        cfg.TEST.NMS_CROSS_CLASS, cfg.TEST.NUM_DET_PER_CLASS_PRE = cross, pre
        cfg.TEST.NMS_SMALL_BOX_IOU, cfg.TEST.NMS_SMALL_BOX_SCORE_THRESHOLD = small, 0.2
        s_out, b_out, cls_boxes = vos_test.box_results_with_nms_and_limit(sc, pred, prev_cls_boxes=prev)
        g["set_" + tag] = np.array([cross, pre, small, 0.2])
        g["out_scores_" + tag], g["out_boxes_" + tag] = s_out, b_out
        g["cls_count_" + tag] = np.array([len(cls_boxes[j]) for j in range(K)], dtype=np.int32)
    assert len(g["out_scores_x"]) < len(np.vstack(base[1:])) and len(g["out_scores_z"]) < len(g["out_scores_x"])

    # ---- segm_results -> nms_with_mask_iou on overlapping detections ------------------------------------
    fh, fw, M, Km = 120, 168, 28, 6
    cfg.MODEL.NUM_CLASSES, cfg.MRCNN.RESOLUTION, cfg.MRCNN.CLS_SPECIFIC_MASK = Km, M, True
    b0, _, m0 = synth.detections(9100, 14, (fh, fw), M, Km)
    boxes, masks = [], []
    for i in range(14):                                   # every base detection + 2 near-duplicates
        for d in range(3):
            jit = rs.uniform(-1.5, 1.5, 4).astype(np.float32) * (d > 0)
            boxes.append(np.concatenate([b0[i] + jit, [rs.uniform(0.3, 1.0)]]).astype(np.float32))
            masks.append(m0[i] + (0.02 * rs.standard_normal(m0[i].shape).astype(np.float32)) * (d > 0))
    boxes, masks = np.stack(boxes), np.stack(masks).astype(np.float32)
    cls = np.sort(rs.randint(1, Km, len(boxes))).astype(np.int32)
    cls_boxes = [[] for _ in range(Km)]
    for j in range(1, Km):
        cls_boxes[j] = boxes[cls == j]
    order = np.concatenate([np.flatnonzero(cls == j) for j in range(1, Km)])
    assert np.array_equal(order, np.arange(len(boxes)))
    vos_test.mask_util.encode = lambda a: [dict(size=[a.shape[0], a.shape[1]],
                                                counts=orc.rle_encode(np.asarray(a[:, :, k]))['counts'].encode('ascii'))
                                           for k in range(a.shape[2])]
    vos_test.mask_util.decode = lambda segms: np.stack(
        [orc.rle_decode(orc.rle_from_string(s['counts']), s['size'][0], s['size'][1]) for s in segms], axis=-1)
    cls_segms = vos_test.segm_results(cls_boxes, masks, boxes[:, :4], fh, fw)
    flat_segms = [s for sl in cls_segms for s in sl]
    g["m_boxes"], g["m_cls"], g["m_masks"], g["m_frame"] = boxes, cls, masks, np.array([fh, fw])
    g["m_counts"] = np.array([s['counts'] for s in flat_segms])
    for tag, (th, per) in {"p": (0.9, 1), "q": (0.5, 3), "r": (0.7, 100)}.items():
        ob, os_ = vos_test.nms_with_mask_iou([np.asarray(c).reshape(-1, 5) if len(c) else [] for c in cls_boxes],
                                             cls_segms, iou_th=th, max_per_class=per)
        g["mset_" + tag] = np.array([th, per])
        g["mout_count_" + tag] = np.array([len(c) for c in ob], dtype=np.int32)
        g["mout_boxes_" + tag] = np.vstack([np.vstack(c) for c in ob if len(c)]) if any(len(c) for c in ob) else np.zeros((0, 5), np.float32)
        g["mout_counts_" + tag] = np.array([s['counts'] for sl in os_ for s in sl])
    assert g["mout_count_r"].sum() < len(boxes) and g["mout_count_p"].sum() <= Km - 1
    np.savez_compressed(os.path.join(HERE, "vos_post.npz"), **g)
    print("vos_post.npz", os.path.getsize(os.path.join(HERE, "vos_post.npz")),
          {k: int(g["cls_count_" + k].sum()) for k in settings}, {k: int(g["mout_count_" + k].sum()) for k in "pqr"})


if __name__ == "__main__":
    main()
