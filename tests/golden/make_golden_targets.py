"""Generates tests/golden/bbox_targets.npz by running the UNMODIFIED lib/roi_data/fast_rcnn.py
(_compute_targets :216-229 -> box_utils.bbox_transform_inv, _expand_bbox_targets :232-260) in the build
container, through oracle/ref_harness.py.

    python oracle/build_ref.py && python tests/golden/make_golden_targets.py
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


def main():
    r = rh.ref()
    cfg = r.cfg
    import roi_data.fast_rcnn as f
    rs = np.random.RandomState(31)
    n, K = 512, 9
    cfg.MODEL.NUM_CLASSES = K
    cfg.MODEL.BBOX_REG_WEIGHTS = (10., 10., 5., 5.)
    ex = synth.random_rois(600, n, (800, 1344), 1, smin=8)[:, 1:5].copy()
    gt = (ex + rs.uniform(-12, 12, ex.shape)).astype(np.float32)
    gt[:, 2:] = np.maximum(gt[:, 2:], gt[:, :2] + 1)
    labels = rs.randint(0, K, n).astype(np.int32)
    labels[:40] = 0
    g = {"ex": ex, "gt": gt, "labels": labels, "num_classes": K}
    for tag, agn in (("k", False), ("a", True)):
        cfg.MODEL.CLS_AGNOSTIC_BBOX_REG = agn
        data = f._compute_targets(ex, gt, labels.copy())
        t, w = f._expand_bbox_targets(data.copy())
        g["targets_" + tag], g["inside_" + tag] = t, w
        g["outside_" + tag] = np.array(w > 0, dtype=w.dtype)             # fast_rcnn.py:206-208
    cfg.MODEL.CLS_AGNOSTIC_BBOX_REG = False
    np.savez_compressed(os.path.join(HERE, "bbox_targets.npz"), **g)
    print("bbox_targets.npz", g["targets_k"].shape, g["targets_a"].shape)


if __name__ == "__main__":
    main()
