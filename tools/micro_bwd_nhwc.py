#!/usr/bin/env python
"""Times the RoIAlign backward of BASELINE config 4 (1024 x 256 x 7 x 7 and 256 x 256 x 14 x 14 over P2-P5 of two 800x1344
frames) in both memory orders: the separable NCHW kernel and the channels-last kernel.  L2 flushed between iterations.

    python tools/micro_bwd_nhwc.py [--iters 20] [--out gpurun_out/micro_bwd_nhwc.json]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vosdetectron_b200 import ops, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    N = 2
    shapes = [(N, 256, int(np.ceil(synth.COCO_BLOB[0] / 2 ** l)), int(np.ceil(synth.COCO_BLOB[1] / 2 ** l))) for l in synth.ROI_LEVELS]
    sc = [1.0 / 2 ** l for l in synth.ROI_LEVELS]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    res = {}
    for name, pooled, R in (("box7", 7, 1024), ("mask14", 14, 256)):
        rois = torch.from_numpy(synth.random_rois(31 + pooled, R, synth.COCO_BLOB, N)).cuda()
        level, _, _, _ = ops.distribute_cuda(rois)
        lv = (level - 2).to(torch.int32)
        top = torch.randn((R, 256, pooled, pooled), device="cuda")
        for cl in (False, True):
            ts = []
            for i in range(a.iters + 3):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                g = ops.roi_align_ml_backward(top, shapes, sc, rois, lv, pooled, pooled, 2, channels_last=cl)
                e1.record()
                torch.cuda.synchronize()
                if i >= 3:
                    ts.append(e0.elapsed_time(e1))
            res["%s_%s" % (name, "channels_last" if cl else "nchw")] = {"ms_median": float(np.median(ts)), "ms_min": float(min(ts))}
            print(name, "channels_last" if cl else "nchw", res["%s_%s" % (name, "channels_last" if cl else "nchw")])
    if a.out:
        json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
