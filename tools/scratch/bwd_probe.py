import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from vosdetectron_b200 import ops, synth, _lib
mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
_lib.load().vosd_debug_force_generic(mode)
lvls = synth.ROI_LEVELS
feats = synth.fpn_features(2000, synth.COCO_BLOB, 2, lvls, 256)
shapes = [feats[l].shape for l in lvls]
sc = [1.0 / 2 ** l for l in lvls]
for R, res in ((1024, 7), (256, 14)):
    rois = torch.from_numpy(synth.random_rois(2001, R, synth.COCO_BLOB, 2)).cuda()
    level, lc, order, restore = ops.distribute_cuda(rois)
    lv0 = (level - 2).to(torch.int32)
    g = torch.randn((R, 256, res, res), device="cuda")
    for _ in range(3):
        ops.roi_align_ml_backward(g, shapes, sc, rois, lv0, res, res, 2)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        ops.roi_align_ml_backward(g, shapes, sc, rois, lv0, res, res, 2)
    b.record(); torch.cuda.synchronize()
    print("mode", mode, "res", res, "ms", a.elapsed_time(b) / 10)
