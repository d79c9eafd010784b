import sys, os
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from vosdetectron_b200 import ops, synth, _lib
mode = int(sys.argv[1])
_lib.load().vosd_debug_force_generic(mode)
lvls = synth.ROI_LEVELS
feats = synth.fpn_features(2000, synth.COCO_BLOB, 2, lvls, 256)
fl = [torch.from_numpy(feats[l]).cuda() for l in lvls]
rois = torch.from_numpy(synth.random_rois(2001, 1024, synth.COCO_BLOB, 2)).cuda()
level, lc, order, restore = ops.distribute_cuda(rois)
lv0 = (level - 2).to(torch.int32)
g = torch.randn((1024, 256, 7, 7), device="cuda")
sc = [1.0 / 2 ** l for l in lvls]
for _ in range(4):
    ops.roi_align_ml_backward(g, [f.shape for f in fl], sc, rois, lv0, 7, 7, 2)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    ops.roi_align_ml_backward(g, [f.shape for f in fl], sc, rois, lv0, 7, 7, 2)
b.record(); torch.cuda.synchronize()
print("mode", mode, "ms", a.elapsed_time(b) / 10)
