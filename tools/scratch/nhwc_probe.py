"""ncu target: the bench-sized channels-last RoIAlign forward (10 frames x 1000 RoIs x 256 ch x 7x7)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from vosdetectron_b200 import ops, synth
B = 10
lvls = synth.ROI_LEVELS
feats = synth.fpn_features(7000, synth.DAVIS_BLOB, B, lvls, 256)
fl = [torch.from_numpy(feats[l]).cuda().contiguous(memory_format=torch.channels_last) for l in lvls]
rois = torch.from_numpy(synth.random_rois(7001, 1000 * B, synth.DAVIS_BLOB, B)).cuda()
level, _, _, _ = ops.distribute_cuda(rois)
lv0 = (level - 2).to(torch.int32)
scales = [1.0 / 2 ** l for l in lvls]
for _ in range(3):
    ops.roi_align_ml_forward(fl, scales, rois, lv0, 7, 7, 2)
torch.cuda.synchronize()
print("ok")
