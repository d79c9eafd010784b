import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/oracle')
import numpy as np, torch
from vosdetectron_b200 import ops, synth
import region_oracle as orc
lvls = synth.ROI_LEVELS
blob = synth.DAVIS_BLOB if hasattr(synth, 'DAVIS_BLOB') else (768, 1344)
N = 10
feats = synth.fpn_features(2000, blob, N, lvls, 256)
fl = [torch.from_numpy(feats[l]).cuda() for l in lvls]
sc = [1.0 / 2 ** l for l in lvls]
def run(rois_np, tag):
    rois = torch.from_numpy(rois_np).cuda()
    level = torch.from_numpy(orc.map_rois_to_fpn_levels(rois_np[:, 1:5]).astype(np.int32) - 2).cuda()
    for res in (7, 14):
        for _ in range(3): ops.roi_align_ml_forward(fl, sc, rois, level, res, res, 2)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): ops.roi_align_ml_forward(fl, sc, rois, level, res, res, 2)
        b.record(); torch.cuda.synchronize()
        print(tag, "res", res, "rois", len(rois_np), "ms", a.elapsed_time(b) / 10)
allr = synth.random_rois(5, 30000, blob, N)
lv = orc.map_rois_to_fpn_levels(allr[:, 1:5])
wtex = (allr[:, 3] - allr[:, 1]) / 2 ** lv
small = allr[wtex <= 26][:10000]
run(allr[:10000], "all")
run(small, "w<=26tex")
big = allr[wtex > 30][:2000]
run(big, "w>30tex")
