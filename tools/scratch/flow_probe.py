"""ncu target: one multi-level FlowAlign forward + backward over 4 DAVIS frames (5 FPN levels, 256 ch)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from vosdetectron_b200 import ops, synth

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
shapes = [synth.level_shape(synth.DAVIS_BLOB, l) for l in synth.FPN_LEVELS]
feats = [torch.randn((B, 256, h, w), device="cuda") for h, w in shapes]
flows = [torch.from_numpy(synth.flow_field(4000 + i, B, h, w, "smooth", 2.0)).cuda() for i, (h, w) in enumerate(shapes)]
grads = [torch.randn_like(f) for f in feats]
for _ in range(2):
    ops.flow_align_ml_forward(feats, flows)
    ops.flow_align_ml_backward(grads, feats, flows)
torch.cuda.synchronize()
print("ok")
