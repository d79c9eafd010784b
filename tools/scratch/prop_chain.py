#!/usr/bin/env python
"""The proposal chain of the bench step (10 DAVIS frames: top-k + decode, IoU bitmask, NMS reduce, collect) on its own,
for `ncu` captures:  ncu --set full --import-source on -k regex:topk_decode -c 1 python tools/scratch/prop_chain.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from vosdetectron_b200 import synth  # noqa: E402
from vosdetectron_b200.config import RegionConfig  # noqa: E402
from vosdetectron_b200.pipeline import RegionPipeline  # noqa: E402

B = 10
blob = synth.DAVIS_BLOB if hasattr(synth, "DAVIS_BLOB") else (768, 1344)
r = synth.rpn_outputs(1, blob, B)
rpn = {l: (torch.from_numpy(r[l][0]).cuda(), torch.from_numpy(r[l][1]).cuda()) for l in synth.FPN_LEVELS}
info = torch.tensor([[blob[0], blob[1], 1.6]] * B, dtype=torch.float32, device="cuda")
pipe = RegionPipeline(RegionConfig(), training=False)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    out = pipe.proposals(rpn, info)
torch.cuda.synchronize()
print("ok", int(out["count"].sum()))
