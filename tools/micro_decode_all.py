#!/usr/bin/env python
"""Streaming decode of every anchor of one level (vosd_decode_anchors: the full-argsort branch's decode, and the roofline
launch the round-1 review asked for): N images of a P2 map, >= 64 MB per launch, L2 flushed between iterations.
Algorithmic bytes: 16 B of deltas read + 16 B of box written per anchor.

    python tools/micro_decode_all.py [--images 32] [--iters 20] [--out gpurun_out/micro_decode_all.json]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vosdetectron_b200 import ops  # noqa: E402
from vosdetectron_b200.modeling.generate_anchors import fpn_level_anchors  # noqa: E402

PEAK = 6466.8
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:  # noqa: BLE001
    pass


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=32)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    N, A, H, W = a.images, 3, 200, 336
    deltas = (0.2 * torch.randn((N, 4 * A, H, W), device="cuda")).contiguous()
    info = torch.tensor([[800., 1344., 1.0]] * N, device="cuda")
    anchors = fpn_level_anchors(2)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for i in range(a.iters + 3):
        flush.zero_()
        torch.cuda._sleep(2_000_000)       # the host enqueues event, allocation and launch while the GPU waits: no launch gap inside the marks
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        boxes = ops.decode_anchors_cuda(deltas, anchors, 4.0, info)
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(e0.elapsed_time(e1))
    nbytes = deltas.numel() * 4 + boxes.numel() * 4
    ms = float(np.median(ts))
    res = {"images": N, "anchors": N * A * H * W, "bytes": nbytes, "ms_median": ms, "ms_min": float(min(ts)),
           "gbs": nbytes / (ms * 1e-3) / 1e9, "frac_of_measured_hbm": nbytes / (ms * 1e-3) / 1e9 / PEAK,
           "note": "the timed call allocates the output tensor (caching allocator) and launches one kernel"}
    print(json.dumps(res))
    if a.out:
        json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
