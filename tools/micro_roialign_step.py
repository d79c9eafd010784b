#!/usr/bin/env python
"""Micro-benchmark of the two RoIAlign launches of the bench step (cfg 3: 10 DAVIS frames, 1000 proposal RoIs per
frame 7x7, 100 detections per frame 14x14) on the RoIs the step itself produces, L2 flushed between iterations.

    python tools/micro_roialign_step.py [--frames 10] [--iters 20] [--out gpurun_out/micro_step.json] [--old]

Reports ms, algorithmic bytes (output + unique texels touched + rois), GB/s, fraction of the measured HBM peak and
max |err| against the workspace-free separable kernel (--old times that one)."""
import argparse
import ctypes
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vosdetectron_b200 import _lib, ops, synth  # noqa: E402
from vosdetectron_b200.pipeline import RegionPipeline  # noqa: E402
import bench  # noqa: E402

PEAK = 6466.8
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:  # noqa: BLE001
    pass


def old_forward(fl, sc, rois, lv, res):
    R, C = rois.shape[0], fl[0].shape[1]
    out = torch.empty((R, C, res, res), dtype=torch.float32, device=rois.device)
    ptrs, hs, ws, scs = ops._level_arrays(fl, sc)
    _lib.call("vosd_roialign_ml_fwd", ptrs, hs, ws, scs, len(fl), C, res, res, 2, R, ops._ptr(rois), ops._ptr(lv), None,
              ops._ptr(out), ops._stream())
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=10)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "micro_step.json"))
    ap.add_argument("--old", action="store_true")
    ap.add_argument("--once", action="store_true", help="one launch of each (for ncu)")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    host = bench.make_host_inputs(args.frames, 1234)
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    rpn = {l: (cu(v[0]), cu(v[1])) for l, v in host["rpn"].items()}
    feats = {l: cu(v) for l, v in host["feats"].items()}
    im_info = cu(host["im_info"])
    pipe = RegionPipeline()
    prop = pipe.proposals(rpn, im_info)
    B = args.frames
    rois = prop["rois"].view(-1, 5).contiguous()
    level = prop["level"].view(-1).clamp(2, 5)
    lv = (level - 2).to(torch.int32)
    det_boxes = cu(host["det_boxes"])
    frame_idx = torch.arange(B, device=dev, dtype=torch.float32).view(B, 1, 1).expand(B, det_boxes.shape[1], 1)
    mrois = torch.cat([frame_idx, det_boxes * synth.DAVIS_SCALE], dim=2).view(-1, 5).contiguous()
    mlevel = ops.distribute_cuda(mrois, 2, 5, 224, 4)[0]
    mlv = (mlevel - 2).to(torch.int32)
    fl = [feats[l] for l in synth.ROI_LEVELS]
    sc = [1.0 / 2 ** l for l in synth.ROI_LEVELS]
    shapes = {l: tuple(feats[l].shape[2:]) for l in synth.ROI_LEVELS}
    scratch = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    res_out = []
    for name, r, l, lvl_full, res in (("box7", rois, lv, level, 7), ("mask14", mrois, mlv, mlevel, 14)):
        new = lambda: ops.roi_align_ml_forward(fl, sc, r, l, res, res, 2)
        old = lambda: old_forward(fl, sc, r, l, res)
        fn = old if args.old else new
        if args.once:
            fn()
            torch.cuda.synchronize()
            continue
        a, b = new(), old()
        err = float((a - b).abs().max())
        tol = 1e-5 * b.abs() + 1e-6 * float(b.abs().max())
        nbad = int(((a - b).abs() > tol).sum())
        for _ in range(3):
            fn()
        ts = []
        for _ in range(args.iters):
            scratch.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        rh = r.cpu().numpy()
        touched = bench.touched_texel_bytes(rh, lvl_full.cpu().numpy(), res, 2, shapes, fl[0].shape[1])
        alg = r.shape[0] * fl[0].shape[1] * res * res * 4 + touched + 20 * r.shape[0]
        med = float(np.median(ts))
        rec = {"launch": name, "kernel": "old_sep" if args.old else "row_window", "rois": int(r.shape[0]), "ms_median": med,
               "ms_min": ts[0], "algorithmic_bytes": int(alg), "gbs": alg / med / 1e6,
               "frac_of_measured_hbm": alg / med / 1e6 / PEAK, "max_abs_err_vs_sep": err, "out_of_tolerance": nbad}
        print(json.dumps(rec), flush=True)
        res_out.append(rec)
    if not args.once:
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        json.dump(res_out, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
