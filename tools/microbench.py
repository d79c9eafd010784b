#!/usr/bin/env python
"""Per-kernel micro-benchmarks at the BASELINE.json config sizes (run on the GPU box).

    python tools/microbench.py [--iters 30] [--out gpurun_out/micro.json]

For every kernel: CUDA-event time with the L2 flushed between iterations (a 512 MB scratch write)
and "hot" (no flush), algorithmic bytes, achieved GB/s and the fraction of the
measured HBM peak.  RoIAlign is timed next to the REFERENCE kernel (oracle/_ref/libref_roialign.so:
the unmodified roi_align_kernel.cu built for sm_100a) driven exactly like the reference's
roi_feature_transform (one launch per level + cat + index_select).  This script is measurement
tooling, not the product path; it loads oracle/_ref only as the comparison arm.
"""
import argparse
import os
os.environ.setdefault("VOSD_B200_TEST_HOOKS", "1")
import ctypes
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

from vosdetectron_b200 import ops, synth  # noqa: E402
from vosdetectron_b200.modeling.generate_anchors import fpn_level_anchors  # noqa: E402
from bench import touched_texel_bytes  # noqa: E402

PEAK = 6466.8
try:
    PEAK = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:  # noqa: BLE001
    pass


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


class Timer:
    def __init__(self, iters):
        self.iters = iters
        self.scratch = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")

    def run(self, fn, flush):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(self.iters):
            if flush:
                self.scratch.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        return float(np.median(ts)), float(ts[0])


def report(out, name, timer, fn, alg_bytes, extra=None):
    cold, cold_min = timer.run(fn, True)
    hot, hot_min = timer.run(fn, False)
    rec = {"kernel": name, "ms_cold_median": cold, "ms_cold_min": cold_min, "ms_hot_median": hot,
           "algorithmic_bytes": int(alg_bytes), "gbs_cold": alg_bytes / cold / 1e6, "gbs_hot": alg_bytes / hot / 1e6,
           "frac_of_measured_hbm_cold": alg_bytes / cold / 1e6 / PEAK, "frac_of_measured_hbm_hot": alg_bytes / hot / 1e6 / PEAK}
    if extra:
        rec.update(extra)
    out.append(rec)
    print(json.dumps(rec), flush=True)
    return rec


def load_ref():
    so = os.path.join(ROOT, "oracle", "_ref", "libref_roialign.so")
    if not os.path.exists(so):
        return None
    lib = ctypes.CDLL(so)
    vp = ctypes.c_void_p
    lib.ROIAlignForwardLaucher.argtypes = [vp, ctypes.c_float] + [ctypes.c_int] * 7 + [vp, vp, vp]
    lib.ROIAlignBackwardLaucher.argtypes = [vp, ctypes.c_float] + [ctypes.c_int] * 8 + [vp, vp, vp]
    return lib


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "micro.json"))
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    timer = Timer(args.iters)
    ref = load_ref()
    out = []
    want = lambda k: not args.only or k in args.only
    stream = lambda: torch.cuda.current_stream().cuda_stream
    lvls = synth.ROI_LEVELS
    scales = [1.0 / 2 ** l for l in lvls]

    # ---------------- cfg 2: multi-level RoIAlign fwd, 1000 RoIs x 256 ch, COCO blob ----------------
    for N, R, tag in (((1, 1000, "cfg2"), (2, 1024, "cfg4box"), (2, 256, "cfg4mask")) if want("fwd") or want("bwd") else ()):
        feats = synth.fpn_features(2000, synth.COCO_BLOB, N, lvls, 256)
        fl = [cu(feats[l]) for l in lvls]
        shapes = {l: feats[l].shape[2:] for l in lvls}
        rois_h = synth.random_rois(2001, R, synth.COCO_BLOB, N)
        rois = cu(rois_h)
        level, lc, order, restore = ops.distribute_cuda(rois)
        lv0 = (level - 2).to(torch.int32)
        level_h = level.cpu().numpy()
        order_h = order.cpu().numpy().astype(np.int64)
        per_level = [cu(rois_h[order_h[level_h[order_h] == l]]) for l in lvls]
        restore_l = restore.long()
        for res in ((7, 14) if tag == "cfg2" else ((7,) if tag == "cfg4box" else (14,))):
            touched = touched_texel_bytes(rois_h, level_h, res, 2, shapes, 256)
            out_bytes = R * 256 * res * res * 4
            alg = out_bytes + touched + 20 * R
            upper = out_bytes + sum(f.numel() * 4 for f in fl)
            if want("fwd"):
                report(out, "%s_roialign_fwd_%dx%d_vosd" % (tag, res, res), timer,
                       lambda: ops.roi_align_ml_forward(fl, scales, rois, lv0, res, res, 2), alg,
                       {"rois": R, "images": N, "upper_bound_bytes_whole_maps": int(upper)})
                if ref is not None:
                    def ref_fwd():
                        parts = []
                        for i, l in enumerate(lvls):
                            r = per_level[i]
                            if r.shape[0]:
                                o = torch.zeros((r.shape[0], 256, res, res), device="cuda")    # features.new(...).zero_()
                                ref.ROIAlignForwardLaucher(fl[i].data_ptr(), scales[i], r.shape[0], fl[i].shape[2],
                                                           fl[i].shape[3], 256, res, res, 2, r.data_ptr(), o.data_ptr(), stream())
                                parts.append(o)
                        return torch.cat(parts)[restore_l]
                    report(out, "%s_roialign_fwd_%dx%d_REFERENCE_kernel" % (tag, res, res), timer, ref_fwd, alg,
                           {"rois": R, "note": "unmodified reference .cu for sm_100a: 4 launches + zero-fill + cat + index_select"})
            # ---------------- backward (cfg 4 accounting: read top_diff + write every grad texel) -------
            if want("bwd"):
                g = torch.randn((R, 256, res, res), device="cuda")
                alg_b = out_bytes + sum(f.numel() * 4 for f in fl)
                report(out, "%s_roialign_bwd_%dx%d_vosd" % (tag, res, res), timer,
                       lambda: ops.roi_align_ml_backward(g, [f.shape for f in fl], scales, rois, lv0, res, res, 2), alg_b,
                       {"rois": R, "images": N})
                if ref is not None:
                    g_sh = g[order.long()].contiguous()
                    offs = np.concatenate([[0], np.cumsum([p.shape[0] for p in per_level])])

                    def ref_bwd():
                        outs = []
                        for i, l in enumerate(lvls):
                            r = per_level[i]
                            gi = torch.zeros_like(fl[i])                                       # rois.new(...).zero_()
                            if r.shape[0]:
                                ref.ROIAlignBackwardLaucher(g_sh[offs[i]:offs[i + 1]].data_ptr(), scales[i], N, r.shape[0],
                                                            fl[i].shape[2], fl[i].shape[3], 256, res, res, 2, r.data_ptr(),
                                                            gi.data_ptr(), stream())
                            outs.append(gi)
                        return outs
                    report(out, "%s_roialign_bwd_%dx%d_REFERENCE_kernel" % (tag, res, res), timer, ref_bwd, alg_b, {"rois": R})
        del fl, feats

    # ---------------- proposals: cfg 1 (one COCO image) and a 10-frame batch ----------------
    if want("prop"):
        for N in (1, 10):
            rpn = {}
            for f in range(N):
                r = synth.rpn_outputs(1000 + f, synth.COCO_BLOB, 1)
                for l in r:
                    rpn.setdefault(l, [[], []])
                    rpn[l][0].append(r[l][0])
                    rpn[l][1].append(r[l][1])
            inputs = [(cu(np.concatenate(rpn[l][0])), cu(np.concatenate(rpn[l][1])), fpn_level_anchors(l), float(2 ** l))
                      for l in synth.FPN_LEVELS]
            info = cu(np.tile(np.array([[800, 1344, 1.6667]], np.float32), (N, 1)))
            n_anchor = sum(int(i[0].shape[1] * i[0].shape[2] * i[0].shape[3]) for i in inputs)
            for pre, post in ((2000, 1000), (1000, 1000)):
                alg = N * (4 * n_anchor + 5 * pre * 40)
                report(out, "cfg1_generate_proposals_%dimg_pre%d_post%d" % (N, pre, post), timer,
                       lambda: ops.generate_proposals_cuda(inputs, info, pre, post, 0.7, 0.0), alg,
                       {"images": N, "anchors_per_image": n_anchor, "iou_pairs_per_image_upper": 5 * pre * (pre - 1) // 2})
            rois, probs, count = ops.generate_proposals_cuda(inputs, info, 2000, 1000, 0.7, 0.0)
            report(out, "collect_distribute_%dimg" % N, timer,
                   lambda: ops.collect_distribute_cuda(rois, probs, count, 1000, 1), N * 5 * 1000 * 24, {"images": N})
            # streaming decode of every anchor (P2 level)
            sc, dl, anc, st = inputs[0]
            alg = dl.numel() * 4 * 2
            report(out, "decode_all_anchors_P2_%dimg" % N, timer, lambda: ops.decode_anchors_cuda(dl, anc, st, info), alg,
                   {"anchors": int(dl.numel() // 4)})

    # ---------------- NMS alone ----------------
    if want("nms"):
        for n in (1000, 2000):
            d = cu(synth.clustered_dets(n, n))
            report(out, "nms_%d_clustered" % n, timer, lambda: ops.nms_cuda(d, 0.7), 28 * n, {"iou_pairs": n * (n - 1) // 2})

    # ---------------- paste: cfg 3 (100 dets, 480x854) and a 10-frame batch ----------------
    if want("paste"):
        for B in (1, 10):
            det = [synth.detections(3000 + b, 100) for b in range(B)]
            boxes = cu(np.concatenate([d[0] for d in det]))
            cls = cu(np.concatenate([d[1] for d in det]))
            masks = cu(np.concatenate([d[2] for d in det]))
            alg = B * 100 * (480 * 854 + 28 * 28 * 4 + 16)
            report(out, "paste_%dframes_100dets_480x854" % B, timer,
                   lambda: ops.paste_masks_cuda(masks, cls, boxes, 480, 854, 0.5), alg, {"dets": B * 100})
            report(out, "paste_dense_plus_packed_%dframes_100dets_480x854" % B, timer,
                   lambda: ops.paste_masks_packed_cuda(masks, cls, boxes, 480, 854, 0.5), alg + alg // 8, {"dets": B * 100})
            report(out, "paste_packed_only_%dframes_100dets_480x854" % B, timer,
                   lambda: ops.paste_masks_packed_cuda(masks, cls, boxes, 480, 854, 0.5, want_dense=False),
                   B * 100 * (480 * 854 // 8 + 28 * 28 * 4 + 16), {"dets": B * 100})
            # fused paste -> COCO RLE: algorithmic bytes = masks in + runs and string out (measured once)
            r = ops.paste_rle_cuda(masks, cls, boxes, 480, 854, 0.5)
            used = r["cursors"].cpu().numpy()
            rc, rs = int(used[0]), int(used[1])
            cap_r, cap_s = r["runs"].numel(), r["chars"].numel()
            report(out, "paste_rle_%dframes_100dets_480x854" % B, timer,
                   lambda: ops.paste_rle_cuda(masks, cls, boxes, 480, 854, 0.5, run_capacity=cap_r, str_capacity=cap_s),
                   B * 100 * (28 * 28 * 4 + 16 + 28) + 4 * rc + rs,
                   {"dets": B * 100, "runs": rc, "string_bytes": rs, "dense_equivalent_bytes": B * 100 * 480 * 854})

    # ---------------- FlowAlign: hidden states of the 5 FPN levels of DAVIS frames ----------------
    if want("flow"):
        from vosdetectron_b200 import _lib
        so = os.path.join(ROOT, "oracle", "_ref", "libref_flowalign.so")
        rlib = ctypes.CDLL(so) if os.path.exists(so) else None
        if rlib is not None:
            vp = ctypes.c_void_p
            rlib.FlowAlignForward.argtypes = [ctypes.c_int] * 4 + [vp, vp, vp, vp]
            rlib.FlowAlignBackward.argtypes = [ctypes.c_int] * 4 + [vp, vp, vp, vp, vp, vp]
        for B in (1, 4):
            shapes = [synth.level_shape(synth.DAVIS_BLOB, l) for l in synth.FPN_LEVELS]
            feats = [torch.randn((B, 256, h, w), device="cuda") for h, w in shapes]
            flows = [cu(synth.flow_field(4000 + i, B, h, w, "smooth", 2.0)) for i, (h, w) in enumerate(shapes)]
            grads = [torch.randn_like(f) for f in feats]
            elems = sum(f.numel() for f in feats)
            flow_bytes = sum(f.numel() for f in flows) * 4
            # forward: read every feature texel once + the flow, write every output element
            alg_f = 2 * elems * 4 + flow_bytes
            # backward: read topdiff + features + flow, write both gradients once (zero or sum)
            alg_b = 3 * elems * 4 + 2 * flow_bytes
            for mode, tag in ((0, ""), (2, "_aluvariant"), (1, "_fp32variant")):
                old = _lib.load().vosd_debug_flow_align_fast(mode)
                report(out, "flowalign_fwd%s_5lvl_%dframes" % (tag, B), timer,
                       lambda: ops.flow_align_ml_forward(feats, flows), alg_f, {"elements": elems})
                if mode != 1:
                    report(out, "flowalign_bwd%s_5lvl_%dframes" % (tag, B), timer,
                           lambda: ops.flow_align_ml_backward(grads, feats, flows), alg_b, {"elements": elems})
                _lib.load().vosd_debug_flow_align_fast(old)
            if rlib is not None:
                tops = [torch.empty_like(f) for f in feats]
                gfs = [torch.empty_like(f) for f in feats]
                gfls = [torch.empty_like(f) for f in flows]

                def ref_f():
                    for f, fl_, t in zip(feats, flows, tops):
                        rlib.FlowAlignForward(B, f.shape[2], f.shape[3], 256, f.data_ptr(), fl_.data_ptr(), t.data_ptr(), stream())

                def ref_b():
                    for g, f, fl_, gf, gfl in zip(grads, feats, flows, gfs, gfls):
                        gf.zero_(), gfl.zero_()
                        rlib.FlowAlignBackward(B, f.shape[2], f.shape[3], 256, g.data_ptr(), f.data_ptr(), fl_.data_ptr(),
                                               gf.data_ptr(), gfl.data_ptr(), stream())
                report(out, "REFERENCE_flowalign_fwd_5lvl_%dframes" % B, timer, ref_f, alg_f, {"elements": elems})
                report(out, "REFERENCE_flowalign_bwd_5lvl_%dframes" % B, timer, ref_b, alg_b, {"elements": elems})


    # ---------------- yardsticks: what the memory system gives plain fill / copy kernels of the same sizes ----------------
    if want("yard"):
        fill = torch.empty(410 << 20, dtype=torch.uint8, device="cuda")           # the paste output of 10 frames
        report(out, "YARDSTICK_memset_410MB (cudaMemsetAsync via torch.zero_)", timer, lambda: fill.zero_(), fill.numel())
        a = torch.empty(353 << 20, dtype=torch.uint8, device="cuda")              # FlowAlign forward: read 353 MB, write 353 MB
        b = torch.empty_like(a)
        report(out, "YARDSTICK_copy_353MB_read+353MB_write (torch.copy_)", timer, lambda: b.copy_(a), 2 * a.numel())
        if want("flow") or True:
            from vosdetectron_b200 import _lib as _l
            shapes = [synth.level_shape(synth.DAVIS_BLOB, l) for l in synth.FPN_LEVELS]
            feats = [torch.randn((4, 256, h, w), device="cuda") for h, w in shapes]
            flows = [cu(synth.flow_field(4000 + i, 4, h, w, "smooth", 2.0)) for i, (h, w) in enumerate(shapes)]
            grads = [torch.randn_like(f) for f in feats]
            elems = sum(f.numel() for f in feats)
            report(out, "flowalign_bwd_feature_grad_only_5lvl_4frames", timer,
                   lambda: ops.flow_align_ml_backward(grads, feats, flows, want_flow_grad=False),
                   2 * elems * 4 + sum(f.numel() for f in flows) * 4, {"elements": elems})


    # ---------------- channels-last RoIAlign forward (TMA-fed) next to the NCHW kernels: the bench launch (10 frames) ----------------
    if want("nhwc"):
        B = 10
        feats = synth.fpn_features(7000, synth.DAVIS_BLOB, B, lvls, 256)
        fl = [cu(feats[l]) for l in lvls]
        fl_cl = [f.contiguous(memory_format=torch.channels_last) for f in fl]
        shapes = {l: feats[l].shape[2:] for l in lvls}
        for R_per, res in ((1000, 7), (100, 14)):
            rois_h = synth.random_rois(7001, R_per * B, synth.DAVIS_BLOB, B)
            rois = cu(rois_h)
            level, lc, order, restore = ops.distribute_cuda(rois)
            lv0 = (level - 2).to(torch.int32)
            touched = touched_texel_bytes(rois_h, level.cpu().numpy(), res, 2, shapes, 256)
            alg = R_per * B * 256 * res * res * 4 + touched + 20 * R_per * B
            a = report(out, "bench_roialign_fwd_%dx%d_%drois_nchw" % (res, res, R_per * B), timer,
                       lambda: ops.roi_align_ml_forward(fl, scales, rois, lv0, res, res, 2), alg, {"rois": R_per * B})
            b = report(out, "bench_roialign_fwd_%dx%d_%drois_channels_last_tma" % (res, res, R_per * B), timer,
                       lambda: ops.roi_align_ml_forward(fl_cl, scales, rois, lv0, res, res, 2), alg, {"rois": R_per * B})
            print("speedup channels-last / nchw: %.2fx" % (a["ms_cold_median"] / b["ms_cold_median"]), flush=True)


    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump({"peak_gbs_measured": PEAK, "gpu": torch.cuda.get_device_name(0), "results": out}, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
