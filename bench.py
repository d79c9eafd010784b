#!/usr/bin/env python
"""Benchmark of the per-frame region pipeline (BASELINE.json metric: frames/s at 1/2/4/8 B200;
RoIAlign achieved HBM GB/s vs peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" = one pass of the hot path over one batch of --frames-per-gpu synthetic frames per GPU
(BASELINE config 3: 480x854 frame -> 768x1344 blob, R-50-FPN RPN outputs for 5 levels, TEST
1000/1000 proposals, box RoIAlign 1000x256x7x7, mask RoIAlign 100x256x14x14, paste of 100
detections).  With 10 frames per GPU, 8 GPUs process the 80-frame DAVIS-shaped clip of config 5
per step; frames are sharded with no data-path collective, only the final all-gather of
detections + bit-packed masks ("scaling": "weak").

  value : frames/s, inputs resident in HBM, CUDA events around exactly K steps, max over ranks.  The step is
          captured once as a CUDA graph (both streams) and the timed region replays it (--no-cuda-graph: eager).
  e2e   : same metric through the host-buffer API (vosdetectron_b200.pipeline.HostPipeline): per step H2D of
          every input from pinned host memory and D2H of the step's results (rois, counts, pasted masks),
          all inside the timed region; uploads, kernels and downloads of consecutive steps overlap on 3 streams.
  roofline : the dominant kernel of the step (box RoIAlign: most algorithmic bytes), algorithmic bytes / its mean
          launch duration from CUDA events on the launching stream -- taken in an eager pass of the same K steps right
          after the graph-replayed timed region (events cannot be read inside a replayed graph), each RoIAlign alone.
  alt_layout : the same device-resident step with the FPN maps in torch.channels_last memory order (TMA-fed
          RoIAlign kernel); never mixed into `value` (N=1 only).
  cpu_baseline : the oracle port of the reference's CPU path on this box's host cores (N=1 only).
`--impl reference` times that CPU path alone (the reference has no CPU RoIAlign at all --
functions/roi_align.py:29-30 raises -- so its RoIAlign leg is the oracle's C restatement).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "region_pipeline_frames_per_sec"
UNIT = "frames/s"
POST = 1000
DETS = 100
K_CLASSES = 81
MASK_M = 28
C_FPN = 256


def workload_name(frames):
    return ("cfg3 per-frame region pipeline: 480x854 frame (blob 768x1344), 5-level RPN top-k/decode/NMS "
            "1000/1000 + collect/distribute, box RoIAlign 1000x256x7x7, mask RoIAlign 100x256x14x14, "
            "paste 100 dets; %d frames/GPU/step" % frames)


# ------------------------------------------------------------------------------------------
# synthetic inputs (host)
# ------------------------------------------------------------------------------------------
def make_host_inputs(frames, seed0, full=True):
    """Per-rank batch of `frames` frames.  RPN outputs from the seeded generator of SURVEY 8d
    (clustered, tie-free); FPN features N(0,1); detections + sigmoid masks."""
    import torch
    from vosdetectron_b200 import synth
    blob = synth.DAVIS_BLOB
    rpn = {l: [[], []] for l in synth.FPN_LEVELS}
    for f in range(frames):
        r = synth.rpn_outputs(seed0 + f, blob, 1)
        for l in synth.FPN_LEVELS:
            rpn[l][0].append(r[l][0])
            rpn[l][1].append(r[l][1])
    rpn = {l: (np.concatenate(v[0]), np.concatenate(v[1])) for l, v in rpn.items()}
    gen = torch.Generator().manual_seed(seed0)
    feats = {l: torch.randn((frames, C_FPN) + synth.level_shape(blob, l), generator=gen).numpy()
             for l in synth.ROI_LEVELS} if full else None
    im_info = np.tile(np.array([[blob[0], blob[1], synth.DAVIS_SCALE]], np.float32), (frames, 1))
    det = [synth.detections(seed0 + 500 + f, DETS, synth.DAVIS_FRAME, MASK_M, K_CLASSES) for f in range(frames)]
    return {"rpn": rpn, "feats": feats, "im_info": im_info,
            "det_boxes": np.stack([d[0] for d in det]), "det_cls": np.stack([d[1] for d in det]),
            "det_masks": np.stack([d[2] for d in det])}


# ------------------------------------------------------------------------------------------
# CPU path (oracle port) -- cpu_baseline leg and --impl reference
# ------------------------------------------------------------------------------------------
_CPU_CACHE = {}


def _cpu_inputs(seed):
    if seed not in _CPU_CACHE:
        import torch
        from vosdetectron_b200 import synth
        blob = synth.DAVIS_BLOB
        gen = torch.Generator().manual_seed(seed)
        _CPU_CACHE.clear()
        _CPU_CACHE[seed] = (
            synth.rpn_outputs(seed, blob, 1),
            {l: torch.randn((1, C_FPN) + synth.level_shape(blob, l), generator=gen).numpy() for l in synth.ROI_LEVELS},
            synth.detections(seed + 500, DETS, synth.DAVIS_FRAME, MASK_M, K_CLASSES))
    return _CPU_CACHE[seed]


def _cpu_one_frame(seed):
    """The reference's per-frame CPU work on one synthetic frame, single thread:
    GenerateProposalsOp x5 levels (NumPy + C NMS), collect, distribute, RoIAlign box + mask
    (C restatement of the CUDA kernel), paste (cv2.resize when present, else its C restatement).
    Returns the seconds spent per stage (input synthesis excluded)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import region_oracle as orc
    from vosdetectron_b200 import synth
    try:
        import cv2
        cv2.setNumThreads(1)
        resize = lambda p, w, h: cv2.resize(p, (w, h))
    except Exception:  # noqa: BLE001
        resize = None
    blob = synth.DAVIS_BLOB
    rpn, feats, (boxes, cls, masks) = _cpu_inputs(seed)
    im_info = np.array([[blob[0], blob[1], synth.DAVIS_SCALE]], np.float32)
    t0 = time.perf_counter()
    rl, pl = [], []
    for l in synth.FPN_LEVELS:
        r, p = orc.generate_proposals(rpn[l][0], rpn[l][1], im_info, orc.fpn_anchors(l), 1. / 2 ** l, 1000, 1000, 0.7, 0)
        rl.append(r)
        pl.append(p)
    rois = orc.collect(rl, pl, POST)
    blobs = orc.distribute(rois)
    t1 = time.perf_counter()
    fl = [feats[l] for l in (5, 4, 3, 2)]
    sc = [1. / 32, 1. / 16, 1. / 8, 1. / 4]
    orc.roi_feature_transform(fl, blobs, 'rois', 7, sc, 2, nthreads=1)
    t2 = time.perf_counter()
    mr = np.concatenate([np.zeros((DETS, 1), np.float32), boxes * np.float32(synth.DAVIS_SCALE)], axis=1)
    orc.roi_feature_transform(fl, orc.distribute(mr), 'rois', 14, sc, 2, nthreads=1)
    t3 = time.perf_counter()
    orc.paste_masks(masks, cls, boxes, synth.DAVIS_FRAME[0], synth.DAVIS_FRAME[1], resize=resize)
    t4 = time.perf_counter()
    return [t1 - t0, t2 - t1, t3 - t2, t4 - t3]


def cpu_path_run(steps, warmup, workers):
    """`steps` timed rounds of `workers` frames, one frame per single-threaded worker process (the
    reference's own sharding model: one process per range of frames, lib/utils/subprocess.py:41-113).
    A round costs the slowest worker's compute time.  Returns (frames/s, ms per round, per-stage
    seconds per frame)."""
    import multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import region_oracle as orc
    orc.build_c()
    seeds = list(range(9000, 9000 + workers))
    ctx = mp.get_context("fork")
    with ctx.Pool(workers) as pool:
        for _ in range(max(1, warmup)):
            pool.map(_cpu_one_frame, seeds, chunksize=1)
        wall, stages = 0.0, []
        for _ in range(steps):
            t0 = time.perf_counter()
            st = pool.map(_cpu_one_frame, seeds, chunksize=1)
            wall += time.perf_counter() - t0
            stages += st
    st = np.mean(np.asarray(stages), axis=0)
    return steps * workers / wall, 1000.0 * wall / steps, {
        "proposals_collect_s": float(st[0]), "roialign_box_s": float(st[1]),
        "roialign_mask_s": float(st[2]), "paste_s": float(st[3])}


def reference_arm(args, rank):
    if rank != 0:
        return
    workers = os.cpu_count() or 1
    steps, warmup = max(1, args.steps), max(1, min(args.warmup, 2))
    fps, ms_step, stages = cpu_path_run(steps, warmup, workers)
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(workers) + " [CPU: one frame per host process per step]"},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": workers, "kind": "port",
                         "sample": "%d steps x %d frames (one per worker process, 1 thread each); "
                                   "NumPy proposal path + C NMS/RoIAlign restatement + cv2.resize paste" % (steps, workers),
                         "stages_per_frame": stages},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------
# helpers for the GPU arm
# ------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0]))
                mx.append(float(p[1]))
            except ValueError:
                continue
            for n, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# dram__bytes_read.sum + dram__bytes_write.sum per launch of the stage's kernel: not measurable inside a timed run
# (ncu replays every kernel ~40 times), so it is read from the committed capture of the same launches
# (profiles/r02_traffic.json, written next to the ncu summary it comes from, with the commit it was taken at).
def _load_traffic():
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
        return {k: {"bytes": v["bytes"], "source": "%s @ %s" % (v["source"], t.get("commit", "?"))}
                for k, v in t.items() if isinstance(v, dict) and "bytes" in v}
    except Exception:  # noqa: BLE001
        return {}


NCU_TRAFFIC = _load_traffic()


def touched_texel_bytes(rois, levels, res, sr, shapes, channels):
    """Exact algorithmic input bytes of a multi-level RoIAlign: 4*C*|unique (frame,level,y,x) texels
    read|, from the RoIs (host replay of the sampling geometry, fp32 like the kernel)."""
    total = 0
    rois = np.asarray(rois, np.float32)
    for li, lvl in enumerate(sorted(shapes)):
        H, W = shapes[lvl]
        sel = np.where(levels == lvl)[0]
        if not len(sel):
            continue
        r = rois[sel]
        scale = np.float32(1.0 / 2 ** lvl)
        frames = r[:, 0].astype(np.int64)
        nf = int(frames.max()) + 1
        touched = np.zeros((nf, H, W), dtype=bool)
        x1, y1 = r[:, 1] * scale, r[:, 2] * scale
        rw = np.maximum(r[:, 3] * scale - x1, 1).astype(np.float32)
        rh = np.maximum(r[:, 4] * scale - y1, 1).astype(np.float32)
        g = np.arange(res * sr, dtype=np.float32)
        p, i = np.floor(g / sr), g % sr

        def taps(start, ext, size):
            b = (ext / np.float32(res))[:, None]
            v = start[:, None] + p[None] * b + (i[None] + np.float32(.5)) * b / np.float32(sr)
            ok = ~((v < -1) | (v > size))
            v = np.maximum(v, 0)
            lo = np.minimum(v.astype(np.int64), size - 1)
            hi = np.minimum(lo + 1, size - 1)
            return lo, hi, ok
        ylo, yhi, yok = taps(y1, rh, H)
        xlo, xhi, xok = taps(x1, rw, W)
        for k in range(len(r)):
            ys = np.unique(np.concatenate([ylo[k][yok[k]], yhi[k][yok[k]]]))
            xs = np.unique(np.concatenate([xlo[k][xok[k]], xhi[k][xok[k]]]))
            if len(ys) and len(xs):
                touched[frames[k]][np.ix_(ys, xs)] = True
        total += int(touched.sum()) * 4 * channels
    return total


# ------------------------------------------------------------------------------------------
# BASELINE config 4: training step (2 frames / GPU): TRAIN 2000/2000 proposals, collect over the minibatch,
# box RoIAlign 1024 x 256 x 7 x 7 and mask RoIAlign 256 x 256 x 14 x 14, forward AND backward
# ------------------------------------------------------------------------------------------
CFG4_IM_INFO = [[800.0, 1333.0, 1.6667], [800.0, 1067.0, 1.25]]     # train-style per-image true sizes
CFG4_BOX_ROIS, CFG4_MASK_ROIS = 1024, 256


def cfg4_train_step(dev, steps, warmup, seed=4000, with_ref=True):
    """One GPU's training step of the region path on a COCO-shaped minibatch of 2 frames.  The label assignment
    between collect and RoIAlign (sampling 512 RoIs per image, fg subset for the mask head) is outside this library:
    the step takes the first 1024 collected RoIs for the box head and the first 256 of them for the mask head.
    Returns ms per step, per-stage ms (CUDA events, a device-side sleep in front of every step keeps launch gaps out),
    the roofline of both RoIAlign backward launches and the unmodified reference kernel (oracle/_ref) timed on the
    same RoIs in the same run, driven like roi_feature_transform / RoIAlignFunction.backward (one launch per level)."""
    import ctypes
    import torch
    from vosdetectron_b200 import ops, synth
    from vosdetectron_b200.config import RegionConfig
    from vosdetectron_b200.pipeline import RegionPipeline
    blob, B = synth.COCO_BLOB, 2
    cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    r = synth.rpn_outputs(seed, blob, B)
    rpn = {l: (cu(r[l][0]), cu(r[l][1])) for l in synth.FPN_LEVELS}
    feats_h = synth.fpn_features(seed + 1, blob, B, synth.ROI_LEVELS, C_FPN)
    feats = {l: cu(feats_h[l]) for l in synth.ROI_LEVELS}
    im_info = cu(np.asarray(CFG4_IM_INFO, np.float32))
    pipe = RegionPipeline(RegionConfig(), training=True)
    fl = [feats[l] for l in synth.ROI_LEVELS]
    sc = [1.0 / 2 ** l for l in synth.ROI_LEVELS]
    shapes = [tuple(f.shape) for f in fl]
    gen = torch.Generator(device="cpu").manual_seed(seed + 2)
    top_box = torch.randn((CFG4_BOX_ROIS, C_FPN, 7, 7), generator=gen).to(dev)
    top_mask = torch.randn((CFG4_MASK_ROIS, C_FPN, 14, 14), generator=gen).to(dev)
    names = ["proposals", "collect_distribute", "sample_rois", "roialign_box_fwd", "roialign_box_bwd", "roialign_mask_fwd",
             "roialign_mask_bwd", "end"]

    def step(mark):
        mark("proposals")
        prop = pipe.proposals(rpn, im_info, images_per_group=B, mark=mark)
        mark("sample_rois")
        rois = prop["rois"].view(-1, 5)[:CFG4_BOX_ROIS].contiguous()
        lv = (ops.distribute_cuda(rois)[0] - 2).to(torch.int32)
        mrois, mlv = rois[:CFG4_MASK_ROIS].contiguous(), lv[:CFG4_MASK_ROIS].contiguous()
        mark("roialign_box_fwd")
        bf = ops.roi_align_ml_forward(fl, sc, rois, lv, 7, 7, 2)
        mark("roialign_box_bwd")
        gb = ops.roi_align_ml_backward(top_box, shapes, sc, rois, lv, 7, 7, 2)
        mark("roialign_mask_fwd")
        mf = ops.roi_align_ml_forward(fl, sc, mrois, mlv, 14, 14, 2)
        mark("roialign_mask_bwd")
        gm = ops.roi_align_ml_backward(top_mask, shapes, sc, mrois, mlv, 14, 14, 2)
        mark("end")
        return rois, lv, bf, gb, mf, gm

    for _ in range(max(3, warmup)):
        out = step(lambda n: None)
    torch.cuda.synchronize()
    # value: K steps back to back
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = step(lambda n: None)
    e1.record()
    torch.cuda.synchronize()
    ms_step = e0.elapsed_time(e1) / steps
    # stage pass
    evs = []

    def mark(n):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        evs[-1].append((n, e))
    for _ in range(steps):
        evs.append([])
        torch.cuda._sleep(6_000_000)
        out = step(mark)
    torch.cuda.synchronize()
    per = {n: [] for n in names[:-1]}
    for ev in evs:
        for (n0, a), (_, b) in zip(ev[:-1], ev[1:]):
            if n0 in per:
                per[n0].append(a.elapsed_time(b))
    stage = {n: float(np.median(v)) for n, v in per.items()}      # median over the steps
    rois, lv = out[0], out[1]
    map_bytes = sum(int(np.prod(s)) * 4 for s in shapes)
    peak = 6650.0
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:  # noqa: BLE001
        pass
    alg_bwd = {"roialign_box_bwd": top_box.numel() * 4 + map_bytes, "roialign_mask_bwd": top_mask.numel() * 4 + map_bytes}
    lv_full = (lv + 2).cpu().numpy()
    lshape = {l: tuple(feats[l].shape[2:]) for l in synth.ROI_LEVELS}
    alg_fwd = {"roialign_box_fwd": top_box.numel() * 4 + 20 * CFG4_BOX_ROIS
               + touched_texel_bytes(rois.cpu().numpy(), lv_full, 7, 2, lshape, C_FPN),
               "roialign_mask_fwd": top_mask.numel() * 4 + 20 * CFG4_MASK_ROIS
               + touched_texel_bytes(rois[:CFG4_MASK_ROIS].cpu().numpy(), lv_full[:CFG4_MASK_ROIS], 14, 2, lshape, C_FPN)}
    roof = {}
    for k, bts in list(alg_bwd.items()) + list(alg_fwd.items()):
        gbs = bts / (stage[k] * 1e-3) / 1e9 if stage[k] > 0 else None
        roof[k] = {"ms": stage[k], "algorithmic_bytes": int(bts), "gbs": gbs, "frac": None if gbs is None else gbs / peak}
    res = {"workload": "cfg4 training step: 2 frames/GPU (blob 800x1344), TRAIN 2000/2000 proposals + collect over the "
                       "minibatch, box RoIAlign 1024x256x7x7 and mask RoIAlign 256x256x14x14 forward + backward "
                       "(label assignment outside the library: first 1024 / 256 collected RoIs)",
           "value": B * 1000.0 / ms_step, "unit": UNIT, "ms_per_step": ms_step, "frames_per_gpu": B,
           "stages_ms": stage, "roofline": roof, "peak": peak,
           "bwd_bytes_rule": "read top_diff once + write every gradient texel once (SURVEY 8d cfg 4)"}
    # ---- the unmodified reference kernel on the same RoIs, driven like the reference's per-level loop
    so = os.path.join(ROOT, "oracle", "_ref", "libref_roialign.so")
    if with_ref and os.path.exists(so):
        lib = ctypes.CDLL(so)
        vp = ctypes.c_void_p
        lib.ROIAlignForwardLaucher.argtypes = [vp, ctypes.c_float] + [ctypes.c_int] * 7 + [vp, vp, vp]
        lib.ROIAlignBackwardLaucher.argtypes = [vp, ctypes.c_float] + [ctypes.c_int] * 8 + [vp, vp, vp]
        st = torch.cuda.current_stream().cuda_stream

        ref_out = {}

        def ulp_hist(mine, theirs):
            """Distance of the default forward to the reference kernel's output in units in the last place of the
            reference value: counts of 0, 1, 2, 3, 4-15, >= 16 ulp (the default kernel re-associates the bilinear sum)."""
            a_ = mine.contiguous().view(torch.int32).to(torch.int64)
            b_ = theirs.contiguous().view(torch.int32).to(torch.int64)
            a_ = torch.where(a_ < 0, -(a_ & 0x7fffffff), a_)              # sign-magnitude -> monotone integer line
            b_ = torch.where(b_ < 0, -(b_ & 0x7fffffff), b_)
            d = (a_ - b_).abs().flatten()
            edges = [0, 1, 2, 3, 4, 16]
            counts = [int(((d >= lo) & (d < hi)).sum()) for lo, hi in zip(edges, edges[1:] + [1 << 62])]
            return {"ulp_0": counts[0], "ulp_1": counts[1], "ulp_2": counts[2], "ulp_3": counts[3], "ulp_4_15": counts[4],
                    "ulp_ge_16": counts[5], "max_abs_err": float((mine - theirs).abs().max()),
                    "max_abs_ref": float(theirs.abs().max())}

        def ref_pass(rr, ll, top, res_):
            idx = [torch.nonzero(ll == i).flatten() for i in range(len(fl))]
            per = [rr[i].contiguous() for i in idx]
            tops = [top[i].contiguous() for i in idx]
            outs, grads = [], []
            a, b, c = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            a.record()
            for i, f in enumerate(fl):                     # forward: one launch per level + cat + restore
                o = torch.zeros((per[i].shape[0], C_FPN, res_, res_), device=dev)
                if per[i].shape[0]:
                    lib.ROIAlignForwardLaucher(f.data_ptr(), sc[i], per[i].shape[0], f.shape[2], f.shape[3], C_FPN, res_, res_, 2,
                                               per[i].data_ptr(), o.data_ptr(), st)
                outs.append(o)
            shuffled = torch.cat(outs)
            restore = torch.argsort(torch.cat(idx))
            ref_out["fwd"] = shuffled[restore]
            b.record()
            for i, f in enumerate(fl):                     # backward: zero-filled maps + one launch per level
                g = torch.zeros_like(f)
                if per[i].shape[0]:
                    lib.ROIAlignBackwardLaucher(tops[i].data_ptr(), sc[i], f.shape[0], per[i].shape[0], f.shape[2], f.shape[3],
                                                C_FPN, res_, res_, 2, per[i].data_ptr(), g.data_ptr(), st)
                grads.append(g)
            c.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b), b.elapsed_time(c)
        ref = {}
        for name, rr, ll, top, res_ in (("box", rois, lv, top_box, 7),
                                        ("mask", rois[:CFG4_MASK_ROIS], lv[:CFG4_MASK_ROIS], top_mask, 14)):
            for _ in range(2):
                ref_pass(rr, ll, top, res_)
            ts = []
            for _ in range(max(5, steps // 2)):
                torch.cuda._sleep(4_000_000)
                ts.append(ref_pass(rr, ll, top, res_))
            ref["roialign_%s_fwd_ms" % name] = float(np.median([t[0] for t in ts]))
            ref["roialign_%s_bwd_ms" % name] = float(np.median([t[1] for t in ts]))
            ref["roialign_%s_fwd_vs_reference" % name] = ulp_hist(ops.roi_align_ml_forward(fl, sc, rr, ll, res_, res_, 2), ref_out["fwd"])
        ref["how"] = ("oracle/_ref/libref_roialign.so = the unmodified roi_align_kernel.cu built for sm_100a; per-level launches + "
                      "cat + restore (forward), zero-filled maps + per-level launches (backward), as the reference drives it")
        res["ref_gpu_kernel"] = ref
    # the same two backward launches with the gradient maps in the OTHER memory order (torch.channels_last, the
    # channels-last kernel): an extra key, not part of the step above
    alt = {}
    for name, top, rr, ll, pooled in (("roialign_box_bwd", top_box, rois, lv, 7),
                                      ("roialign_mask_bwd", top_mask, rois[:CFG4_MASK_ROIS].contiguous(),
                                       lv[:CFG4_MASK_ROIS].contiguous(), 14)):
        ts = []
        for i in range(max(5, steps // 2) + 2):
            torch.cuda._sleep(4_000_000)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            ops.roi_align_ml_backward(top, shapes, sc, rr, ll, pooled, pooled, 2, channels_last=True)
            b.record()
            torch.cuda.synchronize()
            if i >= 2:
                ts.append(a.elapsed_time(b))
        ms = float(np.median(ts))
        alt[name] = {"ms": ms, "gbs": alg_bwd[name] / (ms * 1e-3) / 1e9, "frac": alg_bwd[name] / (ms * 1e-3) / 1e9 / peak}
    res["alt_layout"] = {"features_layout": "channels_last", "roofline": alt}
    return res


def gpu_arm(args, rank, world, local_rank):
    # CPU baseline first (fork-based pool before any CUDA context exists in this process)
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        workers = os.cpu_count() or 1
        fps, _, stages = cpu_path_run(3, 1, workers)
        cpu_baseline = {"value": fps, "unit": UNIT, "cores": workers, "kind": "port",
                        "sample": "3 rounds x %d frames of the same workload, one per worker process (1 thread each)" % workers,
                        "stages_per_frame": stages}

    import torch
    import torch.distributed as dist
    from vosdetectron_b200 import _lib, synth
    from vosdetectron_b200.config import RegionConfig
    from vosdetectron_b200.pipeline import FrameGather, RegionPipeline, STEP_LAUNCHES, all_gather_frames, pack_mask_bits

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # stdout carries exactly ONE JSON line.  Libraries write to file descriptor 1 behind Python's back (NCCL prints its
    # "NCCL version ..." banner there even with NCCL_DEBUG_FILE set), so fd 1 is pointed at stderr for the whole run
    # and the JSON line goes to a private duplicate of the original stdout.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        if args.nccl_channels > 0:
            os.environ["NCCL_MAX_NCHANNELS"] = str(args.nccl_channels)
        dist.init_process_group("nccl", device_id=dev)
    # Pinned host buffers should live on the NUMA node the GPU hangs off: the e2e leg is PCIe-bound and measured
    # 23 ms vs 36 ms per step depending on where the pages landed.  The container hides the topology
    # (/sys/bus/pci/devices/*/numa_node = -1), so the placement is probed: pin a 64 MB buffer from each half of the
    # allowed CPUs, time host->device copies, and keep the process on the faster half if it is clearly faster.
    try:
        allowed = sorted(os.sched_getaffinity(0))
        if len(allowed) >= 4:
            halves = [set(allowed[:len(allowed) // 2]), set(allowed[len(allowed) // 2:])]
            rates = []
            d_probe = torch.empty(64 << 20, dtype=torch.uint8, device=dev)
            for h in halves:
                os.sched_setaffinity(0, h)
                h_probe = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
                h_probe.fill_(1)
                for _ in range(2):
                    d_probe.copy_(h_probe, non_blocking=True)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(8):
                    d_probe.copy_(h_probe, non_blocking=True)
                torch.cuda.synchronize()
                rates.append(8 * 64 / 1024.0 / (time.perf_counter() - t0))
                del h_probe
            best = 0 if rates[0] >= rates[1] else 1
            keep = halves[best] if rates[best] > 1.1 * rates[1 - best] else set(allowed)
            os.sched_setaffinity(0, keep)
            print("rank %d: H2D probe %.1f / %.1f GB/s from the two CPU halves -> %d cpus kept" % (rank, rates[0], rates[1], len(keep)),
                  file=sys.stderr)
            del d_probe
    except Exception:  # noqa: BLE001
        pass
    B = args.frames_per_gpu
    host = make_host_inputs(B, 3000 + 100 * rank)
    pipe = RegionPipeline(RegionConfig())
    frame_hw, im_scale = synth.DAVIS_FRAME, synth.DAVIS_SCALE

    # pinned host buffers (e2e) and device-resident copies (value)
    wc_keep = []

    def pin(a):
        a = np.ascontiguousarray(a)
        if args.wc_pinned:
            # write-combined pinned memory (cudaHostAllocWriteCombined): the device's PCIe reads do not snoop the CPU
            # caches; the host only ever writes these input buffers
            import ctypes
            rt = ctypes.CDLL("libcudart.so")
            ptr = ctypes.c_void_p()
            if rt.cudaHostAlloc(ctypes.byref(ptr), ctypes.c_size_t(max(a.nbytes, 16)), ctypes.c_uint(0x04)) != 0:
                raise RuntimeError("cudaHostAlloc(write-combined) failed")
            buf = (ctypes.c_uint8 * max(a.nbytes, 16)).from_address(ptr.value)
            arr = np.frombuffer(buf, dtype=a.dtype, count=a.size).reshape(a.shape)
            arr[...] = a
            wc_keep.append(buf)
            t = torch.from_numpy(arr)
            assert t.is_pinned()
            return t
        return torch.from_numpy(a).pin_memory()
    h_rpn = {l: (pin(s), pin(d)) for l, (s, d) in host["rpn"].items()}
    h_feats = {l: pin(f) for l, f in host["feats"].items()}
    if args.features_layout == "channels_last":
        h_feats = {l: f.contiguous(memory_format=torch.channels_last).pin_memory() for l, f in h_feats.items()}
    h_info, h_boxes, h_cls = pin(host["im_info"]), pin(host["det_boxes"]), pin(host["det_cls"])
    # the mask head output the reference moves is (D,81,28,28); the API uploads what paste reads
    h_masks = pin(host["det_masks"])
    h2d_tensors = [t for p in h_rpn.values() for t in p] + list(h_feats.values()) + [h_info, h_boxes, h_cls, h_masks]
    h2d_bytes = sum(t.numel() * t.element_size() for t in h2d_tensors)

    def upload():
        return ({l: (s.to(dev, non_blocking=True), d.to(dev, non_blocking=True)) for l, (s, d) in h_rpn.items()},
                h_info.to(dev, non_blocking=True), {l: f.to(dev, non_blocking=True) for l, f in h_feats.items()},
                h_boxes.to(dev, non_blocking=True), h_cls.to(dev, non_blocking=True), h_masks.to(dev, non_blocking=True))

    d_rpn, d_info, d_feats, d_boxes, d_cls, d_masks = upload()
    torch.cuda.synchronize()

    stage_names = ["proposals", "collect_distribute", "join_wait", "roialign_box", "paste_rle", "mask_rois", "roialign_mask", "paste"]
    pipe.overlap = False if args.no_overlap else (True if args.join_overlap else "full")
    # all-gather payload of the sharded clip: the 1-bit-per-pixel masks the paste kernel writes itself (default), or
    # COCO RLE strings from the fused paste -> RLE kernel (--gather-rle: 50x fewer bytes, but the extra kernel costs
    # more than the gather saves: 2 GPUs 1.44 vs 1.34 ms, 4 GPUs 1.51 vs 1.40 ms, 8 GPUs 1.93 vs 1.71 ms per step)
    pipe.packed_masks = False if world == 1 else ("rle" if args.gather_rle else True)
    events = []

    pending = []          # all-gathers in flight (N > 1): they overlap the next step's kernels

    graph = {"g": None, "out": None, "launches": 0, "alt": None, "n": 0}

    fg = {"g": None}      # FrameGather of the bit-packed payload (copy-engine peer pushes; NCCL if unavailable)

    def run_step(mark=None, replay=False):
        direct = False
        if replay and graph["g"] is not None:
            if graph["alt"] is not None and fg["g"] is not None and fg["g"].transport == "ce":
                # two captured graphs with their own output buffers take turns, so the gather can push straight from
                # the step's packed masks (no staging copy between consecutive replays); before a graph is replayed
                # again, the gather that still reads its outputs (two steps back) must be done
                k = graph["n"] % (1 + len(graph["alt"]))
                graph["n"] += 1
                ev = graph["busy"][k]
                if ev is not None:
                    torch.cuda.current_stream().wait_event(ev)
                (graph["g"] if k == 0 else graph["alt"][k - 1][0]).replay()
                out = graph["out"] if k == 0 else graph["alt"][k - 1][1]
                direct = True
            else:
                graph["g"].replay()
                out = graph["out"]
        else:
            out = pipe.step(d_rpn, d_info, d_feats, d_boxes, d_cls, d_masks, frame_hw, im_scale, mark=mark)
        if world > 1:
            dets = torch.cat([d_boxes, d_cls.unsqueeze(-1).float(), torch.ones_like(d_cls).unsqueeze(-1).float()], dim=2)
            if out["masks_packed"] is not None:
                if fg["g"] is None:
                    fg["g"] = FrameGather(dets.shape, dets.dtype, out["masks_packed"].shape, out["masks_packed"].dtype, dev,
                                          transport=args.gather_transport, slots=max(2, args.gather_slots))
                    print("rank %d: frame gather transport = %s%s" % (rank, fg["g"].transport,
                          "" if fg["g"].why is None else " (copy-engine path unavailable: %s)" % fg["g"].why), file=sys.stderr)
                g_ = fg["g"]
                # results are handed over (finish) and their slot given back to the senders (release) one step after
                # the gather started -- not only when the slot is about to be reused: with > 2 slots every sender then
                # holds its acknowledgement whole steps before it needs it
                depth = 1 if (g_.slots > 2 and g_.transport == "ce") else g_.slots - 1
                while len(pending) > depth:
                    s_old = pending.pop(0)[3]
                    g_.finish(s_old)
                    g_.release(s_old)
                payload = out["masks_packed"]
                if g_.transport != "ce" and replay and graph["g"] is not None:
                    payload = payload.clone()                  # NCCL reads it after the next replay has started
                slot_ = g_.start(dets, payload, stage=not direct)
                if direct:
                    graph["busy"][(graph["n"] - 1) % (1 + len(graph["alt"]))] = g_.done_event(slot_)
                pending.append((None, None, [], slot_))
            else:
                while len(pending) > 1:                       # at most two gathers outstanding
                    for w in pending.pop(0)[2]:
                        w.wait()
                r = out["masks_rle"]                      # record = box, class, score, RLE offset / length in the arena
                rec = torch.cat([dets, r["str_offset"].view(B, -1, 1).float(), r["str_len"].view(B, -1, 1).float()], dim=2)
                chars = r["chars"].clone() if (replay and graph["g"] is not None) else r["chars"]
                pending.append(all_gather_frames(rec, chars.view(B, -1), async_op=True))
        return out

    def drain():
        while pending:
            p_ = pending.pop(0)
            for w in p_[2]:
                w.wait()
            if len(p_) > 3:
                fg["g"].finish(p_[3])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        out = run_step()
    drain()
    barrier()

    # ---- value: K steps, device-resident inputs ------------------------------------------
    def mark(name):
        e = torch.cuda.Event(enable_timing=True)
        e.record()              # on the current stream: the mask chain marks its own (second) stream
        events[-1].append((name, e, torch.cuda.current_stream().cuda_stream))

    # The step is a fixed sequence of ~10 kernel launches and ~50 small tensor ops on static shapes: captured once
    # as a CUDA graph (both streams, fork / join included) it costs one launch per step.  That matters at N > 1,
    # where 8 Python processes share the host cores and the eager step becomes launch-bound (1.21 ms on 1 GPU,
    # 1.71 ms on 8).  Any capture failure falls back to the eager step.
    if not args.no_cuda_graph:
        try:
            torch.cuda.synchronize()
            l0 = _lib.launch_count()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                g_out = pipe.step(d_rpn, d_info, d_feats, d_boxes, d_cls, d_masks, frame_hw, im_scale, mark=None)
            nrot = max(2, args.gather_slots)
            graph.update(g=g, out=g_out, launches=_lib.launch_count() - l0, busy=[None] * nrot)
            if world > 1 and not args.gather_rle and args.gather_transport != "nccl":
                # as many captured copies of the step (each with its own outputs) as the gather has slots: the push of
                # step i reads the outputs of copy i % slots, which is replayed again only after that push is done, so
                # a rank that falls behind by a hiccup of up to (slots - 1) steps does not stall its peers
                alts = []
                for _ in range(nrot - 1):
                    g2 = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g2):
                        g2_out = pipe.step(d_rpn, d_info, d_feats, d_boxes, d_cls, d_masks, frame_hw, im_scale, mark=None)
                    alts.append((g2, g2_out))
                graph["alt"] = alts
            for _ in range(3):
                run_step(replay=True)
            drain()
            torch.cuda.synchronize()
        except Exception as exc:  # noqa: BLE001
            print("CUDA graph capture failed (%s: %s); timing the eager step" % (type(exc).__name__, str(exc)[:200]),
                  file=sys.stderr)
            graph.update(g=None, out=None, launches=0)
            torch.cuda.synchronize()
    barrier()

    sampler = ClockSampler(local_rank) if rank == 0 else None     # one nvidia-smi poller per job, not per rank
    if sampler:
        sampler.start()
    launches0 = _lib.launch_count()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t_host = time.perf_counter()
    for _ in range(args.steps):
        events.append([])
        out = run_step(mark, replay=True)
    host_enqueue_ms = (time.perf_counter() - t_host) * 1e3 / args.steps   # host time to enqueue one step (+ its gather)
    drain()               # every gather of the timed steps has completed before the closing event
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count() - launches0
    if graph["g"] is not None:
        launches = graph["launches"] * args.steps      # replayed launches do not pass through the host counter
        # stage durations: CUDA events cannot be read inside a replayed graph, so the same K steps run once more
        # eagerly with the per-stage marks (this pass is not part of `value`)
        # ... with the two chains JOINED before the box RoIAlign, so each RoIAlign is timed alone on the GPU: the
        # roofline figure is a property of the kernel, not of what happens to run beside it
        overlap_mode, pipe.overlap = pipe.overlap, (pipe.overlap if pipe.overlap is False else True)
        host_s = 0.0
        for _ in range(3):      # torch.cuda.graph() emptied the allocator cache: re-populate it outside the timed marks
            torch.cuda.synchronize()
            t_h = time.perf_counter()
            run_step(None)
            host_s = max(host_s, time.perf_counter() - t_h)    # host time to ENQUEUE one eager step
        drain()
        torch.cuda.synchronize()
        events.clear()
        # the eager step is host-bound (~60 tensor ops + 12 launches): a device-side sleep in front of it lets the host
        # enqueue the whole step first, so the marks bracket kernels, not launch gaps.  Its length follows the measured
        # enqueue time (a slower or shared host -- N ranks on one box -- needs a longer head start), at least 2 ms
        sleep_cycles = int(min(max(0.002, 2.0 * host_s), 0.05) * 2.0e9)
        for _ in range(args.steps):
            events.append([])
            torch.cuda._sleep(sleep_cycles)
            out = run_step(mark)
        drain()
        torch.cuda.synchronize()
        pipe.overlap = overlap_mode
        # the proposal chain with the GPU to itself (in the step it runs beside the mask RoIAlign, whose persistent CTAs
        # hold every SM: the "proposals" stage above includes that wait)
        alone = {"proposals": [], "collect_distribute": []}
        for _ in range(max(5, args.steps // 2)):
            torch.cuda._sleep(sleep_cycles)
            evs_ = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            evs_[0].record()
            pipe.proposals(d_rpn, d_info, mark=lambda n: evs_[1].record())
            evs_[2].record()
            torch.cuda.synchronize()
            alone["proposals"].append(evs_[0].elapsed_time(evs_[1]))
            alone["collect_distribute"].append(evs_[1].elapsed_time(evs_[2]))
        stages_alone = {k: float(np.median(v)) for k, v in alone.items()}
    if graph["g"] is None:
        stages_alone = None
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * B * args.steps / (ms / 1000.0)

    # stage duration = distance to the next mark on the SAME stream ("end" / "mask_end" close a chain)
    # (median over the K steps: one step whose launches the host delivered late does not move the figure)
    per_step = {n: [] for n in stage_names}
    for ev in events:
        acc = {n: 0.0 for n in stage_names}
        for sid in {x[2] for x in ev}:
            chain = [x for x in ev if x[2] == sid]
            for (n0, a, _), (_, b, _) in zip(chain[:-1], chain[1:]):
                if n0 in acc:
                    acc[n0] += a.elapsed_time(b)
        for n in stage_names:
            per_step[n].append(acc[n])
    stage_ms = {n: (float(np.median(v)) if v else 0.0) for n, v in per_step.items()}

    # ---- e2e: host buffers through HostPipeline (H2D + step + D2H per batch, 3 streams, 2 slots) ----
    from vosdetectron_b200.pipeline import HostPipeline
    host_batch = {"rpn": h_rpn, "im_info": h_info, "feats": h_feats, "det_boxes": h_boxes, "det_cls": h_cls,
                  "det_masks": h_masks}
    hp = HostPipeline(pipe, frame_hw, im_scale, dev, depth=2)
    for _ in range(3):
        slot = hp.submit(host_batch)
    hp.synchronize()
    res = hp.results(slot)
    d2h_bytes = sum(res[k].numel() * res[k].element_size() for k in ("rois", "roi_count", "masks_packed"))
    # bytes that really cross PCIe per step: every input, but ONE class channel of each detection's (K,M,M) mask
    h2d_bytes_e2e = h2d_bytes - h_masks.numel() * h_masks.element_size() + h_masks.numel() // h_masks.shape[2] * h_masks.element_size()
    barrier()
    e2e_steps = max(4, min(args.steps, 12))
    # The e2e region is PCIe- and host-memory-bound (~0.2 s for 12 steps): a neighbour on the host or the page cache
    # left by the CPU baseline moves one pass by 20 % (453 vs 550 frames/s on one box, minutes apart), so the region is
    # timed three times and the MEDIAN pass is reported (all three are in the line)

    def e2e_pass():
        e0.record()
        hp.s_h2d.wait_event(e0)
        e2e_pending = []
        for _ in range(e2e_steps):
            slot = hp.submit(host_batch)
            if world > 1 and fg["g"] is not None:
                # the sharded clip's exchange is part of the end-to-end step: gather this batch's records + packed masks
                with torch.cuda.stream(hp.s_cmp):
                    while len(e2e_pending) >= fg["g"].slots:
                        fg["g"].finish(e2e_pending.pop(0))
                    o_ = hp.dev_out[slot]
                    dets = torch.cat([hp.dev[slot]["det_boxes"], hp.dev[slot]["det_cls"].unsqueeze(-1).float(),
                                      torch.ones_like(hp.dev[slot]["det_cls"]).unsqueeze(-1).float()], dim=2)
                    e2e_pending.append(fg["g"].start(dets, o_["masks_packed"]))
        if world > 1 and fg["g"] is not None:
            with torch.cuda.stream(hp.s_cmp):
                while e2e_pending:
                    fg["g"].finish(e2e_pending.pop(0))
        for s_ in (hp.s_h2d, hp.s_cmp, hp.s_d2h):
            torch.cuda.current_stream().wait_stream(s_)
        e1.record()
        barrier()
        ms_ = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms_], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_ = float(t.item())
        return ms_

    e2e_all = [e2e_pass() for _ in range(3)]
    e2e_ms = float(np.median(e2e_all))
    e2e_value = world * B * e2e_steps / (e2e_ms / 1000.0)
    # ---- the same device-resident step with the FPN maps in the OTHER memory order (single GPU only): the default
    #      line is the reference's NCHW; torch.channels_last maps take the TMA-fed RoIAlign kernel.  Reported as an
    #      extra key, never mixed into `value`.
    alt = None
    if world == 1:
        alt_layout = "channels_last" if args.features_layout == "nchw" else "nchw"
        keep = d_feats
        if alt_layout == "channels_last":
            d_feats = {l: f.contiguous(memory_format=torch.channels_last) for l, f in keep.items()}
        else:
            d_feats = {l: f.contiguous() for l, f in keep.items()}
        torch.cuda.synchronize()
        for _ in range(max(3, args.warmup)):
            run_step(None)
        torch.cuda.synchronize()
        events.clear()
        e0.record()
        for _ in range(args.steps):
            events.append([])
            run_step(mark)
        e1.record()
        torch.cuda.synchronize()
        alt_ms = e0.elapsed_time(e1)
        alt_stage = {n: 0.0 for n in stage_names}
        for ev in events:
            for sid in {x[2] for x in ev}:
                chain = [x for x in ev if x[2] == sid]
                for (n0, a, _), (_, b, _) in zip(chain[:-1], chain[1:]):
                    if n0 in alt_stage:
                        alt_stage[n0] += a.elapsed_time(b)
        alt = {"features_layout": alt_layout, "value": B * args.steps / (alt_ms / 1000.0), "unit": UNIT,
               "ms_per_step": alt_ms / args.steps,
               "stages_ms": {n: v / args.steps for n, v in alt_stage.items()}}
        d_feats = keep
    train = None
    if world == 1:
        try:
            train = cfg4_train_step(dev, max(5, min(args.steps, 20)), args.warmup)
        except Exception as exc:  # noqa: BLE001
            print("cfg4 training-step leg failed (%s: %s)" % (type(exc).__name__, str(exc)[:300]), file=sys.stderr)
    clocks = sampler.stop() if sampler else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ---------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    shapes = {l: synth.level_shape(synth.DAVIS_BLOB, l) for l in synth.ROI_LEVELS}
    cnt = out["roi_count"].cpu().numpy()
    rois_h = out["rois"].cpu().numpy()
    lvl_h = out["roi_level"].cpu().numpy()
    box_rois = np.concatenate([rois_h[b, :cnt[b]] for b in range(B)])
    box_lvls = np.concatenate([lvl_h[b, :cnt[b]] for b in range(B)])
    bytes_box = (len(box_rois) * C_FPN * 49 * 4 + 20 * len(box_rois)
                 + touched_texel_bytes(box_rois, box_lvls, 7, 2, shapes, C_FPN))
    m_rois, m_lvl = out["mask_rois"].cpu().numpy(), out["mask_level"].cpu().numpy()
    bytes_mask = (len(m_rois) * C_FPN * 196 * 4 + 20 * len(m_rois)
                  + touched_texel_bytes(m_rois, m_lvl, 14, 2, shapes, C_FPN))
    bytes_paste = B * DETS * (frame_hw[0] * frame_hw[1] + MASK_M * MASK_M * 4 + 16)
    n_anchor = sum(3 * h * w for h, w in (synth.level_shape(synth.DAVIS_BLOB, l) for l in synth.FPN_LEVELS))
    bytes_prop = B * (4 * n_anchor + 5 * 1000 * 40)
    alg = {"roialign_box": bytes_box, "roialign_mask": bytes_mask, "paste": bytes_paste, "proposals": bytes_prop}
    # dominant kernel = the one with the most algorithmic bytes (the box RoIAlign; also the longest launch of the
    # serialised ncu list).  With the two chains running concurrently a stage's event-bracketed duration includes the
    # time its CTAs wait behind the other chain's, so "longest stage" is not a property of the kernel any more.
    dom = max(alg, key=lambda k: alg[k])
    ach = alg[dom] / (stage_ms[dom] * 1e-3) / 1e9
    roofline = {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": NCU_TRAFFIC.get(dom, {}).get("bytes"), "traffic_source": NCU_TRAFFIC.get(dom, {}).get("source"),
                "algorithmic_bytes_per_launch": alg[dom], "ms_per_launch": stage_ms[dom],
                "peak_source": peak_src,
                "proposal_chain_alone_ms": stages_alone,
                "stages": {k: {"ms": stage_ms[k], "algorithmic_bytes": alg.get(k),
                               "gbs": (alg[k] / (stage_ms[k] * 1e-3) / 1e9) if k in alg and stage_ms[k] > 0 else None}
                           for k in stage_ms}}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(B), "frames_per_gpu": B, "global_frames_per_step": B * world,
                   "l2": "inputs larger than L2 (%.2f GB touched per step)" % (h2d_bytes / 1e9),
                   "streams": ("1" if args.no_overlap else
                               "2: proposal chain on a high-priority stream beside mask RoIAlign + paste, joined before the box RoIAlign"
                               if args.join_overlap else
                               "2: proposals -> collect -> box RoIAlign on a high-priority stream beside mask RoIAlign + paste "
                               "(stage durations of concurrent kernels overlap)"),
                   "features_layout": args.features_layout,
                   "cuda_graph": ("the timed region replays a CUDA graph of the step; stage durations from an eager pass "
                                  "of the same steps with the chains joined before the box RoIAlign (each kernel alone)"
                                  if graph["g"] is not None else "off (eager step)"),
                   "parallelism": "frame-sharded x%d%s" % (world, (", all-gather of dets + %s per step%s" % (
                       "COCO RLE strings (fused paste -> RLE kernel)" if args.gather_rle else "bit-packed masks",
                       "" if fg["g"] is None else " [%s]" % ("copy-engine peer pushes over NVLink (symmetric memory), no SMs"
                                                             if fg["g"].transport == "ce" else "NCCL all_gather_into_tensor"))) if world > 1 else "")},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes_e2e, "d2h_bytes_per_step": d2h_bytes,
                "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
                "passes_ms_per_step": [m_ / e2e_steps for m_ in e2e_all], "reported": "median of 3 passes of `steps` steps",
                "what": "HostPipeline: pinned host inputs -> H2D (class channel of every detection mask gathered on the host "
                        "inside the timed region) -> step -> D2H of rois, counts and the 1-bit-per-pixel pasted masks"},
        "gpu_launches": int(launches),
        "host_enqueue_ms_per_step": host_enqueue_ms,
        "launches_per_step_expected": STEP_LAUNCHES + (1 if pipe.packed_masks == "rle" else 0),
        "roofline": roofline,
    }
    if cpu_baseline is not None:
        line["cpu_baseline"] = cpu_baseline
    if train is not None:
        line["train_step_cfg4"] = train
    if alt is not None:
        # achieved bytes/s of the box RoIAlign with the maps in the other memory order (same algorithmic bytes)
        if alt["stages_ms"].get("roialign_box"):
            alt["roialign_box_gbs"] = alg["roialign_box"] / (alt["stages_ms"]["roialign_box"] * 1e-3) / 1e9
            alt["roialign_box_frac"] = alt["roialign_box_gbs"] / peak
        line["alt_layout"] = alt
    sys.stdout.flush()
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    os.close(json_fd)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames-per-gpu", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-overlap", action="store_true", help="run the mask chain behind the box chain on one stream")
    ap.add_argument("--no-cuda-graph", action="store_true", help="time the eager step instead of a CUDA-graph replay of it")
    ap.add_argument("--nccl-channels", type=int, default=0,
                    help="N > 1 experiment: cap NCCL at this many channels (NCCL_MAX_NCHANNELS) so the all-gather holds fewer SMs")
    ap.add_argument("--wc-pinned", action="store_true",
                    help="experiment: write-combined pinned host input buffers for the e2e leg")
    ap.add_argument("--gather-transport", default="auto", choices=["auto", "ce", "nccl"],
                    help="N > 1: exchange of the packed masks -- copy-engine peer pushes over symmetric memory (auto: when "
                         "available) or NCCL all_gather_into_tensor")
    ap.add_argument("--max-connections", type=int, default=0,
                    help="experiment: CUDA_DEVICE_MAX_CONNECTIONS for the run (hardware work queues the streams map to)")
    ap.add_argument("--gather-slots", type=int, default=2,
                    help="N > 1: rotating slots of the frame gather = captured copies of the step (>= 2; measured at 8 GPUs: "
                         "2 -> 74.2k, 3 -> 68.7k, 4 -> 66.5k frames/s, profiles/README.md)")
    ap.add_argument("--gather-rle", action="store_true",
                    help="N > 1: all-gather COCO RLE strings (fused paste -> RLE kernel) instead of 1-bit-per-pixel masks")
    ap.add_argument("--join-overlap", action="store_true",
                    help="join the two streams before the box RoIAlign (it then runs alone) instead of letting it run "
                         "on the proposal stream beside mask RoIAlign + paste (default)")
    ap.add_argument("--features-layout", default="nchw", choices=["nchw", "channels_last"],
                    help="memory order of the synthetic FPN maps: the reference's NCHW (default) or torch.channels_last "
                         "(N,H,W,C), which routes RoIAlign through the TMA-fed channels-last kernel")
    args = ap.parse_args()
    if args.max_connections > 0:
        os.environ["CUDA_DEVICE_MAX_CONNECTIONS"] = str(args.max_connections)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    gpu_arm(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
