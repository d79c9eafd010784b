"""Seeded synthetic inputs for the region pipeline.

NumPy only; shared by tests/, bench.py and __graft_entry__.smoke() so that the
CUDA path, the oracle and the reference all see byte-identical inputs.
All randomness comes from ``numpy.random.RandomState(seed)``.
"""
import math

import numpy as np

FPN_LEVELS = (2, 3, 4, 5, 6)          # RPN levels (lib/core/config.py FPN.RPN_MIN/MAX_LEVEL)
ROI_LEVELS = (2, 3, 4, 5)             # RoI levels (FPN.ROI_MIN/MAX_LEVEL)
NUM_ANCHORS = 3                       # one size x 3 aspect ratios per level (FPN.py:343-350)
FPN_DIM = 256

COCO_BLOB = (800, 1344)               # 800x1333 frame padded to stride 32
DAVIS_BLOB = (768, 1344)              # 480x854 frame at TEST.SCALE 800 -> 749x1333 -> padded
DAVIS_FRAME = (480, 854)
DAVIS_SCALE = 1333.0 / 854.0            # min(800/480, 1333/854) (utils/blob.py:37-61)


def level_shape(blob_hw, lvl):
    s = 2 ** lvl
    return int(math.ceil(blob_hw[0] / s)), int(math.ceil(blob_hw[1] / s))


def rpn_outputs(seed, blob_hw=COCO_BLOB, num_images=1, levels=FPN_LEVELS, A=NUM_ANCHORS,
                clustered=True):
    """Per level: tie-free scores (N,A,H,W) and deltas (N,4A,H,W), fp32.

    Scores are a random permutation of (i+0.5)/n (n < 2**24, so distinct after
    rounding to fp32; image k is offset by 0.05*k/n); with ``clustered`` the permutation is biased by
    a smooth low-frequency field so that the top-k really cluster and NMS
    suppresses.  Deltas: N(0,0.5) for dx,dy, N(0,0.25) for dw,dh, 0.1 % of dw/dh
    set to +-5 to exercise BBOX_XFORM_CLIP and the max(w,1) floor."""
    rs = np.random.RandomState(seed)
    out = {}
    for lvl in levels:
        H, W = level_shape(blob_hw, lvl)
        n = A * H * W
        sc = np.empty((num_images, A, H, W), dtype=np.float32)
        for i in range(num_images):
            if clustered:
                yy, xx = np.mgrid[0:H, 0:W].astype(np.float64)
                field = np.zeros((H, W))
                for _ in range(6):
                    cy, cx = rs.uniform(0, H), rs.uniform(0, W)
                    sg = rs.uniform(0.05, 0.25) * max(H, W)
                    field += np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * sg * sg))
                key = field[None] + 0.35 * rs.uniform(size=(A, H, W))
                rank = np.empty(n, dtype=np.int64)
                rank[np.argsort(key.ravel(), kind='stable')] = np.arange(n)
            else:
                rank = rs.permutation(n)
            # + 0.05*i keeps images of one minibatch tie-free too (collect() sorts across images)
            sc[i] = ((rank + 0.5 + 0.05 * (i % 10)) / n).astype(np.float32).reshape(A, H, W)
        d = np.empty((num_images, 4 * A, H, W), dtype=np.float32)
        d[:, 0::4] = rs.normal(0, 0.5, size=(num_images, A, H, W))
        d[:, 1::4] = rs.normal(0, 0.5, size=(num_images, A, H, W))
        d[:, 2::4] = rs.normal(0, 0.25, size=(num_images, A, H, W))
        d[:, 3::4] = rs.normal(0, 0.25, size=(num_images, A, H, W))
        big = rs.uniform(size=(num_images, 2 * A, H, W)) < 1e-3
        sign = np.where(rs.uniform(size=big.shape) < 0.5, -5.0, 5.0).astype(np.float32)
        wh = np.concatenate([d[:, 2::4], d[:, 3::4]], axis=1)
        wh[big] = sign[big]
        d[:, 2::4], d[:, 3::4] = wh[:, :A], wh[:, A:]
        out[lvl] = (sc, d)
    return out


def fpn_features(seed, blob_hw=COCO_BLOB, num_images=1, levels=ROI_LEVELS, C=FPN_DIM):
    """{lvl: (N,C,H_l,W_l) fp32 ~ N(0,1)}."""
    rs = np.random.RandomState(seed)
    return {lvl: rs.standard_normal((num_images, C) + level_shape(blob_hw, lvl)).astype(np.float32)
            for lvl in levels}


def random_rois(seed, R, im_hw=COCO_BLOB, num_images=1, smin=16.0, smax=800.0):
    """(R,5) fp32 [batch,x1,y1,x2,y2]: centre uniform in the image, sqrt(area)
    log-uniform in [smin,smax] px, aspect log-uniform in [1/3,3], clipped to the image."""
    rs = np.random.RandomState(seed)
    H, W = im_hw
    s = np.exp(rs.uniform(math.log(smin), math.log(smax), R))
    ar = np.exp(rs.uniform(math.log(1 / 3.0), math.log(3.0), R))
    w, h = s * np.sqrt(ar), s / np.sqrt(ar)
    cx, cy = rs.uniform(0, W, R), rs.uniform(0, H, R)
    x1 = np.clip(cx - w / 2, 0, W - 1)
    x2 = np.clip(cx + w / 2, 0, W - 1)
    y1 = np.clip(cy - h / 2, 0, H - 1)
    y2 = np.clip(cy + h / 2, 0, H - 1)
    b = rs.randint(0, num_images, R)
    return np.stack([b, x1, y1, x2, y2], axis=1).astype(np.float32)


def edge_rois(im_hw=COCO_BLOB):
    """Degenerate / out-of-map RoIs (separate edge-case suite)."""
    H, W = im_hw
    r = [
        [0, 10, 10, 10, 10],               # zero area
        [0, 50, 60, 40, 30],               # negative extent
        [0, W - 40, H - 30, W + 60, H + 80],   # hangs over right/bottom edge
        [0, -50, -40, 30, 20],             # hangs over top/left
        [0, 0, 0, W - 1, H - 1],           # whole image
        [0, W - 1, H - 1, W - 1, H - 1],   # touches the last texel
        [0, -500, -500, -400, -450],       # entirely outside
        [0, 100.25, 37.75, 100.75, 38.5],  # sub-texel
    ]
    return np.asarray(r, dtype=np.float32)


def clustered_dets(seed, n, im_hw=COCO_BLOB, n_centres=40):
    """(n,5) fp32 [x1,y1,x2,y2,score] clustered boxes with distinct scores, unsorted."""
    rs = np.random.RandomState(seed)
    H, W = im_hw
    c = rs.randint(0, n_centres, n)
    ctr = np.stack([rs.uniform(0, W, n_centres), rs.uniform(0, H, n_centres)], 1)[c]
    size = np.exp(rs.uniform(math.log(24), math.log(400), n_centres))[c]
    cx = ctr[:, 0] + rs.normal(0, 0.15, n) * size
    cy = ctr[:, 1] + rs.normal(0, 0.15, n) * size
    w = size * np.exp(rs.normal(0, 0.2, n))
    h = size * np.exp(rs.normal(0, 0.2, n))
    x1 = np.clip(cx - w / 2, 0, W - 1); x2 = np.clip(cx + w / 2, 0, W - 1)
    y1 = np.clip(cy - h / 2, 0, H - 1); y2 = np.clip(cy + h / 2, 0, H - 1)
    sc = ((rs.permutation(n) + 0.5) / n)
    return np.stack([x1, y1, x2, y2, sc], 1).astype(np.float32)


def detections(seed, R=100, frame_hw=DAVIS_FRAME, M=28, K=81):
    """Boxes (R,4) fp32 in original-frame coords, classes ascending (class-major
    order of segm_results), masks (R,K,M,M) fp32 = sigmoid of smooth noise."""
    rs = np.random.RandomState(seed)
    boxes = random_rois(seed + 7, R, frame_hw, 1, smin=12.0, smax=400.0)[:, 1:5].copy()
    cls = np.sort(rs.randint(1, K, R)).astype(np.int32)
    base = rs.standard_normal((R, 1, M // 4 + 2, M // 4 + 2)).astype(np.float32)
    up = np.kron(base, np.ones((1, 1, 4, 4), dtype=np.float32))[:, :, 2:2 + M, 2:2 + M]
    noise = 0.3 * rs.standard_normal((R, K, M, M)).astype(np.float32)
    masks = (1.0 / (1.0 + np.exp(-(2.5 * up + noise)))).astype(np.float32)
    return boxes, cls, masks


def box_head_outputs(seed, R=1000, K=81, im_hw=COCO_BLOB, fg_frac=0.25):
    """Synthetic Fast R-CNN head outputs for one image: proposals (R,4) in image pixels (clustered, so per-class
    NMS really suppresses), softmax scores (R,K) with ``fg_frac`` of the rows confident in one of ~12 classes
    (distinct values: tie-free), box deltas (R,4K) ~ N(0, 0.5) / N(0, 0.25) as the RPN deltas."""
    rs = np.random.RandomState(seed)
    props = clustered_dets(seed + 1, R, im_hw, n_centres=max(8, R // 25))[:, :4].copy()
    logits = rs.standard_normal((R, K)).astype(np.float64) * 0.5
    logits[:, 0] += 3.0
    fg = rs.uniform(size=R) < fg_frac
    cls = rs.choice(np.arange(1, K), size=min(12, K - 1), replace=False)
    which = cls[rs.randint(0, len(cls), R)]
    logits[np.arange(R)[fg], which[fg]] += rs.uniform(3.0, 8.0, int(fg.sum()))
    e = np.exp(logits - logits.max(1, keepdims=True))
    scores = (e / e.sum(1, keepdims=True)).astype(np.float32)
    # make every score distinct (tie-free ordering and image_thresh)
    scores = (scores * (1.0 + 1e-4 * rs.permutation(R * K).reshape(R, K) / (R * K))).astype(np.float32)
    d = np.empty((R, 4 * K), dtype=np.float32)
    d[:, 0::4] = rs.normal(0, 0.5, (R, K)); d[:, 1::4] = rs.normal(0, 0.5, (R, K))
    d[:, 2::4] = rs.normal(0, 0.25, (R, K)); d[:, 3::4] = rs.normal(0, 0.25, (R, K))
    return props.astype(np.float32), scores, d


def flow_field(seed, N, H, W, kind="smooth", magnitude=2.0):
    """Synthetic optical flow (N,2,H,W) fp32 at feature-map resolution, plane 0 = x, plane 1 = y.

    ``smooth``: a low-frequency displacement field (a few Fourier modes, |flow| <~ magnitude) plus 2 % pixel
    noise -- what a down-sampled DAVIS flow looks like; ``noise``: i.i.d. N(0, magnitude^2), every pixel its own
    geometry; ``zero`` / ``shift``: constant fields (identity / integer translation by (+2, -1))."""
    rs = np.random.RandomState(seed)
    if kind == "zero":
        return np.zeros((N, 2, H, W), dtype=np.float32)
    if kind == "shift":
        f = np.zeros((N, 2, H, W), dtype=np.float32)
        f[:, 0] = 2.0
        f[:, 1] = -1.0
        return f
    if kind == "noise":
        return (rs.standard_normal((N, 2, H, W)) * magnitude).astype(np.float32)
    ys, xs = np.meshgrid(np.arange(H) / max(H, 1), np.arange(W) / max(W, 1), indexing="ij")
    f = np.zeros((N, 2, H, W), dtype=np.float64)
    for n in range(N):
        for p in range(2):
            for _ in range(4):
                ky, kx = rs.randint(0, 3, size=2)
                ph = rs.uniform(0, 2 * np.pi, size=2)
                f[n, p] += rs.uniform(-1, 1) * np.sin(2 * np.pi * ky * ys + ph[0]) * np.cos(2 * np.pi * kx * xs + ph[1])
            f[n, p] *= magnitude / 2.0
    f += rs.standard_normal(f.shape) * 0.02 * magnitude
    return f.astype(np.float32)
