"""CPU: the C-ABI library loads and exports every symbol include/vosd_b200.h declares, argument
validation returns status codes without touching a GPU, and the host-side logic (config adapter,
anchors, RLE, frame sharding, bit packing) behaves."""
import ctypes
import os
import re
import types

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from vosdetectron_b200 import build, _lib
    build.build()
    return _lib.load()


def test_every_header_symbol_is_exported(lib):
    from vosdetectron_b200 import _lib
    header = open(os.path.join(ROOT, "include", "vosd_b200.h")).read()
    declared = set(re.findall(r"VOSD_API [\w \*]+?(vosd_\w+)\(", header))
    assert len(declared) >= 19
    for name in declared:
        assert hasattr(lib, name), name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert b"sm_100a" in lib.vosd_version()
    assert lib.vosd_status_string(-3) == b"unsupported size (compiled-in limit)"


def test_argument_validation_without_gpu(lib):
    from vosdetectron_b200._lib import RpnLevel
    assert lib.vosd_nms(None, -1, 0.5, None, None, None, 0, None) == -1            # bad shape
    assert lib.vosd_nms(None, 5, 0.5, None, None, None, 0, None) == -2             # null num_keep
    assert lib.vosd_nms_workspace_bytes(2000) >= 2000 * 32 * 8
    assert lib.vosd_paste_masks(None, None, None, 3, 1, 28, 0, 10, 0.5, None, None, None) == -1
    assert lib.vosd_paste_masks(None, None, None, 3, 1, 28, 10, 10, 0.5, None, None, None) == -2
    assert lib.vosd_roialign_fwd(None, 0.25, 4, 10, 10, 8, 7, 7, 2, None, None, None) == -2
    assert lib.vosd_roialign_fwd(ctypes.c_void_p(256), 0.25, 4, 0, 10, 8, 7, 7, 2, None, None, None) == -1
    assert lib.vosd_collect_distribute(None, None, None, 5, 4, 100, 3, 100, 2, 5, 224.0, 4,
                                       None, None, None, None, None, None, None, 0, None) == -1   # 4 % 3 != 0
    lv = (RpnLevel * 2)()
    for i, (h, w) in enumerate(((200, 336), (13, 21))):
        lv[i].height, lv[i].width, lv[i].num_anchors = h, w, 3
    assert lib.vosd_proposals_capacity(lv, 2, 2000, 1000) == 1000
    assert lib.vosd_proposals_capacity(lv, 2, 2000, 0) == 2000
    one = ctypes.cast(ctypes.byref(lv[1]), ctypes.POINTER(RpnLevel))
    assert lib.vosd_proposals_capacity(one, 1, 2000, 1000) == 819                  # 13*21*3 < pre
    assert lib.vosd_proposals_capacity(lv, 2, 0, 1000) == 1000                     # full sort of P2: streamed kernel
    assert lib.vosd_proposals_capacity(lv, 2, 0, 0) == 200 * 336 * 3
    assert lib.vosd_generate_proposals_workspace_bytes(lv, 2, 1, 0, 1000) > 0
    lv[0].num_anchors = 17
    assert lib.vosd_proposals_capacity(lv, 2, 2000, 1000) == -3
    assert lib.vosd_generate_proposals_workspace_bytes(one, 1, 10, 2000, 1000) > 0


def test_missing_library_fails_loudly(monkeypatch):
    from vosdetectron_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libvosd_b200.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()


def test_cpu_tensors_are_rejected():
    from vosdetectron_b200 import ops
    from vosdetectron_b200.modeling.roi_xfrom.roi_align.functions.roi_align import RoIAlignFunction
    with pytest.raises(NotImplementedError):
        RoIAlignFunction(7, 7, 0.25, 2)(torch.zeros(1, 2, 8, 8), torch.zeros(1, 5))
    with pytest.raises(NotImplementedError):
        ops.nms_cuda(torch.zeros(4, 5), 0.5)
    with pytest.raises(NotImplementedError):
        ops.paste_masks_cuda(torch.zeros(1, 1, 28, 28), None, torch.zeros(1, 4), 10, 10)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "vosdetectron_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(d, f)).read()
                assert "region_oracle" not in src and "liboracle" not in src and "ref_harness" not in src, f


def test_anchors_match_golden(golden):
    from vosdetectron_b200.modeling.generate_anchors import generate_anchors, fpn_level_anchors
    g = golden("anchors")
    assert np.array_equal(generate_anchors(16, (128, 256, 512), (0.5, 1, 2)), g["kat_stride16"])
    for lvl in range(2, 7):
        a = fpn_level_anchors(lvl)
        assert a.dtype == np.float64 and np.array_equal(a, g["fpn%d" % lvl])


def test_config_adapter():
    from vosdetectron_b200.config import RegionConfig, set_cfg, get_cfg
    ns = types.SimpleNamespace
    rpn = lambda pre, post: ns(RPN_PRE_NMS_TOP_N=pre, RPN_POST_NMS_TOP_N=post, RPN_NMS_THRESH=0.7, RPN_MIN_SIZE=0,
                               SCORE_THRESH=0.05, NMS=0.5, DETECTIONS_PER_IM=100, SOFT_NMS=ns(ENABLED=False),
                               BBOX_VOTE=ns(ENABLED=False))
    cfg = ns(TRAIN=rpn(2000, 2000), TEST=rpn(1000, 1000),
             FPN=ns(RPN_MIN_LEVEL=2, RPN_MAX_LEVEL=6, ROI_MIN_LEVEL=2, ROI_MAX_LEVEL=5, ROI_CANONICAL_SCALE=224,
                    ROI_CANONICAL_LEVEL=4, RPN_COLLECT_SCALE=1, RPN_ANCHOR_START_SIZE=32, RPN_ASPECT_RATIOS=(0.5, 1, 2)),
             BBOX_XFORM_CLIP=np.log(1000. / 16.), MODEL=ns(NUM_CLASSES=81, BBOX_REG_WEIGHTS=(10., 10., 5., 5.)),
             MRCNN=ns(RESOLUTION=28, THRESH_BINARIZE=0.5, CLS_SPECIFIC_MASK=True))
    rc = RegionConfig.from_cfg(cfg)
    assert rc.mode(True).pre_nms_topN == 2000 and rc.mode(False).post_nms_topN == 1000
    assert rc.collect_post_topN(False) == 1000 and rc.num_classes == 81
    assert rc.test_nms == 0.5 and rc.test_detections_per_im == 100 and rc.test_num_det_per_class == 0
    assert rc.bbox_reg_weights == (10., 10., 5., 5.) and not rc.test_soft_nms
    old = get_cfg()
    try:
        assert set_cfg(cfg).train.post_nms_topN == 2000 and get_cfg().test.pre_nms_topN == 1000
    finally:
        set_cfg(old)


def test_rle_encoder_round_trip():
    from vosdetectron_b200.core.test import rle_encode
    rs = np.random.RandomState(0)
    for shape in ((5, 7), (48, 85), (1, 1)):
        m = (rs.uniform(size=shape) > 0.6).astype(np.uint8)
        m[1:3] = 1 if shape[0] > 3 else m[1:3]
        rle = rle_encode(m)
        # decode (pycocotools rleFrString + rleDecode)
        cnts, p, s = [], 0, rle["counts"]
        while p < len(s):
            x, k, more = 0, 0, True
            while more:
                c = ord(s[p]) - 48
                x |= (c & 0x1f) << (5 * k)
                more = bool(c & 0x20)
                p += 1
                k += 1
                if not more and (c & 0x10):
                    x |= -1 << (5 * k)
            if len(cnts) > 2:
                x += cnts[-2]
            cnts.append(x)
        flat = np.concatenate([np.full(c, i & 1, np.uint8) for i, c in enumerate(cnts)]) if cnts else np.zeros(0, np.uint8)
        assert np.array_equal(flat.reshape(shape[::-1]).T, m)
    assert rle_encode(np.zeros((5, 7), np.uint8))["counts"] == "S1"      # 35 zeros, pycocotools' string


def test_frame_sharding_and_bit_packing():
    from vosdetectron_b200.pipeline import shard_frames, pack_mask_bits, unpack_mask_bits
    parts = [shard_frames(80, 8, r) for r in range(8)]
    assert all(len(p) == 10 for p in parts) and np.array_equal(np.concatenate(parts), np.arange(80))
    parts = [shard_frames(82, 4, r) for r in range(4)]
    assert [len(p) for p in parts] == [21, 21, 20, 20]                   # np.array_split convention
    m = (torch.rand(3, 2, 37, 53) > 0.5).to(torch.uint8)
    p = pack_mask_bits(m)
    assert p.shape == (3, 2, (37 * 53 + 7) // 8)
    assert torch.equal(unpack_mask_bits(p, 37, 53), m)


def test_reference_aliases_install():
    import sys
    import vosdetectron_b200 as v
    names = v.install_reference_aliases()
    assert 'modeling.roi_xfrom.roi_align.functions.roi_align' in names
    mod = sys.modules['modeling.roi_xfrom.roi_align.functions.roi_align']
    assert hasattr(mod, 'RoIAlignFunction')
    for n in names:
        sys.modules.pop(n, None)


def test_rle_string_parser_matches_the_oracle(orc):
    """core.test.rle_counts_from_string (host half of mask_util.decode in the nms_with_mask_iou mirror) against the
    oracle's rleFrString restatement and its encoder, bytes or str input, negative differences included."""
    from vosdetectron_b200.core import test as core_test
    rs = np.random.RandomState(9)
    for _ in range(50):
        h, w = rs.randint(1, 40), rs.randint(1, 40)
        m = (rs.rand(h, w) > rs.rand()).astype(np.uint8)
        enc = orc.rle_encode(m)
        counts = core_test.rle_counts_from_string(enc['counts'])
        assert counts == orc.rle_from_string(enc['counts']) == orc.rle_counts_fast(m)
        assert core_test.rle_counts_from_string(enc['counts'].encode('ascii')) == counts
        assert core_test.rle_encode(m)['counts'] == enc['counts']
    big = [5, 1000, 3, 2, 70000, 1]
    assert core_test.rle_counts_from_string(orc.rle_to_string(big)) == big


def test_argument_validation_of_the_next_row_entry_points(lib):
    """Status codes of the round's new entry points, reachable without a GPU (validation precedes any CUDA call)."""
    vp, ip, fp = ctypes.c_void_p, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_float)
    p = vp(256)
    # FlowAlign: empty problems are no-ops, negative extents / missing pointers are errors, the int index limit holds
    assert lib.vosd_flow_align_fwd(0, 8, 8, 4, None, None, None, None) == 0
    assert lib.vosd_flow_align_fwd(2, 8, 8, 0, None, None, None, None) == 0
    assert lib.vosd_flow_align_fwd(-1, 8, 8, 4, p, p, p, None) == -1
    assert lib.vosd_flow_align_fwd(2, 8, 8, 4, None, p, p, None) == -2
    assert lib.vosd_flow_align_fwd(64, 1024, 1024, 32, p, p, p, None) == -3           # N*C*H*W >= 2^31
    assert lib.vosd_flow_align_bwd(2, 8, 8, 4, p, p, p, None, p, 0, None) == -2        # bottomdiff missing
    assert lib.vosd_debug_flow_align_fast(0) in (0, 1, 2)
    # mask-IoU suppression
    assert lib.vosd_mask_iou_nms(None, -1, 64, None, 0.5, None, None, None, 0, None) == -1
    assert lib.vosd_mask_iou_nms(p, 4, 64, None, 0.5, p, None, None, 0, None) == -2    # num_keep missing
    assert lib.vosd_mask_iou_nms(p, 4, 62, None, 0.5, p, p, p, 1 << 20, None) == -2    # not whole 32-bit words
    assert lib.vosd_mask_iou_nms(p, 5000, 64, None, 0.5, p, p, p, 1 << 30, None) == -3  # > 2048 masks
    assert lib.vosd_mask_iou_nms(p, 8, 64, None, 0.5, p, p, None, 0, None) == -4       # workspace missing
    assert lib.vosd_mask_iou_nms_workspace_bytes(100) >= 100 * 4 + 100 * 2 * 8
    assert lib.vosd_rle_to_bits(None, None, None, 0, 100, None, 0, None) == 0
    assert lib.vosd_rle_to_bits(p, p, p, 3, 100, p, 20000, None) == -3                 # > 12000 runs per mask
    # label-assignment pieces
    assert lib.vosd_bbox_overlaps(None, 0, None, 5, None, None, None, None) == 0
    assert lib.vosd_bbox_overlaps(None, 4, p, 5, None, None, None, None) == -2
    assert lib.vosd_bbox_overlaps(vp(260), 4, p, 5, None, None, None, None) == -2      # not 16-byte aligned
    w = (ctypes.c_float * 4)(10, 10, 5, 5)
    assert lib.vosd_bbox_targets(p, p, p, 0, 81, 0, w, p, p, None, None) == 0
    assert lib.vosd_bbox_targets(p, p, p, 4, 0, 0, w, p, p, None, None) == -1
    assert lib.vosd_bbox_targets(p, p, None, 4, 81, 0, w, p, p, None, None) == -2
    # channels-last RoIAlign: heads outside the supported set are refused so that the caller takes the NCHW entry point
    h, wd = (ctypes.c_int * 1)(50), (ctypes.c_int * 1)(84)
    sc = (ctypes.c_float * 1)(0.0625)
    data = (vp * 1)(256)
    args = lambda C, pw, sr: (data, h, wd, sc, 1, 1, C, 7, pw, sr, 10, p, None, None, p, None)
    assert lib.vosd_roialign_ml_fwd_nhwc(*args(256, 7, 0)) == -3                        # adaptive grid
    assert lib.vosd_roialign_ml_fwd_nhwc(*args(48, 7, 2)) == -3                         # C % 32 != 0
    assert lib.vosd_roialign_ml_fwd_nhwc(*args(256, 6, 2)) == -3                        # pooled width not 7 / 14 / 28
    assert lib.vosd_roialign_ml_fwd_nhwc(*args(0, 7, 2)) == -1


def test_channels_last_dispatch_predicates():
    import torch
    from vosdetectron_b200 import ops
    x = torch.zeros((2, 64, 5, 7))
    assert not ops._is_channels_last(x) and not ops._is_channels_last(x.contiguous(memory_format=torch.channels_last))  # CPU
    assert ops._nhwc_supported(256, 7, 2) and ops._nhwc_supported(64, 28, 2)
    assert not ops._nhwc_supported(48, 7, 2) and not ops._nhwc_supported(256, 7, 0) and not ops._nhwc_supported(256, 6, 2)


def test_cffi_level_mirrors_import_and_reject_cpu_tensors():
    """`_ext.roi_align` / `_ext.flow_align`: the entry points the reference's Function classes call, importable under
    the reference's module names, refusing CPU tensors like everything else here."""
    import sys
    import torch
    import vosdetectron_b200
    names = vosdetectron_b200.install_reference_aliases()
    assert 'modeling.roi_xfrom.roi_align._ext.roi_align' in sys.modules and 'vos_model.flow_align._ext.flow_align' in sys.modules
    # (the parent packages `modeling`, `vos_model` come from the reference tree when it is on sys.path)
    ra = sys.modules['modeling.roi_xfrom.roi_align._ext.roi_align']
    fa = sys.modules['vos_model.flow_align._ext.flow_align']
    assert ra.__all__ == ["roi_align_forward_cuda", "roi_align_backward_cuda"]
    assert fa.__all__ == ["flow_align_forward_cuda", "flow_align_backward_cuda"]
    f, r = torch.zeros((1, 4, 8, 8)), torch.zeros((2, 5))
    with pytest.raises(NotImplementedError):
        ra.roi_align_forward_cuda(7, 7, 0.25, 2, f, r, torch.zeros((2, 4, 7, 7)))
    with pytest.raises(NotImplementedError):
        fa.flow_align_forward_cuda(f, torch.zeros((1, 2, 8, 8)), torch.zeros_like(f))
    assert isinstance(names, list)


def test_rpn_label_host_pieces(orc, golden):
    """Host side of the RPN label mirror (roi_data/rpn.py, roi_data/data_utils.py): the field of anchors equals the
    oracle's restatement of get_field_of_anchors (and is cached), the blob names are the reference's (= the keys of the
    golden produced by the unmodified add_rpn_blobs), the config adapter reads the TRAIN.RPN_* keys, and without a GPU
    the mirror raises instead of computing anything."""
    import types
    from vosdetectron_b200.config import RegionConfig
    from vosdetectron_b200.roi_data import data_utils, rpn
    c = RegionConfig()
    c.train_max_size = 384
    foas = rpn._fields(c)
    assert [f.field_size for f in foas] == [96, 48, 24, 12, 6] and all(f.num_cell_anchors == 3 for f in foas)
    for l, f in zip(range(2, 7), foas):
        ref, A, field = orc.field_of_anchors(2. ** l, (32 * 2. ** (l - 2),), (0.5, 1, 2), 384)
        assert (A, field) == (f.num_cell_anchors, f.field_size)
        assert f.field_of_anchors.dtype == np.float32 and np.array_equal(f.field_of_anchors, ref)
    assert rpn._fields(c)[0] is foas[0]                                    # cached like the reference's thread-local cache
    c1 = RegionConfig()
    c1.fpn_on = c1.multilevel_rpn = False
    c1.train_max_size, c1.rpn_sizes = 384, (32, 64, 128, 256)
    (single,) = rpn._fields(c1)
    ref, A, field = orc.field_of_anchors(16, (32, 64, 128, 256), (0.5, 1, 2), 384)
    assert A == 12 and field == 24 and np.array_equal(single.field_of_anchors, ref)
    g = golden("rpn_labels")
    for tag, cfg_ in (("fpn_", c), ("single_", c1)):
        names = set(rpn.get_rpn_blob_names(cfg=cfg_)) - {"roidb"}
        assert names == {k[len(tag):] for k in g.files if k.startswith(tag + "rpn_") or k == tag + "im_info"}
    assert rpn.get_rpn_blob_names(is_training=False, cfg=c) == ["im_info"]
    fake = types.SimpleNamespace(
        TRAIN=types.SimpleNamespace(RPN_PRE_NMS_TOP_N=2000, RPN_POST_NMS_TOP_N=2000, RPN_NMS_THRESH=0.7, RPN_MIN_SIZE=0,
                                    RPN_POSITIVE_OVERLAP=0.6, RPN_NEGATIVE_OVERLAP=0.2, RPN_FG_FRACTION=0.25,
                                    RPN_BATCH_SIZE_PER_IM=128, RPN_STRADDLE_THRESH=-1, MAX_SIZE=1000),
        TEST=types.SimpleNamespace(RPN_PRE_NMS_TOP_N=1000, RPN_POST_NMS_TOP_N=1000, RPN_NMS_THRESH=0.7, RPN_MIN_SIZE=0,
                                   SCORE_THRESH=0.05, NMS=0.5, DETECTIONS_PER_IM=100,
                                   SOFT_NMS=types.SimpleNamespace(ENABLED=False), BBOX_VOTE=types.SimpleNamespace(ENABLED=False)),
        FPN=types.SimpleNamespace(RPN_MIN_LEVEL=2, RPN_MAX_LEVEL=6, ROI_MIN_LEVEL=2, ROI_MAX_LEVEL=5, ROI_CANONICAL_SCALE=224,
                                  ROI_CANONICAL_LEVEL=4, RPN_COLLECT_SCALE=1, RPN_ANCHOR_START_SIZE=32, RPN_ASPECT_RATIOS=(0.5, 1, 2),
                                  FPN_ON=True, MULTILEVEL_RPN=False, COARSEST_STRIDE=64),
        RPN=types.SimpleNamespace(STRIDE=8, SIZES=(16, 32), ASPECT_RATIOS=(1,)),
        BBOX_XFORM_CLIP=4.135, MODEL=types.SimpleNamespace(NUM_CLASSES=81, BBOX_REG_WEIGHTS=(10., 10., 5., 5.)),
        MRCNN=types.SimpleNamespace(RESOLUTION=28, THRESH_BINARIZE=0.5, CLS_SPECIFIC_MASK=True))
    r = RegionConfig.from_cfg(fake)
    assert (r.train_rpn_positive_overlap, r.train_rpn_negative_overlap, r.train_rpn_fg_fraction) == (0.6, 0.2, 0.25)
    assert (r.train_rpn_batch_size_per_im, r.train_rpn_straddle_thresh, r.train_max_size) == (128, -1.0, 1000)
    assert (r.fpn_on, r.multilevel_rpn, r.fpn_coarsest_stride) == (True, False, 64)
    assert (r.rpn_stride, r.rpn_sizes, r.rpn_single_aspect_ratios) == (8, (16, 32), (1,))
    if not torch.cuda.is_available():
        blobs = {k: [] for k in rpn.get_rpn_blob_names(cfg=c)}
        roidb = [{"height": 100, "width": 120, "boxes": np.zeros((0, 4), np.float32), "gt_classes": np.zeros(0, np.int32),
                  "is_crowd": np.zeros(0, bool)}]
        with pytest.raises(Exception):                                     # no CPU path: nothing is computed on the host
            rpn.add_rpn_blobs(blobs, [1.0], roidb, cfg=c)
