"""GPU parity: mask paste-back, the reference-signature shims and the whole frame-batched step.
Bars: binary masks equal to the reference's (golden) except where the probability is within 1e-5
of the 0.5 threshold; pasted probabilities rtol 1e-5 + atol 4e-6 vs cv2.resize."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


class _Ops:
    """module-level `ops` for the tests below (imported lazily: collection must not need the .so)"""
    def __getattr__(self, name):
        from vosdetectron_b200 import ops as real
        return getattr(real, name)


ops = _Ops()


def test_paste_golden(golden, orc):
    from vosdetectron_b200 import ops
    g = golden("paste")
    fh, fw = (int(v) for v in g["frame_hw"])
    ref = np.unpackbits(g["packed"], axis=-1)[..., :fw]
    out, prob = ops.paste_masks_cuda(_cu(g["masks"]), _cu(g["cls"]), _cu(g["boxes"]), fh, fw, 0.5, want_prob=True)
    out, prob = out.cpu().numpy(), prob.cpu().numpy()
    diff = out != ref
    assert np.all(np.abs(prob[diff] - 0.5) < 1e-5) and diff.sum() <= 2
    out2 = ops.paste_masks_cuda(_cu(g["masks"]), _cu(g["cls"]), _cu(g["boxes"]), fh, fw, 0.5).cpu().numpy()
    assert np.array_equal(out, out2)                       # prob / no-prob kernels agree
    _, po = orc.paste_masks(g["masks"], g["cls"], g["boxes"], fh, fw, want_prob=True)
    assert np.allclose(prob, po, rtol=1e-5, atol=4e-6)


def test_paste_full_size_vs_cv2(synth, orc):
    """BASELINE config 3: 100 detections, M=28, 480x854 frame (41 MB of output)."""
    cv2 = pytest.importorskip("cv2")
    from vosdetectron_b200 import ops
    boxes, cls, masks = synth.detections(3000, 100)
    fh, fw = synth.DAVIS_FRAME
    out, prob = ops.paste_masks_cuda(_cu(masks), _cu(cls), _cu(boxes), fh, fw, 0.5, want_prob=True)
    out, prob = out.cpu().numpy(), prob.cpu().numpy()
    ro, rp = orc.paste_masks(masks, cls, boxes, fh, fw, want_prob=True, resize=lambda p, w, h: cv2.resize(p, (w, h)))
    assert np.allclose(prob, rp, rtol=1e-5, atol=4e-6), np.abs(prob - rp).max()
    diff = out != ro
    assert np.all(np.abs(rp[diff] - 0.5) < 1e-5) and diff.sum() < 20, diff.sum()
    assert out.sum() > 100000
    # every byte written: pre-poisoned output buffer is not observable through the API, so check
    # instead that nothing outside the expanded boxes is set
    eb = orc.expand_boxes(boxes, 30.0 / 28.0).astype(np.int32)
    for i in (0, 17, 99):
        m = out[i].copy()
        m[max(eb[i, 1], 0):eb[i, 3] + 1, max(eb[i, 0], 0):eb[i, 2] + 1] = 0
        assert m.sum() == 0


def test_paste_odd_sizes_and_class_agnostic(synth, orc):
    from vosdetectron_b200 import ops
    boxes, cls, masks = synth.detections(31, 7, (37, 53), 14, 3)          # 37*53 not a multiple of 16
    out = ops.paste_masks_cuda(_cu(masks), None, _cu(boxes), 37, 53, 0.5).cpu().numpy()
    ro, rp = orc.paste_masks(masks, cls, boxes, 37, 53, cls_specific=False, want_prob=True)
    diff = out != ro
    assert np.all(np.abs(rp[diff] - 0.5) < 1e-5)
    assert ops.paste_masks_cuda(_cu(masks[:0]), None, _cu(boxes[:0]), 37, 53, 0.5).shape == (0, 37, 53)


def test_segm_results_signature(golden):
    from vosdetectron_b200.core import test as core_test
    from vosdetectron_b200.config import RegionConfig
    g = golden("paste")
    fh, fw = (int(v) for v in g["frame_hw"])
    K = g["masks"].shape[1]
    cfg = RegionConfig(num_classes=K, mrcnn_resolution=28)
    cls_boxes = [np.zeros((int((g["cls"] == j).sum()), 5), np.float32) for j in range(K)]
    segms = core_test.segm_results(cls_boxes, g["masks"], g["boxes"], fh, fw, cfg)
    assert len(segms) == K and segms[0] == []
    assert [len(s) for s in segms] == [len(b) for b in cls_boxes]
    ref = np.unpackbits(g["packed"], axis=-1)[..., :fw]
    i = 0
    for j in range(1, K):
        for rle in segms[j]:
            assert rle["size"] == [fh, fw] and isinstance(rle["counts"], str)
            assert rle == core_test.rle_encode(ref[i]) or np.abs(int(ref[i].sum()) - _rle_area(rle)) <= 2
            i += 1


def _rle_area(rle):
    # decode the compressed counts and sum the odd runs
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
    return sum(cnts[1::2])


def test_pipeline_step_vs_oracle(synth, orc):
    """Whole frame-batched step on a reduced blob against the oracle chain, frame by frame."""
    from vosdetectron_b200.pipeline import RegionPipeline
    from vosdetectron_b200.config import RegionConfig, RpnMode
    B, D, C, K, M = 3, 12, 8, 5, 28
    blob, frame, scale = (192, 256), (96, 128), 2.0
    cfg = RegionConfig(test=RpnMode(300, 100), num_classes=K)
    pipe = RegionPipeline(cfg)
    rpn = synth.rpn_outputs(21, blob, B)
    feats = synth.fpn_features(22, blob, B, synth.ROI_LEVELS, C)
    im_info = np.tile(np.array([[blob[0], blob[1], scale]], np.float32), (B, 1))
    det = [synth.detections(23 + b, D, frame, M, K) for b in range(B)]
    det_boxes = np.stack([d[0] for d in det]); det_cls = np.stack([d[1] for d in det]); det_masks = np.stack([d[2] for d in det])
    out = pipe.step({l: (_cu(rpn[l][0]), _cu(rpn[l][1])) for l in rpn}, _cu(im_info),
                    {l: _cu(feats[l]) for l in feats}, _cu(det_boxes), _cu(det_cls), _cu(det_masks), frame, scale)
    torch.cuda.synchronize()
    cnt = out["roi_count"].cpu().numpy()
    for b in range(B):
        rl, pl = [], []
        for l in synth.FPN_LEVELS:
            r, p = orc.generate_proposals(rpn[l][0][b:b + 1], rpn[l][1][b:b + 1], im_info[b:b + 1], orc.fpn_anchors(l),
                                          1. / 2 ** l, 300, 100, 0.7, 0)
            r[:, 0] = b
            rl.append(r); pl.append(p)
        ro = orc.collect(rl, pl, 100)
        n = int(cnt[b])
        assert n == len(ro)
        got = out["rois"][b, :n].cpu().numpy()
        assert np.allclose(got, ro, rtol=1e-5, atol=1e-5)
        assert np.array_equal(out["roi_level"][b, :n].cpu().numpy(), orc.map_rois_to_fpn_levels(ro[:, 1:5]).astype(np.int32))
        # box features: oracle RoIAlign on the device-produced rois (identical inputs -> 1e-5)
        blobs = orc.distribute(got)
        fo = orc.roi_feature_transform([feats[l] for l in (5, 4, 3, 2)], blobs, 'rois', 7, [1. / 32, 1. / 16, 1. / 8, 1. / 4], 2)
        bf = out["box_feats"].view(B, -1, C, 7, 7)[b, :n].cpu().numpy()
        assert np.abs(bf - fo).max() <= 1e-5 * np.abs(fo).max()
        # mask branch
        mr = np.concatenate([np.full((D, 1), b, np.float32), det_boxes[b] * np.float32(scale)], axis=1)
        mb = orc.distribute(mr)
        mo = orc.roi_feature_transform([feats[l] for l in (5, 4, 3, 2)], mb, 'rois', 14, [1. / 32, 1. / 16, 1. / 8, 1. / 4], 2)
        mf = out["mask_feats"].view(B, D, C, 14, 14)[b].cpu().numpy()
        assert np.abs(mf - mo).max() <= 1e-5 * np.abs(mo).max()
        po, pp = orc.paste_masks(det_masks[b], det_cls[b], det_boxes[b], frame[0], frame[1], want_prob=True)
        diff = out["masks"][b].cpu().numpy() != po
        assert np.all(np.abs(pp[diff] - 0.5) < 1e-5)


def test_launch_counter_counts_library_kernels(synth):
    from vosdetectron_b200 import _lib, ops
    before = _lib.launch_count()
    f = torch.randn(1, 4, 10, 12, device="cuda")
    ops.roi_align_forward(f, torch.tensor([[0, 1, 1, 20, 20.]], device="cuda"), 7, 7, 0.25, 2)
    assert _lib.launch_count() == before + 1


def test_pack_mask_bits_matches_torch_reference():
    from vosdetectron_b200.pipeline import pack_mask_bits, unpack_mask_bits
    for shape in ((3, 5, 480, 854), (2, 37, 53), (1, 1, 8, 8)):
        m = (torch.rand(shape, device="cuda") > 0.5).to(torch.uint8)
        p = pack_mask_bits(m)
        assert torch.equal(p.cpu(), pack_mask_bits(m.cpu()))
        assert torch.equal(unpack_mask_bits(p, shape[-2], shape[-1]), m)


def test_paste_packed_equals_packing_the_dense_paste(synth):
    """vosd_paste_masks_packed: the 1-bit copy written by the paste kernel itself == pack_mask_bits(dense paste),
    with and without the dense output, on the DAVIS frame (fast kernel) and an odd frame (flat kernel + packer)."""
    from vosdetectron_b200.pipeline import pack_mask_bits
    for (fh, fw), R in (((480, 854), 40), ((37, 53), 9)):
        boxes, cls, masks = synth.detections(77, R, (fh, fw), 28, 5)
        dense = ops.paste_masks_cuda(_cu(masks), _cu(cls), _cu(boxes), fh, fw, 0.5)
        want = pack_mask_bits(dense)
        d2, p2 = ops.paste_masks_packed_cuda(_cu(masks), _cu(cls), _cu(boxes), fh, fw, 0.5)
        assert torch.equal(d2, dense) and torch.equal(p2, want)
        if (fh * fw) % 16 == 0:
            d3, p3 = ops.paste_masks_packed_cuda(_cu(masks), _cu(cls), _cu(boxes), fh, fw, 0.5, want_dense=False)
            assert d3 is None and torch.equal(p3, want)
        else:
            with pytest.raises(Exception):
                ops.paste_masks_packed_cuda(_cu(masks), _cu(cls), _cu(boxes), fh, fw, 0.5, want_dense=False)


def _rle_check(orc, masks, cls, boxes, fh, fw):
    dense = ops.paste_masks_cuda(_cu(masks), None if cls is None else _cu(cls), _cu(boxes), fh, fw, 0.5).cpu().numpy()
    small = fh * fw < 4096          # tiny frames with noisy masks: worst-case arenas; else the defaults
    out = ops.paste_rle_cuda(_cu(masks), None if cls is None else _cu(cls), _cu(boxes), fh, fw, 0.5,
                             run_capacity=len(boxes) * (fh * fw + 2) if small else None,
                             str_capacity=len(boxes) * 6 * (fh * fw + 2) if small else None)
    assert int(out["status"].abs().max().item()) == 0
    runs = out["runs"].cpu().numpy().view(np.uint32)
    ro, rc = out["run_offset"].cpu().numpy(), out["run_count"].cpu().numpy()
    chars = out["chars"].cpu().numpy().tobytes()
    so, sl = out["str_offset"].cpu().numpy(), out["str_len"].cpu().numpy()
    used = out["cursors"].cpu().numpy()
    assert used[0] == rc.sum() and used[1] == sl.sum()
    for i in range(len(boxes)):
        want = orc.rle_counts_fast(dense[i])
        got = runs[ro[i]:ro[i] + rc[i]].tolist()
        assert got == want, (i, boxes[i], got[:8], want[:8])
        s = chars[so[i]:so[i] + sl[i]].decode('ascii')
        assert s == orc.rle_to_string(want)
        assert np.array_equal(orc.rle_decode(orc.rle_from_string(s), fh, fw), dense[i])
    return dense


def test_paste_rle_matches_oracle_on_the_dense_paste(orc, synth, golden):
    """vosd_paste_rle: run lengths and 'counts' strings == the oracle's rleEncode / rleToString of the dense paste
    (which is pinned to the reference), bit for bit; decoding the string gives the dense mask back."""
    g = golden("paste")
    fh, fw = (int(v) for v in g["frame_hw"])
    _rle_check(orc, g["masks"], g["cls"], g["boxes"], fh, fw)
    boxes, cls, masks = synth.detections(31, 60, (480, 854), 28, 7)
    _rle_check(orc, masks, cls, boxes, 480, 854)


def test_paste_rle_edge_boxes(orc, synth):
    """Boxes that miss the frame, cover it completely (all-ones mask: counts [0, HW]), touch the bottom but not the top
    edge (the run continues into row 0 of the next column), touch only the top, a 1-pixel box; odd frame size."""
    fh, fw, M = 37, 53, 28
    boxes = np.array([[-50, -50, -20, -20],          # outside
                      [-30, -30, 90, 90],            # covers everything
                      [10, 20, 30, 36.5],            # bottom edge, not top
                      [10, -5, 30, 12],              # top edge, not bottom
                      [5, 0, 6, 36.9],               # full height, narrow
                      [20, 20, 20.4, 20.4],          # ~1 pixel
                      [0, 0, 52.9, 36.9],            # exactly the frame
                      [40, 10, 70, 60]], np.float32)  # hangs over right and bottom
    R = len(boxes)
    masks = np.ones((R, 1, M, M), np.float32)
    dense = _rle_check(orc, masks, None, boxes, fh, fw)
    assert dense[0].sum() == 0 and dense[1].all()
    rs = np.random.RandomState(3)
    masks = rs.rand(R, 1, M, M).astype(np.float32)      # noisy masks: many runs per column
    _rle_check(orc, masks, None, boxes, fh, fw)
    _, _, m2 = synth.detections(5, R, (fh, fw), M, 3)
    _rle_check(orc, m2, np.zeros(R, np.int32) + 2, boxes, fh, fw)


def test_paste_rle_overflow_status_and_retry(orc, synth):
    boxes, cls, masks = synth.detections(11, 12, (96, 128), 28, 3)
    out = ops.paste_rle_cuda(_cu(masks), _cu(cls), _cu(boxes), 96, 128, 0.5, run_capacity=20, str_capacity=40)
    st = out["status"].cpu().numpy()
    assert (st != 0).any() and set(st.tolist()) <= {0, 1, 2}
    need = out["cursors"].cpu().numpy()
    assert need[0] == out["run_count"].sum().item()          # what a retry must provide
    rles = ops.rle_results(_cu(masks), _cu(cls), _cu(boxes), 96, 128, 0.5)
    dense = ops.paste_masks_cuda(_cu(masks), _cu(cls), _cu(boxes), 96, 128, 0.5).cpu().numpy()
    assert rles == [orc.rle_encode(d) for d in dense]
    assert ops.rle_results(_cu(masks[:0]), _cu(cls[:0]), _cu(boxes[:0]), 96, 128, 0.5) == []
