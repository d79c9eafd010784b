"""GPU parity: proposals (top-k + decode + clip + filter + NMS), standalone NMS and
collect/distribute through the C ABI, against the golden vectors produced by the reference and
against the CPU oracle on full-size seeded inputs.
Bars: top-k order, NMS keep indices, collect order, FPN levels, per-level splits and restore
permutation bit-exact (tie-free inputs); decoded boxes rtol 1e-5."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
BOX_RTOL = 1e-5


def _cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _levels(orc, rpn, lvls):
    return [(_cu(rpn[l][0]), _cu(rpn[l][1]), orc.fpn_anchors(l), float(2 ** l)) for l in lvls]


def _split(rois, probs, count, l, N):
    """(L,N,cap,.) device outputs -> reference-format (R,5), (R,1) ndarrays for level index l."""
    c = count.cpu().numpy()
    r, p = rois.cpu().numpy(), probs.cpu().numpy()
    return (np.concatenate([r[l, i, :c[l, i]] for i in range(N)]),
            np.concatenate([p[l, i, :c[l, i], None] for i in range(N)]))


def _assert_rois_close(a, b):
    assert a.shape == b.shape, (a.shape, b.shape)
    assert np.array_equal(a[:, 0], b[:, 0])
    assert np.allclose(a[:, 1:], b[:, 1:], rtol=BOX_RTOL, atol=BOX_RTOL)


def test_generate_proposals_golden(golden, orc):
    from vosdetectron_b200 import ops
    g = golden("proposals")
    lvls = [2, 3, 4, 5, 6]
    inputs = [(_cu(g["scores%d" % l]), _cu(g["deltas%d" % l]), orc.fpn_anchors(l), float(2 ** l)) for l in lvls]
    rois, probs, count = ops.generate_proposals_cuda(inputs, _cu(g["im_info"]), int(g["pre"]), int(g["post"]),
                                                     float(g["thresh"]), 0.0)
    exact = 0
    for i, l in enumerate(lvls):
        r, p = _split(rois, probs, count, i, 2)
        assert np.array_equal(p, g["probs%d" % l]), l          # same boxes kept, same order
        _assert_rois_close(r, g["rois%d" % l])
        exact += int(np.array_equal(r, g["rois%d" % l]))
    assert exact >= 4                                          # decode is bit-identical in practice
    # min_size filter
    rois, probs, count = ops.generate_proposals_cuda(inputs[1:2], _cu(g["im_info"]), int(g["pre"]), int(g["post"]),
                                                     float(g["thresh"]), 16.0)
    r, p = _split(rois, probs, count, 0, 2)
    assert np.array_equal(p, g["probs3_min16"])
    _assert_rois_close(r, g["rois3_min16"])


@pytest.mark.parametrize("pre,post", [(2000, 1000), (1000, 1000), (2000, 2000), (0, 300), (-1, 1000), (20000, 2000),
                                      (30000, 0)])
def test_generate_proposals_full_size_vs_oracle(synth, orc, pre, post):
    """BASELINE config 1: 5 levels of an 800x1344 blob, A=3 (268 569 anchors / image).  pre <= 0 is the full argsort
    branch (generate_proposals.py:131-132) on P3-P6: P3 sends all 50 400 boxes into NMS, beyond VOSD_MAX_TOPK, so the
    call takes the streamed kernel (P2's 201 600 are left to the pre = 20 000 / 30 000 cases: the CPU oracle's NMS is
    O(kept * n)); post = 0 returns every survivor of the 30 000."""
    from vosdetectron_b200 import ops
    N = 2
    rpn = synth.rpn_outputs(1000, synth.COCO_BLOB, N)
    im_info = np.array([[800, 1344, 1.6667], [800, 1067, 1.25]], dtype=np.float32)
    lvls = [2, 3, 4, 5, 6] if pre > 0 else [3, 4, 5, 6]
    rois, probs, count = ops.generate_proposals_cuda(_levels(orc, rpn, lvls), _cu(im_info), pre, post, 0.7, 0.0)
    for i, l in enumerate(lvls):
        ro, po = orc.generate_proposals(rpn[l][0], rpn[l][1], im_info, orc.fpn_anchors(l), 1. / 2 ** l,
                                        pre, post, 0.7, 0)
        r, p = _split(rois, probs, count, i, N)
        assert np.array_equal(p, po), (l, len(p), len(po))
        _assert_rois_close(r, ro)


def test_no_nms_branch_and_limits(synth, orc):
    from vosdetectron_b200 import ops, _lib
    rpn = synth.rpn_outputs(5, (192, 256), 1)
    im_info = np.array([[192, 256, 1.0]], dtype=np.float32)
    rois, probs, count = ops.generate_proposals_cuda(_levels(orc, rpn, [3]), _cu(im_info), 200, 50, 0.0, 0.0)
    ro, po = orc.generate_proposals(rpn[3][0], rpn[3][1], im_info, orc.fpn_anchors(3), 1. / 8, 200, 50, 0.0, 0)
    r, p = _split(rois, probs, count, 0, 1)
    assert len(po) == 200 and np.array_equal(p, po)
    _assert_rois_close(r, ro)
    # no NMS on a segment longer than VOSD_MAX_TOPK (streamed kernel): the 20 000 best boxes in score order
    big = synth.rpn_outputs(6, synth.COCO_BLOB, 1, levels=(3,))
    im_big = np.array([[800, 1344, 1.0]], dtype=np.float32)
    rois, probs, count = ops.generate_proposals_cuda(_levels(orc, big, [3]), _cu(im_big), 20000, 0, 0.0, 0.0)
    ro, po = orc.generate_proposals(big[3][0], big[3][1], im_big, orc.fpn_anchors(3), 1. / 8, 20000, 0, 0.0, 0)
    r, p = _split(rois, probs, count, 0, 1)
    assert len(po) == 20000 and np.array_equal(p, po)
    _assert_rois_close(r, ro)


def test_decode_anchors_and_nan_guard(synth, orc):
    from vosdetectron_b200 import ops
    rpn = synth.rpn_outputs(8, (192, 256), 2)
    im_info = np.array([[192, 256, 1.0], [150, 200, 1.0]], dtype=np.float32)
    d = rpn[3][1]
    boxes = ops.decode_anchors_cuda(_cu(d), orc.fpn_anchors(3), 8.0, _cu(im_info)).cpu().numpy()
    H, W = d.shape[-2:]
    anc = orc.all_anchors_for_level(orc.fpn_anchors(3), H, W, 8.0)
    for i in range(2):
        ref = orc.bbox_transform(anc, d[i].transpose(1, 2, 0).reshape(-1, 4))
        ref = orc.clip_tiled_boxes(ref, im_info[i, :2])
        assert np.allclose(boxes[i], ref, rtol=BOX_RTOL, atol=BOX_RTOL)
        assert np.mean(boxes[i] == ref) > 0.999
    t = _cu(d)
    assert int(ops.any_nan_cuda(t).item()) == 0
    t.view(-1)[12345] = float("nan")
    assert int(ops.any_nan_cuda(t).item()) == 1


def test_nms_golden(golden):
    from vosdetectron_b200 import ops
    g = golden("nms")
    for t in (0.3, 0.5, 0.7):
        for kind in ("unsorted", "sorted"):
            keep, num = ops.nms_cuda(_cu(g["dets_" + kind]), t)
            k = keep[:int(num.item())].cpu().numpy()
            assert k.dtype == np.int64
            assert np.array_equal(k, g["keep_%s_%02d" % (kind, int(t * 10))]), (t, kind)


@pytest.mark.parametrize("n,thresh", [(1, 0.5), (63, 0.5), (64, 0.5), (65, 0.5), (2000, 0.7), (2000, 0.3), (6000, 0.5)])
def test_nms_vs_oracle(synth, orc, n, thresh):
    from vosdetectron_b200 import ops
    d = synth.clustered_dets(n, n)
    keep, num = ops.nms_cuda(_cu(d), thresh)
    k = keep[:int(num.item())].cpu().numpy()
    ko = orc.nms(d, thresh)
    assert np.array_equal(k, ko)
    # idempotence: the survivors do not suppress one another
    keep2, num2 = ops.nms_cuda(_cu(d[k]), thresh)
    assert int(num2.item()) == len(k)


def test_nms_edge_cases(orc):
    from vosdetectron_b200 import ops
    from vosdetectron_b200.utils import boxes as box_utils
    assert box_utils.nms(np.zeros((0, 5), np.float32), 0.5) == []            # boxes.py:331-332
    same = np.tile(np.array([[10, 10, 50, 50, 0.0]], np.float32), (200, 1))
    same[:, 4] = np.linspace(0.1, 0.9, 200)
    assert np.array_equal(box_utils.nms(same, 0.5), orc.nms(same, 0.5))      # one survivor
    ties = np.tile(np.array([[10, 10, 50, 50, 0.5]], np.float32), (70, 1))
    ties[:, 0] += np.arange(70) * 100                                        # disjoint boxes, equal scores
    ties[:, 2] += np.arange(70) * 100
    assert np.array_equal(box_utils.nms(ties, 0.5), np.arange(70))
    d = np.array([[0, 0, 10, 10, 0.9], [0, 0, 10, 10, 0.8], [20, 20, 30, 30, 0.7]], np.float32)
    assert np.array_equal(box_utils.nms(d, 1.0), orc.nms(d, 1.0))            # IoU == thresh -> suppressed (>=)


def test_collect_distribute_golden(golden):
    from vosdetectron_b200.modeling import collect_and_distribute_fpn_rpn_proposals as cd
    from vosdetectron_b200.config import RegionConfig, RpnMode
    g = golden("collect_distribute")
    cfg = RegionConfig(test=RpnMode(300, int(g["post"])))
    inputs = [g["in_rois%d" % l] for l in range(2, 7)] + [g["in_probs%d" % l] for l in range(2, 7)]
    rois = cd.collect(inputs, False, cfg)
    assert rois.dtype == np.float32 and np.array_equal(rois, g["rois"])
    for blobs in (cd.distribute(rois, None, cfg), cd.collect_and_distribute(inputs, False, cfg),
                  cd.CollectAndDistributeFpnRpnProposalsOp(cfg).eval()(inputs, None, None)):
        for k in ("rois", "rois_fpn2", "rois_fpn3", "rois_fpn4", "rois_fpn5", "rois_idx_restore_int32"):
            assert blobs[k].dtype == g[k].dtype and np.array_equal(blobs[k], g[k]), k
    wide = cd.distribute(g["wide_rois"], None, cfg)
    for k in ("rois_fpn2", "rois_fpn3", "rois_fpn4", "rois_fpn5", "rois_idx_restore_int32"):
        assert np.array_equal(wide[k], g["wide_" + k]), k


def test_collect_distribute_batched_vs_oracle(synth, orc):
    """Frame-batched device path: 6 frames, per-frame groups and a 2-image minibatch group."""
    from vosdetectron_b200 import ops
    N = 6
    rpn = synth.rpn_outputs(77, synth.DAVIS_BLOB, N)
    im_info = np.tile(np.array([[768, 1344, synth.DAVIS_SCALE]], dtype=np.float32), (N, 1))
    lvls = [2, 3, 4, 5, 6]
    rois, probs, count = ops.generate_proposals_cuda(_levels(orc, rpn, lvls), _cu(im_info), 1000, 1000, 0.7, 0.0)
    per_level = [_split(rois, probs, count, i, N) for i in range(5)]
    for ipg, post in ((1, 1000), (2, 1500)):
        out = ops.collect_distribute_cuda(rois, probs, count, post, ipg)
        cnt = out["count"].cpu().numpy()
        for gi in range(N // ipg):
            sel = [np.isin(r[:, 0], np.arange(gi * ipg, (gi + 1) * ipg)) for r, _ in per_level]
            ro = orc.collect([r[s] for (r, _), s in zip(per_level, sel)], [p[s] for (_, p), s in zip(per_level, sel)], post)
            n = int(cnt[gi])
            assert n == len(ro)
            got = out["rois"][gi, :n].cpu().numpy()
            assert np.array_equal(got, ro), (ipg, gi)
            d = orc.distribute(ro)
            assert np.array_equal(out["level"][gi, :n].cpu().numpy(), orc.map_rois_to_fpn_levels(ro[:, 1:5]).astype(np.int32))
            assert np.array_equal(out["restore"][gi, :n].cpu().numpy(), d["rois_idx_restore_int32"])
            order = out["order"][gi, :n].cpu().numpy()
            lc = out["level_count"][gi].cpu().numpy()
            off = 0
            for j, lvl in enumerate((2, 3, 4, 5)):
                assert np.array_equal(got[order[off:off + lc[j]]], d["rois_fpn%d" % lvl])
                off += lc[j]
            assert off == n


def test_generate_proposals_op_signature(golden, orc):
    """GenerateProposalsOp(anchors, spatial_scale).forward(...) -> (ndarray (R,5), ndarray (R,1))."""
    from vosdetectron_b200.modeling.generate_proposals import GenerateProposalsOp
    from vosdetectron_b200.config import RegionConfig, RpnMode
    g = golden("proposals")
    cfg = RegionConfig(test=RpnMode(int(g["pre"]), int(g["post"]), float(g["thresh"]), 0.0))
    op = GenerateProposalsOp(orc.fpn_anchors(4), 1. / 16, cfg).eval()
    rois, probs = op.forward(_cu(g["scores4"]), _cu(g["deltas4"]), torch.from_numpy(g["im_info"]))
    assert isinstance(rois, np.ndarray) and rois.dtype == np.float32 and rois.shape[1] == 5
    assert probs.shape == (rois.shape[0], 1) and probs.dtype == np.float32
    assert np.array_equal(probs, g["probs4"])
    _assert_rois_close(rois, g["rois4"])
    bad = _cu(g["deltas4"]).clone()
    bad[0, 0, 0, 0] = float("nan")
    with pytest.raises(ValueError):
        op.forward(_cu(g["scores4"]), bad, torch.from_numpy(g["im_info"]))
    with pytest.raises(NotImplementedError):
        op.forward(torch.from_numpy(g["scores4"]), torch.from_numpy(g["deltas4"]), torch.from_numpy(g["im_info"]))


@pytest.mark.parametrize("kind", ["one_bin", "two_bins", "near_zero"])
def test_topk_select_on_degenerate_score_distributions(synth, orc, kind):
    """The cluster top-k keeps the keys of the leading digit's threshold bin in a shared-memory candidate buffer and
    falls back to rescanning when they do not fit.  `one_bin`: every score of a P2 plane (201 600) lies in ONE
    11-bit bin of the key (fallback path); `two_bins`: the top-1000 cut falls inside a bin holding half of the plane;
    `near_zero`: the usual RPN picture, almost everything near 0 and a thin foreground tail.  Order and boxes against
    the oracle."""
    from vosdetectron_b200 import ops
    rpn = synth.rpn_outputs(77, synth.COCO_BLOB, 1, levels=(2,))
    sc = rpn[2][0]
    n = sc.size
    rank = np.argsort(np.argsort(sc.ravel())).astype(np.float64)          # 0 .. n-1, the plane's own (tie-free) order
    if kind == "one_bin":
        new = 0.90 + 0.04 * (rank + 0.5) / n                               # [0.90, 0.94): inside [0.875, 1.0)
    elif kind == "two_bins":
        new = np.where(rank < n // 2, 0.30 + 0.1 * rank / n, 0.76 + 0.2 * rank / n)
    else:
        new = np.where(rank < n - 1500, 1e-4 + 1e-3 * rank / n, 0.2 + 0.79 * (rank - (n - 1500)) / 1500.0)
    new = new.astype(np.float32).reshape(sc.shape)
    assert len(np.unique(new)) == n                                        # still tie-free in fp32
    rpn = {2: (new, rpn[2][1])}
    im_info = np.array([[800, 1344, 1.0]], dtype=np.float32)
    for pre, post in ((1000, 1000), (6000, 300)):
        rois, probs, count = ops.generate_proposals_cuda(_levels(orc, rpn, [2]), _cu(im_info), pre, post, 0.7, 0.0)
        ro, po = orc.generate_proposals(rpn[2][0], rpn[2][1], im_info, orc.fpn_anchors(2), 0.25, pre, post, 0.7, 0)
        r, p = _split(rois, probs, count, 0, 1)
        assert np.array_equal(p, po), (kind, pre, len(p), len(po))
        _assert_rois_close(r, ro)
