"""GPU parity: RoIAlign forward / backward through the C ABI against
  (i)   the reference kernel itself (oracle/_ref/libref_roialign.so = the unmodified
        roi_align_kernel.cu built for sm_100a) run on the same GPU;
  (ii)  the CPU oracle (oracle/oracle.c);
  (iii) the torchvision golden (loose anchor).
Tolerances (north_star: RoIAlign within 1e-5 relative, fp32):
  * "staged" and "generic" kernel families: forward BIT-IDENTICAL to the reference kernel (0 ulp);
  * "default" family: the separable forward (csrc/roialign_sep.cuh) evaluates the same sum in a different
    order -> |out - ref| <= 1e-5*|ref| + 1e-6*max|ref|; wherever the separable kernel does not apply
    (sampling_ratio != 2, pooled width not 7/14/28, C % 32 != 0) the default is the staged kernel: 0 ulp;
  * backward: rtol 1e-5 + atol 1e-6*max|grad| (atomics reorder the fp32 sum, also in the reference)."""
import ctypes
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_roialign.so")


@pytest.fixture(scope="module")
def refk():
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref/libref_roialign.so not built")
    lib = ctypes.CDLL(REF_SO)
    vp = ctypes.c_void_p
    lib.ROIAlignForwardLaucher.argtypes = [vp, ctypes.c_float] + [ctypes.c_int] * 7 + [vp, vp, vp]
    lib.ROIAlignBackwardLaucher.argtypes = [vp, ctypes.c_float] + [ctypes.c_int] * 8 + [vp, vp, vp]

    class R:
        @staticmethod
        def fwd(f, rois, ph, pw, scale, sr):
            N, C, H, W = f.shape
            out = torch.zeros((rois.shape[0], C, ph, pw), device=f.device)
            if rois.shape[0]:
                lib.ROIAlignForwardLaucher(f.data_ptr(), scale, rois.shape[0], H, W, C, ph, pw, sr, rois.data_ptr(),
                                           out.data_ptr(), torch.cuda.current_stream().cuda_stream)
            return out

        @staticmethod
        def bwd(g, rois, shape, ph, pw, scale, sr):
            N, C, H, W = shape
            out = torch.zeros(shape, device=g.device)
            if rois.shape[0]:
                lib.ROIAlignBackwardLaucher(g.data_ptr(), scale, N, rois.shape[0], H, W, C, ph, pw, sr,
                                            rois.data_ptr(), out.data_ptr(), torch.cuda.current_stream().cuda_stream)
            return out
    return R


@pytest.fixture(params=["default", "generic", "staged"])
def path(request):
    """Run a test through every RoIAlign kernel family: default (staged forward + atomic backward),
    generic un-staged kernels, and staged forward + staged backward."""
    from vosdetectron_b200 import _lib
    old = _lib.load().vosd_debug_force_generic({"default": 0, "generic": 1, "staged": 2}[request.param])
    yield request.param
    _lib.load().vosd_debug_force_generic(old)


def ulp_diff(a, b):
    ia = a.view(torch.int32).long()
    ib = b.view(torch.int32).long()
    ia = torch.where(ia < 0, -(ia & 0x7fffffff), ia)
    ib = torch.where(ib < 0, -(ib & 0x7fffffff), ib)
    return (ia - ib).abs()


def sep_applies(path, C, pw, sr):
    """True where the default family routes the forward through the separable kernel."""
    return path == "default" and sr == 2 and pw in (7, 14, 28) and C % 32 == 0


def assert_forward(out, ref, exact, what=""):
    if exact:
        d = ulp_diff(out, ref)
        assert int(d.max()) == 0, "%s ulp histogram: %s" % (what, torch.bincount(d.flatten().clamp(max=8)).tolist())
    else:
        tol = 1e-5 * ref.abs() + 1e-6 * float(ref.abs().max())
        bad = (out - ref).abs() > tol
        assert not bool(bad.any()), "%s max err %g (max|ref| %g), %d elements out of tolerance" % (
            what, float((out - ref).abs().max()), float(ref.abs().max()), int(bad.sum()))


def _case(synth, lvl, R=150, C=16, N=2, seed=0, blob=None):
    blob = blob or synth.COCO_BLOB
    feats = synth.fpn_features(100 + seed, blob, N, (lvl,), C)[lvl]
    rois = synth.random_rois(200 + seed, R, blob, N)
    return torch.from_numpy(feats).cuda(), torch.from_numpy(rois).cuda()


@pytest.mark.parametrize("lvl", [2, 3, 4, 5])
@pytest.mark.parametrize("res,sr,C", [(7, 2, 16), (14, 2, 16), (7, 0, 16), (3, 1, 16), (7, 2, 64), (14, 2, 32), (28, 2, 32)])
def test_forward_vs_reference_kernel(refk, synth, orc, path, lvl, res, sr, C):
    from vosdetectron_b200 import ops
    f, rois = _case(synth, lvl, C=C, seed=lvl)
    scale = 1.0 / 2 ** lvl
    out = ops.roi_align_forward(f, rois, res, res, scale, sr)
    ref = refk.fwd(f, rois, res, res, scale, sr)
    torch.cuda.synchronize()
    assert_forward(out, ref, not sep_applies(path, C, res, sr), "lvl %d res %d" % (lvl, res))
    o = orc.roi_align_forward(f.cpu().numpy(), rois.cpu().numpy(), res, res, scale, sr)
    err = np.abs(out.cpu().numpy() - o).max()
    assert err <= 1e-5 * np.abs(o).max(), err


def test_forward_edge_rois_vs_reference_kernel(refk, synth, orc, path):
    from vosdetectron_b200 import ops
    for lvl in (2, 5):
        for C in (8, 32):
            f = torch.from_numpy(synth.fpn_features(7, synth.COCO_BLOB, 1, (lvl,), C)[lvl]).cuda()
            rois = torch.from_numpy(synth.edge_rois()).cuda()
            for res, sr in ((7, 2), (14, 2), (7, 0)):
                out = ops.roi_align_forward(f, rois, res, res, 1.0 / 2 ** lvl, sr)
                ref = refk.fwd(f, rois, res, res, 1.0 / 2 ** lvl, sr)
                exact = not sep_applies(path, C, res, sr)
                assert_forward(out, ref, exact, "edge lvl %d C %d res %d sr %d" % (lvl, C, res, sr))
                o = orc.roi_align_forward(f.cpu().numpy(), rois.cpu().numpy(), res, res, 1.0 / 2 ** lvl, sr)
                if exact:
                    assert np.array_equal(out.cpu().numpy(), o), (lvl, res, sr)


@pytest.mark.parametrize("C", [40, 64])
def test_forward_oversize_rois_and_odd_channels(refk, synth, path, C):
    """RoIs far larger than the staged tile / the separable kernel's ring (banding + in-kernel direct
    gather), C not a multiple of 32, non-square pooled sizes and a large fixed sampling grid."""
    from vosdetectron_b200 import ops
    f = torch.from_numpy(synth.fpn_features(77, synth.COCO_BLOB, 2, (2,), C)[2]).cuda()     # 200x336
    rois = torch.tensor([[0, 0, 0, 1343, 799], [1, 10, 20, 1300, 90], [0, 5, 5, 90, 790], [1, 300, 300, 340, 330],
                         [0, 100, 100, 900, 700], [1, 40, 40, 160, 700], [0, 0, 0, 127, 127]],
                        dtype=torch.float32, device="cuda")
    for (ph, pw, sr) in ((7, 7, 2), (14, 14, 2), (10, 7, 2), (3, 14, 2), (7, 7, 0), (5, 9, 3), (2, 2, 20)):
        out = ops.roi_align_forward(f, rois, ph, pw, 0.25, sr)
        ref = refk.fwd(f, rois, ph, pw, 0.25, sr)
        assert_forward(out, ref, not sep_applies(path, C, pw, sr), "C %d %dx%d sr %d" % (C, ph, pw, sr))


def test_forward_vs_torchvision_golden(golden):
    from vosdetectron_b200 import ops
    g = golden("roialign_tv")
    rois = torch.from_numpy(g["rois"]).cuda()
    for lvl in (2, 3, 4, 5):
        f = torch.from_numpy(g["feat%d" % lvl]).cuda()
        for res in (7, 14):
            out = ops.roi_align_forward(f, rois, res, res, 1.0 / 2 ** lvl, 2).cpu().numpy()
            assert np.abs(out - g["tv_fwd_l%d_r%d" % (lvl, res)]).max() < 1e-4


def test_empty_and_cpu_inputs(synth):
    from vosdetectron_b200 import ops
    from vosdetectron_b200.modeling.roi_xfrom.roi_align.functions.roi_align import RoIAlignFunction
    f = torch.randn(1, 4, 10, 12, device="cuda")
    out = ops.roi_align_forward(f, torch.zeros((0, 5), device="cuda"), 7, 7, 0.25, 2)
    assert out.shape == (0, 4, 7, 7)
    with pytest.raises(NotImplementedError):
        RoIAlignFunction(7, 7, 0.25, 2)(f.cpu(), torch.zeros((1, 5)))
    with pytest.raises(ValueError):
        ops.roi_align_forward(f, torch.zeros((3, 4), device="cuda"), 7, 7, 0.25, 2)


@pytest.mark.parametrize("lvl,res,sr,C", [(2, 7, 2, 8), (3, 14, 2, 8), (4, 7, 0, 8), (5, 14, 2, 8),
                                          (2, 7, 2, 64), (3, 7, 2, 32), (4, 14, 2, 32), (5, 7, 2, 32), (5, 14, 2, 64),
                                          (3, 28, 2, 32), (4, 7, 0, 32)])
def test_backward_vs_reference_kernel_and_oracle(refk, synth, orc, path, lvl, res, sr, C):
    """C % 32 == 0 with sampling_ratio 2 and pooled width 7 / 14 / 28 routes the default family through the
    separable scatter (csrc/roialign_sep.cuh); everything else through the record-based atomics."""
    from vosdetectron_b200 import ops
    f, rois = _case(synth, lvl, R=120, C=C, seed=10 + lvl)
    scale = 1.0 / 2 ** lvl
    g = torch.from_numpy(np.random.RandomState(lvl).standard_normal((rois.shape[0], C, res, res)).astype(np.float32)).cuda()
    mine = ops.roi_align_backward(g, rois, f.shape, res, res, scale, sr)
    ref = refk.bwd(g, rois, tuple(f.shape), res, res, scale, sr)
    tol = 1e-5 * ref.abs() + 1e-6 * float(ref.abs().max())
    assert bool(((mine - ref).abs() <= tol).all()), float((mine - ref).abs().max())
    o = orc.roi_align_backward(g.cpu().numpy(), rois.cpu().numpy(), tuple(f.shape), res, res, scale, sr)
    o = torch.from_numpy(o).cuda()
    tol = 1e-5 * o.abs() + 1e-6 * float(o.abs().max())
    assert bool(((mine - o).abs() <= tol).all())


def test_autograd_function_and_modules(synth):
    from vosdetectron_b200.modeling.roi_xfrom.roi_align.functions.roi_align import RoIAlignFunction
    from vosdetectron_b200.modeling.roi_xfrom.roi_align.modules.roi_align import RoIAlign, RoIAlignAvg, RoIAlignMax
    torchvision = pytest.importorskip("torchvision")
    f, rois = _case(synth, 3, R=40, C=4, seed=3)
    f.requires_grad_(True)
    out = RoIAlignFunction(7, 7, 0.125, 2)(f, rois)
    w = torch.randn_like(out)
    (out * w).sum().backward()
    g_mine = f.grad.clone()
    f2 = f.detach().clone().requires_grad_(True)
    out2 = torchvision.ops.roi_align(f2, rois, (7, 7), 0.125, sampling_ratio=2, aligned=False)
    (out2 * w).sum().backward()
    assert (out - out2).abs().max() < 1e-4
    assert (g_mine - f2.grad).abs().max() < 1e-4 * max(1.0, float(f2.grad.abs().max()))
    assert RoIAlign(7, 7, 0.125, 2)(f, rois).shape == (40, 4, 7, 7)
    assert RoIAlignAvg(7, 7, 0.125, 2)(f, rois).shape == (40, 4, 7, 7)
    assert RoIAlignMax(7, 7, 0.125, 2)(f, rois).shape == (40, 4, 7, 7)


def test_multilevel_matches_reference_loop(refk, synth, orc, path):
    """roi_feature_transform: one launch == per-level launches + cat + restore (model_builder.py:271-303)."""
    from vosdetectron_b200.modeling.model_builder import roi_feature_transform
    N, C = 2, 8
    feats = synth.fpn_features(55, synth.COCO_BLOB, N, synth.ROI_LEVELS, C)
    rois = synth.random_rois(56, 400, synth.COCO_BLOB, N)
    blobs = orc.distribute(rois)                       # reference-format rpn_ret (ndarrays)
    blobs_in = [torch.from_numpy(feats[l]).cuda() for l in (5, 4, 3, 2)]      # coarsest first
    scales = [1. / 32, 1. / 16, 1. / 8, 1. / 4]
    for res in (7, 14):
        out = roi_feature_transform(blobs_in, blobs, 'rois', 'RoIAlign', res, scales, 2)
        parts = []
        for lvl in (2, 3, 4, 5):
            r = blobs['rois_fpn%d' % lvl]
            if len(r):
                parts.append(refk.fwd(blobs_in[5 - lvl], torch.from_numpy(r).cuda(), res, res, scales[5 - lvl], 2))
        ref = torch.cat(parts)[torch.from_numpy(blobs['rois_idx_restore_int32'].astype(np.int64)).cuda()]
        assert torch.equal(out, ref)                    # C = 8: every family is bit-exact here
    # tensors in rpn_ret work too, and gradients flow to every level
    t_blobs = {k: torch.from_numpy(v).cuda() for k, v in blobs.items()}
    leaf = [b.clone().requires_grad_(True) for b in blobs_in]
    out = roi_feature_transform(leaf, t_blobs, 'rois', 'RoIAlign', 7, scales, 2)
    assert torch.equal(out, roi_feature_transform(blobs_in, blobs, 'rois', 'RoIAlign', 7, scales, 2))
    gout = torch.randn_like(out)
    out.backward(gout)
    order = np.concatenate([np.where(orc.map_rois_to_fpn_levels(rois[:, 1:5]) == l)[0] for l in (2, 3, 4, 5)])
    for lvl in (2, 3, 4, 5):
        idx = np.where(orc.map_rois_to_fpn_levels(rois[:, 1:5]) == lvl)[0]
        ref = refk.bwd(gout[torch.from_numpy(idx).cuda()].contiguous(), torch.from_numpy(rois[idx]).cuda(),
                       tuple(blobs_in[5 - lvl].shape), 7, 7, scales[5 - lvl], 2)
        got = leaf[5 - lvl].grad
        tol = 1e-5 * ref.abs() + 1e-6 * float(ref.abs().max())
        assert bool(((got - ref).abs() <= tol).all()), lvl
    assert len(order) == len(rois)


def test_full_size_vs_reference_kernel(refk, synth, path):
    """BASELINE config 2 at full size (1000 RoIs x 256 ch over P2-P5, 7x7 and 14x14) against the reference
    kernel launched per level (model_builder.py:271-303), every kernel family."""
    from vosdetectron_b200 import ops
    feats = synth.fpn_features(2000, synth.COCO_BLOB, 1, synth.ROI_LEVELS, 256)
    rois = torch.from_numpy(synth.random_rois(2001, 1000, synth.COCO_BLOB, 1)).cuda()
    fl = [torch.from_numpy(feats[l]).cuda() for l in synth.ROI_LEVELS]
    sc = [1.0 / 2 ** l for l in synth.ROI_LEVELS]
    level, _, order, restore = ops.distribute_cuda(rois)
    lv = (level - 2).to(torch.int32)
    for res in (7, 14):
        out = ops.roi_align_ml_forward(fl, sc, rois, lv, res, res, 2)
        ref = torch.empty_like(out)
        for i in range(len(fl)):
            idx = torch.nonzero(lv == i).flatten()
            if len(idx):
                ref[idx] = refk.fwd(fl[i], rois[idx].contiguous(), res, res, sc[i], 2)
        assert_forward(out, ref, not sep_applies(path, 256, res, 2), "full size res %d" % res)
        if sep_applies(path, 256, res, 2):
            d = ulp_diff(out, ref).flatten()
            print("separable forward res %d: ulp histogram (0..7, >=8) %s, max abs err %.3g" % (
                res, torch.bincount(d.clamp(max=8), minlength=9).tolist(), float((out - ref).abs().max())))


def test_backward_edge_and_oversize_rois(refk, synth, path):
    """Degenerate / out-of-map RoIs and footprints beyond the separable kernel's ring, C % 32 == 0."""
    from vosdetectron_b200 import ops
    f = torch.from_numpy(synth.fpn_features(78, synth.COCO_BLOB, 2, (2,), 32)[2]).cuda()
    big = np.array([[0, 0, 0, 1343, 799], [1, 10, 20, 1300, 90], [0, 5, 5, 90, 790], [1, 300, 300, 340, 330],
                    [0, 100, 100, 900, 700], [1, 40, 40, 160, 700], [0, 0, 0, 127, 127]], np.float32)
    rois = torch.from_numpy(np.concatenate([synth.edge_rois(), big])).cuda()
    for res in (7, 14):
        g = torch.randn((rois.shape[0], 32, res, res), device="cuda")
        mine = ops.roi_align_backward(g, rois, f.shape, res, res, 0.25, 2)
        ref = refk.bwd(g, rois, tuple(f.shape), res, res, 0.25, 2)
        tol = 1e-5 * ref.abs() + 1e-6 * float(ref.abs().max())
        assert bool(((mine - ref).abs() <= tol).all()), (res, float((mine - ref).abs().max()))


def test_full_size_backward_vs_reference_kernel(refk, synth, path):
    """BASELINE config 4 box head at full size: 1024 RoIs x 256 ch x 7x7 over 2 frames, and the 14x14 mask head
    on 256 RoIs, against the reference kernel per level."""
    from vosdetectron_b200 import ops
    feats = synth.fpn_features(2000, synth.COCO_BLOB, 2, synth.ROI_LEVELS, 256)
    shapes = [feats[l].shape for l in synth.ROI_LEVELS]
    sc = [1.0 / 2 ** l for l in synth.ROI_LEVELS]
    for R, res in ((1024, 7), (256, 14)):
        rois = torch.from_numpy(synth.random_rois(2001 + res, R, synth.COCO_BLOB, 2)).cuda()
        level, _, order, restore = ops.distribute_cuda(rois)
        lv = (level - 2).to(torch.int32)
        g = torch.randn((R, 256, res, res), device="cuda")
        grads = ops.roi_align_ml_backward(g, shapes, sc, rois, lv, res, res, 2)
        for i in range(len(shapes)):
            idx = torch.nonzero(lv == i).flatten()
            ref = refk.bwd(g[idx].contiguous(), rois[idx].contiguous(), tuple(shapes[i]), res, res, sc[i], 2)
            tol = 1e-5 * ref.abs() + 1e-6 * float(ref.abs().max())
            assert bool(((grads[i] - ref).abs() <= tol).all()), (res, i, float((grads[i] - ref).abs().max()))


def test_full_size_properties(synth):
    """BASELINE config 2 sizes: 1000 RoIs x 256 ch over P2-P5 -- size-independent checks."""
    from vosdetectron_b200 import ops
    feats = synth.fpn_features(2000, synth.COCO_BLOB, 1, synth.ROI_LEVELS, 256)
    rois = torch.from_numpy(synth.random_rois(2001, 1000, synth.COCO_BLOB, 1)).cuda()
    fl = [torch.from_numpy(feats[l]).cuda() for l in synth.ROI_LEVELS]
    sc = [1.0 / 2 ** l for l in synth.ROI_LEVELS]
    level, _, order, restore = ops.distribute_cuda(rois)
    assert torch.equal(order[restore.long()].long(), torch.arange(1000, device="cuda"))
    lv = (level - 2).to(torch.int32)
    for res in (7, 14):
        a = ops.roi_align_ml_forward(fl, sc, rois, lv, res, res, 2)
        # linearity in the features
        b = ops.roi_align_ml_forward([2.0 * f for f in fl], sc, rois, lv, res, res, 2)
        assert torch.equal(b, 2.0 * a)
        # constant features -> constant output for RoIs inside the map
        ones = ops.roi_align_ml_forward([torch.ones_like(f) for f in fl], sc, rois, lv, res, res, 2)
        assert float((ones - 1).abs().max()) < 1e-6
        # adjoint identity <A f, g> == <f, A^T g>
        g = torch.randn_like(a)
        grads = ops.roi_align_ml_backward(g, [f.shape for f in fl], sc, rois, lv, res, res, 2)
        lhs = float((a.double() * g.double()).sum())
        rhs = float(sum((f.double() * gr.double()).sum() for f, gr in zip(fl, grads)))
        assert abs(lhs - rhs) <= 1e-5 * max(1.0, abs(lhs)), (lhs, rhs)
