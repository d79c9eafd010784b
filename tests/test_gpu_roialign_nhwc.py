"""GPU parity: channels-last RoIAlign forward (TMA-fed, csrc/roialign_nhwc.cuh) through the C ABI against
  (i)  the reference kernel (oracle/_ref/libref_roialign.so) on the NCHW copy of the same features:
       |out - ref| <= 1e-5*|ref| + 1e-6*max|ref| (north_star: 1e-5 relative; same gate as the separable NCHW kernel);
  (ii) the separable NCHW kernel of this library (workspace-free entry point) on the NCHW copy: same summation
       order -> BIT-identical."""
import ctypes
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_roialign.so")


@pytest.fixture(scope="module")
def ref_fwd():
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref/libref_roialign.so not built")
    lib = ctypes.CDLL(REF_SO)
    vp = ctypes.c_void_p
    lib.ROIAlignForwardLaucher.argtypes = [vp, ctypes.c_float] + [ctypes.c_int] * 7 + [vp, vp, vp]

    def fwd(f, rois, ph, pw, scale, sr):
        N, C, H, W = f.shape
        out = torch.zeros((rois.shape[0], C, ph, pw), device=f.device)
        if rois.shape[0]:
            lib.ROIAlignForwardLaucher(f.data_ptr(), scale, rois.shape[0], H, W, C, ph, pw, sr, rois.data_ptr(),
                                       out.data_ptr(), torch.cuda.current_stream().cuda_stream)
        return out
    return fwd


def sep_forward(f, rois, res, scale):
    """The separable NCHW kernel (csrc/roialign_sep.cuh) through the workspace-free entry point it sits behind."""
    from vosdetectron_b200 import _lib, ops
    out = torch.empty((rois.shape[0], f.shape[1], res, res), dtype=torch.float32, device=f.device)
    _lib.call("vosd_roialign_fwd", ops._ptr(f), float(scale), rois.shape[0], f.shape[2], f.shape[3], f.shape[1], res, res, 2,
              ops._ptr(rois), ops._ptr(out), ops._stream())
    return out


def gate(out, ref, what):
    tol = 1e-5 * ref.abs() + 1e-6 * float(ref.abs().max())
    err = (out - ref).abs()
    assert bool((err <= tol).all()), "%s: max err %g" % (what, float(err.max()))


@pytest.mark.parametrize("res", [7, 14, 28])
def test_single_level_edge_and_random_rois(ref_fwd, synth, res):
    from vosdetectron_b200 import ops
    from vosdetectron_b200.modeling.roi_xfrom.roi_align.functions.roi_align import RoIAlignFunction
    N, C = 2, 64
    f = torch.from_numpy(synth.fpn_features(91, synth.COCO_BLOB, N, (3,), C)[3]).cuda()
    f_cl = f.contiguous(memory_format=torch.channels_last)
    assert ops._is_channels_last(f_cl) and not ops._is_channels_last(f)
    big = np.array([[0, 0, 0, 1343, 799], [1, 10, 20, 1300, 90], [0, 5, 5, 90, 790], [1, 300, 300, 340, 330],
                    [1, 1200, 700, 1343, 799], [0, 1330, 2, 1343.5, 40]], np.float32)      # oversize -> gather path
    rois = torch.from_numpy(np.concatenate([synth.edge_rois(), big, synth.random_rois(92, 300, synth.COCO_BLOB, N)])).cuda()
    out = RoIAlignFunction(res, res, 0.125, 2)(f_cl, rois)            # the reference's call form, channels-last input
    assert out.shape == (rois.shape[0], C, res, res) and out.is_contiguous()
    gate(out, ref_fwd(f, rois, res, res, 0.125, 2), "vs reference kernel, res %d" % res)
    nchw = sep_forward(f, rois, res, 0.125)                           # separable NCHW kernel of this library
    assert torch.equal(out, nchw), "not bit-identical to the NCHW kernel: max |diff| %g" % float((out - nchw).abs().max())


def test_full_size_multilevel(ref_fwd, synth):
    """BASELINE config 2 at full size (1000 RoIs x 256 ch over P2-P5, 7x7 and 14x14), channels-last maps."""
    from vosdetectron_b200 import _lib, ops
    feats = synth.fpn_features(2000, synth.COCO_BLOB, 1, synth.ROI_LEVELS, 256)
    rois = torch.from_numpy(synth.random_rois(2001, 1000, synth.COCO_BLOB, 1)).cuda()
    fl = [torch.from_numpy(feats[l]).cuda() for l in synth.ROI_LEVELS]
    fl_cl = [f.contiguous(memory_format=torch.channels_last) for f in fl]
    sc = [1.0 / 2 ** l for l in synth.ROI_LEVELS]
    level, _, order, restore = ops.distribute_cuda(rois)
    lv = (level - 2).to(torch.int32)
    perm = torch.randperm(rois.shape[0], device="cuda").to(torch.int32)
    for res in (7, 14):
        out = ops.roi_align_ml_forward(fl_cl, sc, rois, lv, res, res, 2)
        ref = torch.empty_like(out)
        for i in range(len(fl)):
            idx = torch.nonzero(lv == i).flatten()
            if len(idx):
                ref[idx] = ref_fwd(fl[i], rois[idx].contiguous(), res, res, sc[i], 2)
        gate(out, ref, "full size res %d" % res)
        gate(ops.roi_align_ml_forward(fl, sc, rois, lv, res, res, 2), ref, "NCHW row-window kernel, res %d" % res)
        sep = torch.empty_like(out)                      # the separable NCHW kernel: same summation order as channels-last
        ptrs, hs, ws, scs = ops._level_arrays(fl, sc)
        _lib.call("vosd_roialign_ml_fwd", ptrs, hs, ws, scs, len(fl), 256, res, res, 2, rois.shape[0], ops._ptr(rois), ops._ptr(lv),
                  None, ops._ptr(sep), ops._stream())
        assert torch.equal(out, sep)
        scattered = ops.roi_align_ml_forward(fl_cl, sc, rois, lv, res, res, 2, out_index=perm)
        assert torch.equal(scattered[perm.long()], out)


def test_unsupported_configurations_use_the_nchw_path(synth):
    """Adaptive grid / odd channel counts / other pooled widths: the wrapper converts to NCHW (no silent wrong path)."""
    from vosdetectron_b200 import ops
    f = torch.randn((1, 48, 40, 56), device="cuda")
    rois = torch.from_numpy(synth.random_rois(5, 50, (320, 448), 1)).cuda()
    for C, res, sr in [(48, 7, 2), (32, 6, 2), (32, 7, 0)]:
        g = f[:, :C].contiguous()
        a = ops.roi_align_forward(g.contiguous(memory_format=torch.channels_last), rois, res, res, 0.125, sr)
        assert torch.equal(a, ops.roi_align_forward(g, rois, res, res, 0.125, sr))


# ----------------------------------------------------------------------------- backward (csrc/roialign_nhwc_bwd.cuh)
@pytest.fixture(scope="module")
def ref_bwd():
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref/libref_roialign.so not built")
    lib = ctypes.CDLL(REF_SO)
    vp = ctypes.c_void_p
    lib.ROIAlignBackwardLaucher.argtypes = [vp, ctypes.c_float] + [ctypes.c_int] * 8 + [vp, vp, vp]

    def bwd(top, rois, shape, ph, pw, scale, sr):
        N, C, H, W = shape
        g = torch.zeros(shape, device=top.device)
        if rois.shape[0]:
            lib.ROIAlignBackwardLaucher(top.data_ptr(), scale, N, rois.shape[0], H, W, C, ph, pw, sr, rois.data_ptr(),
                                        g.data_ptr(), torch.cuda.current_stream().cuda_stream)
        return g
    return bwd


@pytest.mark.parametrize("res", [7, 14, 28])
def test_backward_single_level_edge_oversize_and_random_rois(ref_bwd, synth, res):
    """Channels-last backward through autograd (the reference's call form) against the reference kernel on the NCHW
    copy: same gate as the NCHW backward tests (|g - ref| <= 1e-5*|ref| + 1e-6*max|ref|: the sums are re-associated).
    The oversize RoIs exceed the 64-texel tables and take the kernel's direct path."""
    from vosdetectron_b200.modeling.roi_xfrom.roi_align.functions.roi_align import RoIAlignFunction
    N, C = 2, 64
    f = torch.from_numpy(synth.fpn_features(93, synth.COCO_BLOB, N, (3,), C)[3]).cuda()
    f_cl = f.contiguous(memory_format=torch.channels_last).requires_grad_(True)
    big = np.array([[0, 0, 0, 1343, 799], [1, 10, 20, 1300, 90], [0, 5, 5, 90, 790], [1, 300, 300, 340, 330],
                    [1, 1200, 700, 1343, 799], [0, 1330, 2, 1343.5, 40]], np.float32)
    rois = torch.from_numpy(np.concatenate([synth.edge_rois(), big, synth.random_rois(94, 300, synth.COCO_BLOB, N)])).cuda()
    out = RoIAlignFunction(res, res, 0.125, 2)(f_cl, rois)
    top = torch.randn_like(out)
    out.backward(top)
    g = f_cl.grad
    assert g.shape == f.shape and g.is_contiguous(memory_format=torch.channels_last)
    gate(g, ref_bwd(top, rois, tuple(f.shape), res, res, 0.125, 2), "channels-last backward, res %d" % res)


def test_backward_full_size_multilevel(ref_bwd, synth):
    """BASELINE config 4 shapes: 1024 RoIs x 256 ch 7x7 and 256 RoIs 14x14 over P2-P5 of 2 frames, channels-last
    gradients through roi_align_multilevel's autograd, with a scattered out_index."""
    from vosdetectron_b200 import ops
    from vosdetectron_b200.modeling.roi_xfrom.roi_align.functions.roi_align import roi_align_multilevel
    N = 2
    feats = synth.fpn_features(2100, synth.COCO_BLOB, N, synth.ROI_LEVELS, 256)
    fl = [torch.from_numpy(feats[l]).cuda() for l in synth.ROI_LEVELS]
    sc = [1.0 / 2 ** l for l in synth.ROI_LEVELS]
    for res, R in ((7, 1024), (14, 256)):
        rois = torch.from_numpy(synth.random_rois(2101 + res, R, synth.COCO_BLOB, N)).cuda()
        level, _, order, restore = ops.distribute_cuda(rois)
        lv = (level - 2).to(torch.int32)
        perm = torch.randperm(R, device="cuda").to(torch.int32)
        fl_cl = [f.contiguous(memory_format=torch.channels_last).requires_grad_(True) for f in fl]
        out = roi_align_multilevel(fl_cl, sc, rois, lv, res, res, 2, out_index=perm)
        top = torch.randn_like(out)
        out.backward(top)
        for i, f in enumerate(fl):
            idx = torch.nonzero(lv == i).flatten()
            ref = ref_bwd(top[perm[idx].long()].contiguous(), rois[idx].contiguous(), tuple(f.shape), res, res, sc[i], 2)
            g = fl_cl[i].grad
            assert g.is_contiguous(memory_format=torch.channels_last)
            gate(g, ref, "level %d res %d" % (i, res))
        nchw = ops.roi_align_ml_backward(top, [tuple(f.shape) for f in fl], sc, rois, lv, res, res, 2, out_index=perm)
        for i in range(len(fl)):
            gate(fl_cl[i].grad, nchw[i], "vs the NCHW backward of this library, level %d" % i)


def test_backward_unsupported_heads_convert(synth):
    """Adaptive grid / pooled sizes that are not multiples of 7 / odd channel counts: accumulated in NCHW order and
    converted -- the values of the NCHW call, channels-last strides."""
    from vosdetectron_b200 import ops
    rois = torch.from_numpy(synth.random_rois(7, 40, (320, 448), 1)).cuda()
    for C, res, sr in [(48, 7, 2), (32, 6, 2), (32, 7, 0)]:
        top = torch.randn((40, C, res, res), device="cuda")
        a = ops.roi_align_backward(top, rois, (1, C, 40, 56), res, res, 0.125, sr, channels_last=True)
        b = ops.roi_align_backward(top, rois, (1, C, 40, 56), res, res, 0.125, sr)
        assert a.is_contiguous(memory_format=torch.channels_last)
        gate(a, b, "converted backward")                   # two atomic accumulations: same gate, not bit-equal
