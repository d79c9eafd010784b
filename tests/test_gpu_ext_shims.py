"""GPU tests of the cffi-level mirrors (`_ext.roi_align`, `_ext.flow_align`: SURVEY 8a row a16 / 8f rank 4) against the
reference kernels built into oracle/_ref, plus the round-1 review items that need a GPU:
  * a 4-byte aligned (not 16-byte aligned) top_diff through the separable backward,
  * the single-level autograd path with roi_level = None,
  * tensors on a GPU that is not the current device (needs >= 2 GPUs),
  * non-finite texels through the default forward (pins the behaviour: as the reference kernel).
Tolerances as in test_gpu_roialign.py: forward |out - ref| <= 1e-5*|ref| + 1e-6*max|ref|, backward the same on grads;
FlowAlign forward bit-identical."""
import ctypes
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_RA = os.path.join(ROOT, "oracle", "_ref", "libref_roialign.so")
REF_FA = os.path.join(ROOT, "oracle", "_ref", "libref_flowalign.so")
vp = ctypes.c_void_p


def gate(out, ref, what=""):
    tol = 1e-5 * ref.abs() + 1e-6 * float(ref.abs().max())
    err = (out - ref).abs()
    assert bool((err <= tol).all()), "%s: max err %g (max|ref| %g)" % (what, float(err.max()), float(ref.abs().max()))


@pytest.fixture(scope="module")
def ref_ra():
    if not os.path.exists(REF_RA):
        pytest.skip("oracle/_ref/libref_roialign.so not built")
    lib = ctypes.CDLL(REF_RA)
    lib.ROIAlignForwardLaucher.argtypes = [vp, ctypes.c_float] + [ctypes.c_int] * 7 + [vp, vp, vp]
    lib.ROIAlignBackwardLaucher.argtypes = [vp, ctypes.c_float] + [ctypes.c_int] * 8 + [vp, vp, vp]
    return lib


def _st():
    return torch.cuda.current_stream().cuda_stream


def test_ext_roi_align_forward_backward_cuda(ref_ra, synth):
    """roi_align_forward_cuda / roi_align_backward_cuda with the THC shim's calling convention (caller allocates and
    zero-fills; returns 1, or 0 for rois that are not (R,5)) against the reference launchers on the same GPU."""
    import vosdetectron_b200
    vosdetectron_b200.install_reference_aliases()
    import sys
    ra = sys.modules['modeling.roi_xfrom.roi_align._ext.roi_align']
    for lvl, C, res in ((2, 32, 7), (3, 64, 14), (5, 8, 7)):
        f = torch.from_numpy(synth.fpn_features(300 + lvl, synth.COCO_BLOB, 2, (lvl,), C)[lvl]).cuda()
        rois = torch.from_numpy(np.concatenate([synth.random_rois(301 + lvl, 90, synth.COCO_BLOB, 2), synth.edge_rois()])).cuda()
        scale = 1.0 / 2 ** lvl
        N, _, H, W = f.shape
        out = torch.zeros((rois.shape[0], C, res, res), device="cuda")
        assert ra.roi_align_forward_cuda(res, res, scale, 2, f, rois, out) == 1
        ref = torch.zeros_like(out)
        ref_ra.ROIAlignForwardLaucher(f.data_ptr(), scale, rois.shape[0], H, W, C, res, res, 2, rois.data_ptr(), ref.data_ptr(), _st())
        gate(out, ref, "ext fwd lvl %d" % lvl)
        top = torch.randn_like(out)
        grad = torch.zeros_like(f)
        assert ra.roi_align_backward_cuda(res, res, scale, 2, top, rois, grad) == 1
        gref = torch.zeros_like(f)
        ref_ra.ROIAlignBackwardLaucher(top.data_ptr(), scale, N, rois.shape[0], H, W, C, res, res, 2, rois.data_ptr(), gref.data_ptr(), _st())
        gate(grad, gref, "ext bwd lvl %d" % lvl)
        # accumulate semantics: a second call adds onto the caller's buffer (the shim passes zero_init = 0)
        assert ra.roi_align_backward_cuda(res, res, scale, 2, top, rois, grad) == 1
        gate(grad, 2 * gref, "ext bwd accumulates")
    bad = torch.zeros((3, 4), device="cuda")
    assert ra.roi_align_forward_cuda(7, 7, 0.25, 2, f, bad, torch.zeros((3, 8, 7, 7), device="cuda")) == 0
    assert ra.roi_align_backward_cuda(7, 7, 0.25, 2, torch.zeros((3, 8, 7, 7), device="cuda"), bad, torch.zeros_like(f)) == 0


def test_ext_flow_align_forward_backward_cuda(synth):
    if not os.path.exists(REF_FA):
        pytest.skip("oracle/_ref/libref_flowalign.so not built")
    import vosdetectron_b200
    vosdetectron_b200.install_reference_aliases()
    import sys
    fa = sys.modules['vos_model.flow_align._ext.flow_align']
    lib = ctypes.CDLL(REF_FA)
    lib.FlowAlignForward.argtypes = [ctypes.c_int] * 4 + [vp, vp, vp, vp]
    lib.FlowAlignBackward.argtypes = [ctypes.c_int] * 4 + [vp, vp, vp, vp, vp, vp]
    N, C, H, W = 2, 24, 24, 42
    rs = np.random.RandomState(5)
    f = torch.from_numpy(rs.standard_normal((N, C, H, W)).astype(np.float32)).cuda()
    fl = torch.from_numpy(synth.flow_field(6, N, H, W, "smooth", 2.0)).cuda()
    g = torch.from_numpy(rs.standard_normal((N, C, H, W)).astype(np.float32)).cuda()
    top = torch.empty_like(f)
    assert fa.flow_align_forward_cuda(f, fl, top) == 1
    ref = torch.empty_like(f)
    lib.FlowAlignForward(N, H, W, C, f.data_ptr(), fl.data_ptr(), ref.data_ptr(), _st())
    assert torch.equal(top, ref)
    gf, gfl = torch.zeros_like(f), torch.zeros_like(fl)
    assert fa.flow_align_backward_cuda(g, f, fl, gf, gfl) == 1
    rf, rfl = torch.zeros_like(f), torch.zeros_like(fl)
    lib.FlowAlignBackward(N, H, W, C, g.data_ptr(), f.data_ptr(), fl.data_ptr(), rf.data_ptr(), rfl.data_ptr(), _st())
    gate(gf, rf, "ext flow bwd (features)")
    err = (gfl - rfl).abs()
    assert bool((err <= 1e-4 * rfl.abs() + 2e-5 * float(rfl.abs().max())).all()), float(err.max())


@pytest.mark.parametrize("res,C", [(7, 32), (7, 64), (14, 32)])
def test_backward_with_unaligned_top_diff(ref_ra, synth, res, C):
    """top_diff that is only 4-byte aligned (a contiguous view with an odd storage offset): the separable backward
    must not take its 16-byte cp.async fetch."""
    from vosdetectron_b200 import ops
    f = torch.from_numpy(synth.fpn_features(41, synth.COCO_BLOB, 2, (3,), C)[3]).cuda()
    rois = torch.from_numpy(synth.random_rois(42, 130, synth.COCO_BLOB, 2)).cuda()
    n = rois.shape[0] * C * res * res
    store = torch.randn(n + 1, device="cuda")
    g = store[1:].view(rois.shape[0], C, res, res)
    assert g.is_contiguous() and g.data_ptr() % 16 == 4
    mine = ops.roi_align_backward(g, rois, f.shape, res, res, 0.125, 2)
    N, _, H, W = f.shape
    gref = torch.zeros_like(f)
    gc = g.clone()
    ref_ra.ROIAlignBackwardLaucher(gc.data_ptr(), 0.125, N, rois.shape[0], H, W, C, res, res, 2, rois.data_ptr(), gref.data_ptr(), _st())
    gate(mine, gref, "unaligned top_diff")


def test_single_level_multilevel_function_without_roi_level(synth):
    """roi_align_multilevel with one level and roi_level = None: forward AND backward."""
    from vosdetectron_b200 import ops
    from vosdetectron_b200.modeling.roi_xfrom.roi_align.functions.roi_align import roi_align_multilevel
    f = torch.from_numpy(synth.fpn_features(51, synth.COCO_BLOB, 1, (4,), 32)[4]).cuda().requires_grad_(True)
    rois = torch.from_numpy(synth.random_rois(52, 60, synth.COCO_BLOB, 1)).cuda()
    out = roi_align_multilevel([f], [1.0 / 16], rois, None, 7, 7, 2)
    w = torch.randn_like(out)
    (out * w).sum().backward()
    ref = ops.roi_align_backward(w, rois, f.shape, 7, 7, 1.0 / 16, 2)
    gate(f.grad, ref, "single-level ML backward")


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_tensors_on_a_non_current_device(synth):
    """Inputs on cuda:1 while cuda:0 is current: the call runs on cuda:1 behind its producer and leaves the caller's
    current device alone."""
    from vosdetectron_b200 import ops
    torch.cuda.set_device(0)
    f0 = torch.from_numpy(synth.fpn_features(61, synth.COCO_BLOB, 1, (3,), 32)[3])
    rois0 = torch.from_numpy(synth.random_rois(62, 80, synth.COCO_BLOB, 1))
    ref = ops.roi_align_forward(f0.cuda(0), rois0.cuda(0), 7, 7, 0.125, 2)
    out = ops.roi_align_forward(f0.to("cuda:1"), rois0.to("cuda:1"), 7, 7, 0.125, 2)
    assert torch.cuda.current_device() == 0 and out.device.index == 1
    assert torch.equal(out.cpu(), ref.cpu())
    g = ops.roi_align_backward(out, rois0.to("cuda:1"), f0.shape, 7, 7, 0.125, 2)
    assert g.device.index == 1 and torch.cuda.current_device() == 0
    keep, num = ops.nms_cuda(torch.from_numpy(synth.clustered_dets(63, 500)).to("cuda:1"), 0.5)
    assert keep.device.index == 1 and torch.cuda.current_device() == 0


@pytest.mark.parametrize("res", [7, 14])
def test_non_finite_texels_follow_the_reference(ref_ra, synth, res):
    """NaN / Inf texels: the default forward multiplies exactly the taps the reference multiplies (the four taps of
    every valid sample, zero weights included), so the SET of non-finite outputs is the reference's and all finite
    outputs stay within the forward gate."""
    from vosdetectron_b200 import ops
    f = torch.from_numpy(synth.fpn_features(71, synth.COCO_BLOB, 1, (3,), 32)[3]).cuda()
    rs = np.random.RandomState(72)
    ys, xs = rs.randint(0, f.shape[2], 40), rs.randint(0, f.shape[3], 40)
    for i, (y, x) in enumerate(zip(ys, xs)):
        f[0, i % 32, y, x] = float("nan") if i % 2 else float("inf")
    rois = torch.from_numpy(synth.random_rois(73, 200, synth.COCO_BLOB, 1)).cuda()
    out = ops.roi_align_forward(f, rois, res, res, 0.125, 2)
    _, C, H, W = f.shape
    ref = torch.zeros_like(out)
    ref_ra.ROIAlignForwardLaucher(f.data_ptr(), 0.125, rois.shape[0], H, W, C, res, res, 2, rois.data_ptr(), ref.data_ptr(), _st())
    bad_ref = ~torch.isfinite(ref)
    assert int(bad_ref.sum()) > 0
    assert torch.equal(~torch.isfinite(out), bad_ref)
    ok = ~bad_ref
    gate(torch.where(ok, out, torch.zeros_like(out)), torch.where(ok, ref, torch.zeros_like(ref)), "finite outputs")


def test_channels_last_inputs_the_tma_kernel_does_not_take_use_the_nchw_kernel(synth):
    """Channels-last maps the channels-last entry point answers VOSD_ERR_UNSUPPORTED for (five levels; a base that is
    not 16-byte aligned) still come back from the GPU, through the NCHW kernel."""
    from vosdetectron_b200 import ops
    rois = torch.from_numpy(synth.random_rois(61, 80, synth.COCO_BLOB, 1)).cuda()
    f = torch.from_numpy(synth.fpn_features(62, synth.COCO_BLOB, 1, (4,), 32)[4]).cuda()
    ref = ops.roi_align_forward(f, rois, 7, 7, 1.0 / 16, 2)
    # (a) five levels, every one channels-last; all RoIs read level 3
    cl = [f.contiguous(memory_format=torch.channels_last) for _ in range(5)]
    lv = torch.full((rois.shape[0],), 3, dtype=torch.int32, device="cuda")
    out = ops.roi_align_ml_forward(cl, [1.0 / 16] * 5, rois, lv, 7, 7, 2)
    gate(out, ref, "five channels-last levels")
    # (b) channels-last view whose base is 4-byte aligned only
    N, C, H, W = f.shape
    buf = torch.empty(N * C * H * W + 1, dtype=torch.float32, device="cuda")
    odd = buf[1:].view(N, H, W, C).permute(0, 3, 1, 2)
    odd.copy_(f)
    assert odd.data_ptr() % 16 != 0 and odd.is_contiguous(memory_format=torch.channels_last)
    gate(ops.roi_align_forward(odd, rois, 7, 7, 1.0 / 16, 2), ref, "unaligned channels-last base")


def test_flow_align_ml_backward_validates_its_lists():
    from vosdetectron_b200 import ops
    f = [torch.randn(1, 32, 16, 24, device="cuda"), torch.randn(1, 32, 8, 12, device="cuda")]
    fl = [torch.zeros(1, 2, 16, 24, device="cuda"), torch.zeros(1, 2, 8, 12, device="cuda")]
    g = [torch.ones_like(t) for t in f]
    ops.flow_align_ml_backward(g, f, fl)
    with pytest.raises(ValueError):
        ops.flow_align_ml_backward(g[:1], f, fl)                  # short grad list
    with pytest.raises(ValueError):
        ops.flow_align_ml_backward([g[0], g[0]], f, fl)           # grad of the wrong level
    with pytest.raises(ValueError):
        ops.flow_align_ml_backward(g, [f[0], torch.randn(1, 16, 8, 12, device="cuda")], fl)   # channel mismatch
