"""GPU parity: FlowAlign forward / backward through the C ABI against
  (i)  the reference kernel itself (oracle/_ref/libref_flowalign.so = the unmodified
       lib_vos/vos_model/flow_align/src/flow_align_cuda_kernel.cu built for sm_100a) on the same GPU;
  (ii) the CPU oracle (oracle/oracle.c: orc_flow_align_fwd/bwd).
Tolerances:
  * forward: BIT-IDENTICAL to the reference kernel and to the oracle (0 ulp; NaN where the reference has NaN);
    the opt-in fp32 variant (vosd_debug_flow_align_fast) |out - ref| <= 1e-5*|ref| + 1e-6*max|ref|;
  * backward: same addends, different summation order (pre-summed shared texels / channel chunks; the
    reference's atomics are unordered too): |g - ref| <= 1e-5*|ref| + 2e-6*max|ref|."""
import ctypes
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libref_flowalign.so")


@pytest.fixture(scope="module")
def refk():
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref/libref_flowalign.so not built")
    lib = ctypes.CDLL(REF_SO)
    vp = ctypes.c_void_p
    lib.FlowAlignForward.argtypes = [ctypes.c_int] * 4 + [vp, vp, vp, vp]
    lib.FlowAlignBackward.argtypes = [ctypes.c_int] * 4 + [vp, vp, vp, vp, vp, vp]

    class R:
        @staticmethod
        def fwd(f, fl):
            N, C, H, W = f.shape
            out = torch.empty_like(f)
            if f.numel():
                lib.FlowAlignForward(N, H, W, C, f.data_ptr(), fl.data_ptr(), out.data_ptr(),
                                     torch.cuda.current_stream().cuda_stream)
            return out

        @staticmethod
        def bwd(g, f, fl):
            N, C, H, W = f.shape
            gf, gfl = torch.zeros_like(f), torch.zeros_like(fl)
            if f.numel():
                lib.FlowAlignBackward(N, H, W, C, g.data_ptr(), f.data_ptr(), fl.data_ptr(), gf.data_ptr(),
                                      gfl.data_ptr(), torch.cuda.current_stream().cuda_stream)
            return gf, gfl
    return R


@pytest.fixture(params=["alu", "cvt"])
def exact_mode(request):
    """Both bit-exact arithmetic variants of the kernels: 2 = float->double widening on the ALU pipe (default),
    0 = the reference's expression as written (F2F conversions)."""
    from vosdetectron_b200 import _lib
    old = _lib.load().vosd_debug_flow_align_fast({"alu": 2, "cvt": 0}[request.param])
    yield request.param
    _lib.load().vosd_debug_flow_align_fast(old)


def same_bits(a, b):
    """Bitwise equality, except that any NaN matches any NaN (payloads are not part of the contract)."""
    a, b = a.contiguous(), b.contiguous()
    both_nan = torch.isnan(a) & torch.isnan(b)
    return bool(((a.view(torch.int32) == b.view(torch.int32)) | both_nan).all())


def close(a, b, rtol=1e-5, atol_rel=2e-6):
    scale = float(b.abs().max()) if b.numel() else 0.0
    err = (a - b).abs()
    bound = rtol * b.abs() + atol_rel * scale
    return bool((err <= bound).all()), float(err.max()) if err.numel() else 0.0, scale


def make(synth, seed, N, C, H, W, kind, magnitude=2.0):
    rs = np.random.RandomState(seed)
    f = torch.from_numpy(rs.standard_normal((N, C, H, W)).astype(np.float32)).cuda()
    fl = torch.from_numpy(synth.flow_field(seed + 1, N, H, W, kind, magnitude)).cuda()
    g = torch.from_numpy(rs.standard_normal((N, C, H, W)).astype(np.float32)).cuda()
    return f, fl, g


CASES = [
    # N, C, H, W, kind, magnitude
    (2, 8, 24, 42, "smooth", 2.0),
    (1, 37, 13, 21, "smooth", 1.0),       # odd channel count, ragged tiles (P6 shape)
    (2, 16, 48, 84, "noise", 3.0),        # every pixel its own geometry, many out of range
    (1, 64, 17, 33, "noise", 0.4),        # 33 columns: a one-lane second tile
    (1, 12, 20, 40, "zero", 0.0),
    (1, 12, 20, 40, "shift", 0.0),
    (3, 5, 5, 7, "noise", 30.0),          # almost everything out of range
    (1, 256, 24, 42, "smooth", 2.0),      # P5 of a DAVIS frame, full channel count
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "%dx%dx%dx%d-%s" % c[:5])
def test_forward_bit_identical(case, refk, orc, synth, exact_mode):
    from vosdetectron_b200 import ops
    N, C, H, W, kind, mag = case
    f, fl, _ = make(synth, 11, N, C, H, W, kind, mag)
    out = ops.flow_align_forward(f, fl)
    ref = refk.fwd(f, fl)
    assert same_bits(out, ref), "max |diff| %g" % float((out - ref).abs().max())
    o = torch.from_numpy(orc.flow_align_forward(f.cpu().numpy(), fl.cpu().numpy())).cuda()
    assert same_bits(out, o), "oracle: max |diff| %g" % float((out - o).abs().max())


@pytest.mark.parametrize("case", CASES, ids=lambda c: "%dx%dx%dx%d-%s" % c[:5])
def test_backward_vs_reference_kernel_and_oracle(case, refk, orc, synth, exact_mode):
    from vosdetectron_b200 import ops
    N, C, H, W, kind, mag = case
    f, fl, g = make(synth, 13, N, C, H, W, kind, mag)
    gf, gfl = ops.flow_align_backward(g, f, fl)
    rf, rfl = refk.bwd(g, f, fl)
    of, ofl = (torch.from_numpy(a).cuda() for a in orc.flow_align_backward(g.cpu().numpy(), f.cpu().numpy(),
                                                                          fl.cpu().numpy()))
    for name, mine, ref, o in (("feature", gf, rf, of), ("flow", gfl, rfl, ofl)):
        ok, err, scale = close(mine, ref)
        assert ok, "%s grad vs reference kernel: max err %g (scale %g)" % (name, err, scale)
        ok, err, scale = close(mine, o)
        assert ok, "%s grad vs oracle: max err %g (scale %g)" % (name, err, scale)
        # where nothing flows back both must be exactly zero
        assert bool(((ref == 0) == (mine == 0)).all()) or float((mine[ref == 0]).abs().max()) == 0.0


def test_backward_without_flow_gradient(refk, synth):
    """flowdiff == NULL: same feature gradient (the tap loads / dx / dy arithmetic are compiled out)."""
    from vosdetectron_b200 import ops
    from vosdetectron_b200.vos_model.flow_align.functions.flow_align import FlowAlignFunction
    f, fl, g = make(synth, 61, 2, 24, 40, 56, "smooth", 2.0)
    gf, gfl = ops.flow_align_backward(g, f, fl, want_flow_grad=False)
    assert gfl is None
    rf, _ = refk.bwd(g, f, fl)
    ok, err, scale = close(gf, rf)
    assert ok, (err, scale)
    gfs, gfls = ops.flow_align_ml_backward([g], [f], [fl], want_flow_grad=False)
    assert gfls is None and close(gfs[0], rf)[0]
    f2 = f.clone().requires_grad_(True)
    out = FlowAlignFunction.apply(f2, fl)              # flow does not require grad
    out.backward(g)
    assert close(f2.grad, rf)[0]


def test_forward_fast_variant_within_1e5(refk, synth):
    from vosdetectron_b200 import _lib, ops
    f, fl, _ = make(synth, 17, 2, 32, 48, 84, "smooth", 2.0)
    old = _lib.load().vosd_debug_flow_align_fast(1)
    try:
        out = ops.flow_align_forward(f, fl)
    finally:
        _lib.load().vosd_debug_flow_align_fast(old)
    ok, err, scale = close(out, refk.fwd(f, fl), 1e-5, 1e-6)
    assert ok, "max err %g (scale %g)" % (err, scale)


def test_nan_and_inf_flow_follow_the_reference(refk, synth, exact_mode):
    from vosdetectron_b200 import ops
    f, fl, g = make(synth, 19, 1, 6, 12, 16, "smooth", 1.0)
    fl[0, 0, 3, 4] = float("nan")
    fl[0, 1, 5, 6] = float("nan")
    fl[0, 0, 7, 8] = float("inf")
    fl[0, 1, 2, 2] = float("-inf")
    out, ref = ops.flow_align_forward(f, fl), refk.fwd(f, fl)
    assert same_bits(out, ref)
    assert bool(torch.isnan(out[0, :, 3, 4]).all()) and bool((out[0, :, 7, 8] == 0).all())
    gf, gfl = ops.flow_align_backward(g, f, fl)
    rf, rfl = refk.bwd(g, f, fl)
    assert bool((torch.isnan(gf) == torch.isnan(rf)).all()) and bool((torch.isnan(gfl) == torch.isnan(rfl)).all())
    fin = ~torch.isnan(rf)
    assert close(gf[fin], rf[fin])[0]


def test_non_finite_and_denormal_features_are_bit_identical(refk, synth, exact_mode):
    """The ALU widening is exact for zero / denormal taps; Inf / NaN taps take the reference expression."""
    from vosdetectron_b200 import ops
    f, fl, g = make(synth, 23, 1, 8, 16, 40, "smooth", 1.0)
    rs = np.random.RandomState(5)
    idx = rs.randint(0, f.numel(), size=200)
    flat = f.view(-1)
    flat[idx[:50]] = float("inf")
    flat[idx[50:80]] = float("-inf")
    flat[idx[80:120]] = float("nan")
    flat[idx[120:160]] = 1e-42          # denormal
    flat[idx[160:180]] = -0.0
    flat[idx[180:]] = 3e38              # finite, but the 4-tap sum can overflow
    g.view(-1)[rs.randint(0, g.numel(), size=20)] = float("inf")
    g.view(-1)[rs.randint(0, g.numel(), size=20)] = 2e-44
    assert same_bits(ops.flow_align_forward(f, fl), refk.fwd(f, fl))
    gf, gfl = ops.flow_align_backward(g, f, fl)
    rf, rfl = refk.bwd(g, f, fl)
    assert bool((torch.isnan(gf) == torch.isnan(rf)).all()) and bool((torch.isnan(gfl) == torch.isnan(rfl)).all())
    assert bool((torch.isinf(gf) == torch.isinf(rf)).all())
    fin = torch.isfinite(rf)
    assert close(gf[fin], rf[fin])[0]


def test_degenerate_shapes(synth):
    from vosdetectron_b200 import ops
    for shape in [(0, 4, 8, 8), (2, 0, 8, 8), (1, 3, 1, 9), (1, 3, 9, 1), (1, 2, 1, 1)]:
        N, C, H, W = shape
        f = torch.randn(shape, device="cuda")
        fl = torch.zeros((N, 2, H, W), device="cuda")
        out = ops.flow_align_forward(f, fl)
        assert out.shape == f.shape and float(out.abs().sum()) == 0.0     # H-1 == 0 or W-1 == 0: all out of range
        gf, gfl = ops.flow_align_backward(torch.ones_like(f), f, fl)
        assert float(gf.abs().sum()) == 0.0 and float(gfl.abs().sum()) == 0.0
    # NaN flow on a one-row map: the reference would read past the plane; here it is refused (zeros)
    f = torch.randn((1, 2, 1, 8), device="cuda")
    fl = torch.full((1, 2, 1, 8), float("nan"), device="cuda")
    assert float(ops.flow_align_forward(f, fl).abs().sum()) == 0.0


def test_multilevel_launch_equals_per_level(synth):
    from vosdetectron_b200 import ops
    shapes = [(12, 21), (24, 42), (48, 84), (96, 168), (0, 5)]
    N, C = 2, 24
    feats, flows, grads = [], [], []
    for i, (H, W) in enumerate(shapes):
        f, fl, g = make(synth, 30 + i, N, C, H, W, "smooth" if H else "zero", 1.5)
        feats.append(f), flows.append(fl), grads.append(g)
    outs = ops.flow_align_ml_forward(feats, flows)
    gfs, gfls = ops.flow_align_ml_backward(grads, feats, flows)
    for f, fl, g, o, gf, gfl in zip(feats, flows, grads, outs, gfs, gfls):
        assert same_bits(o, ops.flow_align_forward(f, fl))
        a, b = ops.flow_align_backward(g, f, fl)
        assert close(gf, a)[0] and close(gfl, b)[0]


def test_autograd_function_and_module(orc, synth):
    import vosdetectron_b200
    vosdetectron_b200.install_reference_aliases()
    from vos_model.flow_align.functions.flow_align import FlowAlignFunction      # reference import path
    from vos_model.flow_align.modules.flow_align import FlowAlign
    f, fl, g = make(synth, 41, 1, 6, 16, 24, "smooth", 1.0)
    f.requires_grad_(True), fl.requires_grad_(True)
    out = FlowAlignFunction.apply(f, fl)
    out.backward(g)
    of, ofl = orc.flow_align_backward(g.cpu().numpy(), f.detach().cpu().numpy(), fl.detach().cpu().numpy())
    assert close(f.grad, torch.from_numpy(of).cuda())[0] and close(fl.grad, torch.from_numpy(ofl).cuda())[0]
    # module: full-resolution flow -> fixed strided mean conv (scale**3 weights) -> warp
    m = FlowAlign(0.25).cuda()
    assert not any(p.requires_grad for p in m.parameters())
    full = torch.from_numpy(synth.flow_field(43, 1, 64, 96, "smooth", 6.0)).cuda()
    warped = m(f.detach(), full)
    small = full.reshape(1, 2, 16, 4, 24, 4).sum(dim=(3, 5)) * (0.25 ** 3)
    expect = torch.from_numpy(orc.flow_align_forward(f.detach().cpu().numpy(), m.conv_flow_downsample(full).cpu().numpy()))
    assert same_bits(warped.cpu(), expect)
    assert torch.allclose(m.conv_flow_downsample(full), small, rtol=1e-5, atol=1e-6)
    with pytest.raises(NotImplementedError):
        FlowAlignFunction.apply(f.detach().cpu(), fl.detach().cpu())
    with pytest.raises(ValueError):
        FlowAlignFunction.apply(f.detach(), fl.detach()[:, :, :8])


def test_full_size_properties_and_reference(refk, synth):
    """DAVIS-shaped P2 map (1,256,192,336): bit-identical to the reference kernel; identity and integer-shift
    properties hold exactly (weights 1/0 in the reference's expression)."""
    from vosdetectron_b200 import ops
    N, C, H, W = 1, 256, 192, 336
    f, fl, g = make(synth, 51, N, C, H, W, "smooth", 2.0)
    assert same_bits(ops.flow_align_forward(f, fl), refk.fwd(f, fl))
    gf, gfl = ops.flow_align_backward(g, f, fl)
    rf, rfl = refk.bwd(g, f, fl)
    ok, err, scale = close(gf, rf)
    assert ok, (err, scale)
    ok, err, scale = close(gfl, rfl)
    assert ok, (err, scale)
    zero = torch.zeros_like(fl)
    out = ops.flow_align_forward(f, zero)
    assert torch.equal(out[:, :, :H - 1, :W - 1], f[:, :, :H - 1, :W - 1])
    assert float(out[:, :, H - 1].abs().sum()) == 0.0 and float(out[:, :, :, W - 1].abs().sum()) == 0.0
    shift = torch.from_numpy(synth.flow_field(0, N, H, W, "shift")).cuda()      # (+2, -1)
    out = ops.flow_align_forward(f, shift)
    assert torch.equal(out[:, :, 1:H, :W - 3], f[:, :, 0:H - 1, 2:W - 1])
    # adjoint identity <warp(f), g> == <f, warp^T(g)> in fp64
    lhs = float((ops.flow_align_forward(f, fl).double() * g.double()).sum())
    rhs = float((f.double() * gf.double()).sum())
    assert abs(lhs - rhs) <= 1e-5 * max(1.0, abs(lhs))
