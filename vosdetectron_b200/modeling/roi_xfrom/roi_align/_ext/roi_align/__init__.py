"""The two entry points the reference's cffi module exported (``_ext.roi_align``: roi_align_forward_cuda /
roi_align_backward_cuda, lib/modeling/roi_xfrom/roi_align/src/roi_align_cuda.h:1-5, bound by build.py:25-33), on top of
the C ABI of this package.  Same calling convention as the THC shim (roi_align_cuda.c:7-76): the CALLER allocates and
zero-fills ``output`` / ``bottom_grad`` (functions/roi_align.py:23,39-40), the functions write into them and return
1, or 0 when ``rois`` is not (R,5).  ``torch.utils.ffi`` no longer exists, so this module is what a maintainer drops in
its place."""
import ctypes

import torch

from ...... import _lib

__all__ = ["roi_align_forward_cuda", "roi_align_backward_cuda"]


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _check(*tensors):
    for t in tensors:
        if not t.is_cuda:
            raise NotImplementedError("CPU tensor: there is no CPU path")
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise ValueError("float32 contiguous CUDA tensors expected (THCudaTensor)")


def roi_align_forward_cuda(aligned_height, aligned_width, spatial_scale, sampling_ratio, features, rois, output):
    _check(features, rois, output)
    if rois.dim() != 2 or rois.size(1) != 5:
        return 0                                                    # roi_align_cuda.c:15-18
    N, C, H, W = features.size()
    with torch.cuda.device(features.device):                     # the caller's current device is restored on exit
        _lib.call("vosd_roialign_fwd", _p(features), float(spatial_scale), rois.size(0), H, W, C, int(aligned_height),
                  int(aligned_width), int(sampling_ratio), _p(rois), _p(output),
                  ctypes.c_void_p(torch.cuda.current_stream(features.device).cuda_stream))
    return 1


def roi_align_backward_cuda(aligned_height, aligned_width, spatial_scale, sampling_ratio, top_grad, rois, bottom_grad):
    _check(top_grad, rois, bottom_grad)
    if rois.dim() != 2 or rois.size(1) != 5:
        return 0                                                    # roi_align_cuda.c:51-54
    N, C, H, W = bottom_grad.size()
    with torch.cuda.device(top_grad.device):                     # the caller's current device is restored on exit
        _lib.call("vosd_roialign_bwd", _p(top_grad), float(spatial_scale), N, rois.size(0), H, W, C, int(aligned_height),
                  int(aligned_width), int(sampling_ratio), _p(rois), _p(bottom_grad), 0,      # 0: the caller zero-filled
                  ctypes.c_void_p(torch.cuda.current_stream(top_grad.device).cuda_stream))
    return 1
