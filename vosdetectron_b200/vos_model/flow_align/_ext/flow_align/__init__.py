"""The two entry points of the reference's cffi module ``_ext.flow_align`` (flow_align_forward_cuda /
flow_align_backward_cuda, lib_vos/vos_model/flow_align/src/flow_align_cuda.h:1-3; shim flow_align_cuda.c:7-43), on top
of the C ABI of this package.  The caller allocates ``top`` and allocates + zero-fills ``bottom_grad`` / ``flow_grad``
(functions/flow_align.py:27,41-43); returns 1 like the launchers (flow_align_cuda_kernel.cu:137,156)."""
import ctypes

import torch

from ..... import _lib

__all__ = ["flow_align_forward_cuda", "flow_align_backward_cuda"]


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def _check(*tensors):
    for t in tensors:
        if not t.is_cuda:
            raise NotImplementedError("CPU tensor: there is no CPU path")
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise ValueError("float32 contiguous CUDA tensors expected (THCudaTensor)")


def _stream(t):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def flow_align_forward_cuda(bottom, flow, top):
    _check(bottom, flow, top)
    N, C, H, W = bottom.size()                                      # flow_align_cuda.c:14-17
    if tuple(flow.size()) != (N, 2, H, W) or tuple(top.size()) != (N, C, H, W):
        raise ValueError("flow must be (N,2,H,W) and top (N,C,H,W)")     # the reference would read out of bounds
    with torch.cuda.device(bottom.device):                     # the caller's current device is restored on exit
        _lib.call("vosd_flow_align_fwd", N, H, W, C, _p(bottom), _p(flow), _p(top), _stream(bottom))
    return 1


def flow_align_backward_cuda(top_grad, bottom, flow, bottom_grad, flow_grad):
    _check(top_grad, bottom, flow, bottom_grad, flow_grad)
    N, C, H, W = top_grad.size()                                    # flow_align_cuda.c:33-36
    if tuple(bottom.size()) != (N, C, H, W) or tuple(flow.size()) != (N, 2, H, W) or \
            tuple(bottom_grad.size()) != (N, C, H, W) or tuple(flow_grad.size()) != (N, 2, H, W):
        raise ValueError("inconsistent shapes")
    with torch.cuda.device(top_grad.device):                     # the caller's current device is restored on exit
        _lib.call("vosd_flow_align_bwd", N, H, W, C, _p(top_grad), _p(bottom), _p(flow), _p(bottom_grad), _p(flow_grad), 0,
                  _stream(top_grad))                                    # 0: the caller zero-filled both gradients
    return 1
