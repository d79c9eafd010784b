"""Drop-in for lib_vos/vos_model/flow_align/functions/flow_align.py:6-53.

``FlowAlignFunction.apply(features, flows)`` warps ``features (N,C,H,W)`` by ``flows (N,2,H,W)``
(x displacement, y displacement, in pixels of the feature map); differentiable w.r.t. both inputs.
CPU tensors raise NotImplementedError like the reference (:29-30): there is no CPU path.
"""
import torch
from torch.autograd import Function

from .... import ops


class FlowAlignFunction(Function):
    @staticmethod
    def forward(ctx, features, flows):
        if not features.is_cuda:
            raise NotImplementedError
        ctx.save_for_backward(features, flows)
        return ops.flow_align_forward(features, flows)

    @staticmethod
    def backward(ctx, grad_output):
        if not (ctx.needs_input_grad[0] or ctx.needs_input_grad[1]):
            return None, None
        features, flows = ctx.saved_tensors
        if not grad_output.is_cuda:
            raise NotImplementedError("FlowAlign backward needs CUDA tensors")
        # a flow that does not require grad (a frozen estimator's output) skips the flow-gradient half of the kernel
        grad_feature, grad_flow = ops.flow_align_backward(grad_output.contiguous(), features, flows,
                                                          want_flow_grad=ctx.needs_input_grad[1])
        return (grad_feature if ctx.needs_input_grad[0] else None), grad_flow


class _FlowAlignML(Function):
    """Every FPN level in one launch (fast variant for the loop at vos_model_builder.py:329-335)."""

    @staticmethod
    def forward(ctx, num_levels, *tensors):
        feats, flows = tensors[:num_levels], tensors[num_levels:]
        ctx.num_levels = num_levels
        ctx.save_for_backward(*tensors)
        return tuple(ops.flow_align_ml_forward(feats, flows))

    @staticmethod
    def backward(ctx, *grads):
        L = ctx.num_levels
        feats, flows = ctx.saved_tensors[:L], ctx.saved_tensors[L:]
        grads = [g.contiguous() if g is not None else torch.zeros_like(f) for g, f in zip(grads, feats)]
        want_flow = any(ctx.needs_input_grad[1 + L:])
        gfs, gfls = ops.flow_align_ml_backward(grads, feats, flows, want_flow_grad=want_flow)
        return (None,) + tuple(gfs) + (tuple(gfls) if gfls is not None else (None,) * L)


def flow_align_multilevel(level_features, level_flows):
    """[features_k], [flows_k] -> [warped_k], one launch for all levels."""
    if len(level_features) != len(level_flows):
        raise ValueError("one flow per level")
    return list(_FlowAlignML.apply(len(level_features), *level_features, *level_flows))
