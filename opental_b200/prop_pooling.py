"""BoundaryMaxPooling operator with the reference's Python API (AFSD/prop_pooling/boundary_pooling_op.py:7-32):
`BoundaryMaxPoolingFunction.apply(input, segments)` and the ctor-argument-free `BoundaryMaxPooling()` module.
Saves (input, segments), makes grad_output contiguous, returns (grad_input, None)."""
import torch.nn as nn
from torch.autograd import Function

from . import ops


class BoundaryMaxPoolingFunction(Function):
    # class-level switch: True reproduces the reference backward's tscale quirk (kernel.cu:121)
    compat_tscale_bug = False

    @staticmethod
    def forward(ctx, input, segments):
        output = ops.bmp_forward(input, segments)
        ctx.save_for_backward(input, segments)
        return output

    @staticmethod
    def backward(ctx, grad_output):
        if not grad_output.is_contiguous():
            grad_output = grad_output.contiguous()
        input, segments = ctx.saved_tensors
        grad_input = ops.bmp_backward(grad_output, input, segments, BoundaryMaxPoolingFunction.compat_tscale_bug)
        return grad_input, None


class BoundaryMaxPooling(nn.Module):
    def __init__(self):
        super().__init__()

    def forward(self, input, segments):
        return BoundaryMaxPoolingFunction.apply(input, segments)
