"""Drop-in replacement for the reference's compiled extension module `boundary_max_pooling_cuda`
(built by the reference's setup.py:15-18 from AFSD/prop_pooling/*.cpp/*.cu; pybind surface
boundary_max_pooling_cuda.cpp:52-55).  Put this directory on PYTHONPATH (or call
`opental_b200.install_shim()`) and `AFSD/prop_pooling/boundary_pooling_op.py` works unchanged.

forward(input[B,C,T], segments[B,K,4]) -> Tensor[B,C,K]
backward(grad_output[B,C,K], input, segments) -> Tensor[B,C,T]

Errors: non-CUDA / non-contiguous tensors raise RuntimeError like the reference's TORCH_CHECKs
(boundary_max_pooling_cuda.cpp:4-6).  Additionally validated (the reference assumes them): dtype equality,
segments.size(0) == B, last dim 4, even C.

The reference backward takes the time extent from grad_output (boundary_max_pooling_kernel.cu:121); that
behaviour is selected with OPENTAL_B200_BMP_COMPAT=1 (default 0 = mathematically correct gradient).
"""
import os

from opental_b200 import ops as _ops


def _compat() -> bool:
    return os.environ.get("OPENTAL_B200_BMP_COMPAT", "0") not in ("0", "", "false", "False")


def forward(input, segments):
    return _ops.bmp_forward(input, segments)


def backward(grad_output, input, segments):
    return _ops.bmp_backward(grad_output, input, segments, compat_tscale_bug=_compat())
