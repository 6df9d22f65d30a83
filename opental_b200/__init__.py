"""opental_b200 — B200-native (sm_100a) implementation of the OpenTAL training/inference hot path.

Host side: PyTorch for device memory / streams / torch.distributed.  Compute: hand-written CUDA kernels in
libopental_b200.so behind a C ABI (include/opental_b200.h).  There is no CPU fallback.
"""
import os
import sys

__version__ = "0.1.0"


def install_shim() -> None:
    """Make `import boundary_max_pooling_cuda` resolve to the drop-in module (opental_b200/shim)."""
    d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shim")
    if d not in sys.path:
        sys.path.insert(0, d)
