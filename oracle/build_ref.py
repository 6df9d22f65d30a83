"""Compile the reference's OWN CUDA extension, unmodified, from the sources where they lie under /root/reference
into oracle/_ref/ (git-ignored, travels to the GPU box)  —  TEST INFRASTRUCTURE.

The reference .cu targets torch 1.9 (THC era); on torch 2.11 it needs exactly two names that no longer exist
(at::cuda::getCurrentCUDAStream's header and THCudaCheck).  They are supplied by a force-included shim header and a
stub <THC/THCAtomics.cuh>; the reference sources themselves are neither copied nor edited.
"""
from __future__ import annotations

import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = os.environ.get("OPENTAL_REFERENCE", "/root/reference")
NAME = "ref_boundary_max_pooling_cuda"

SHIM = """#pragma once
#include <ATen/cuda/CUDAContext.h>
#include <ATen/cuda/Atomic.cuh>
#include <c10/cuda/CUDAException.h>
#define THCudaCheck(x) C10_CUDA_CHECK(x)
"""


def so_path() -> str | None:
    if os.path.isdir(OUT):
        for f in os.listdir(OUT):
            if f.startswith(NAME) and f.endswith(".so"):
                return os.path.join(OUT, f)
    return None


def build() -> str | None:
    """Build if the reference is present (build container); otherwise return the prebuilt .so or None."""
    src_dir = os.path.join(REF, "AFSD", "prop_pooling")
    if not os.path.isdir(src_dir):
        return so_path()
    if so_path():
        return so_path()
    from torch.utils.cpp_extension import load

    os.makedirs(os.path.join(OUT, "shim", "THC"), exist_ok=True)
    with open(os.path.join(OUT, "shim", "ref_shim.h"), "w") as fh:
        fh.write(SHIM)
    with open(os.path.join(OUT, "shim", "THC", "THCAtomics.cuh"), "w") as fh:
        fh.write("#pragma once\n#include <ATen/cuda/Atomic.cuh>\n")
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
    load(name=NAME,
         sources=[os.path.join(src_dir, "boundary_max_pooling_cuda.cpp"), os.path.join(src_dir, "boundary_max_pooling_kernel.cu")],
         extra_include_paths=[os.path.join(OUT, "shim")],
         extra_cuda_cflags=["-include", os.path.join(OUT, "shim", "ref_shim.h"), "-gencode", "arch=compute_100a,code=sm_100a"],
         build_directory=OUT, is_python_module=False, verbose=False)
    return so_path()


def load_module():
    """Import the compiled reference extension (exposes forward / backward like the reference's module)."""
    import torch  # noqa: F401  (must be loaded first: the .so links against libtorch)
    p = so_path()
    if p is None:
        return None
    spec = importlib.util.spec_from_file_location(NAME, p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    print(build())
