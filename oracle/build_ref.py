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


PY_OUT = os.path.join(OUT, "reference_src")
# the Python modules of the hot path and the two configurations the bench runs
PY_FILES = ("AFSD/common/config.py", "AFSD/common/i3d_backbone.py", "AFSD/common/layers.py", "AFSD/common/segment_utils.py",
            "AFSD/prop_pooling/boundary_pooling_op.py", "AFSD/thumos14/BDNet.py", "AFSD/thumos14/multisegment_loss.py",
            "AFSD/thumos14/cls_loss.py", "AFSD/anet/BDNet.py", "AFSD/anet/multisegment_loss.py", "AFSD/anet/cls_loss.py",
            "configs/thumos14_opental_final.yaml", "configs/thumos14.yaml", "configs/anet_opental.yaml")


def build_py() -> str | None:
    """Place the reference's own Python modules of the hot path, byte for byte, under oracle/_ref/reference_src/ (git-ignored
    build output that travels to the GPU box like the compiled checker) so that `bench.py --impl reference` and the
    cpu_baseline leg can time the REFERENCE's code on the GPU box's host cores (`cpu_baseline.kind = "reference"`).  Returns the
    directory, or None when neither /root/reference nor an earlier copy exists."""
    import filecmp
    import shutil
    if not os.path.isdir(os.path.join(REF, "AFSD")):
        return PY_OUT if os.path.isfile(os.path.join(PY_OUT, PY_FILES[5])) else None
    for rel in PY_FILES:
        src, dst = os.path.join(REF, rel), os.path.join(PY_OUT, rel)
        if not os.path.isfile(src):
            continue
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if not (os.path.isfile(dst) and filecmp.cmp(src, dst, shallow=False)):
            shutil.copyfile(src, dst)
    for d in ("AFSD", "AFSD/common", "AFSD/prop_pooling", "AFSD/thumos14", "AFSD/anet"):
        src = os.path.join(REF, d, "__init__.py")
        if os.path.isfile(src):
            shutil.copyfile(src, os.path.join(PY_OUT, d, "__init__.py"))
    return PY_OUT


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
    print(build_py())
