"""CPU: the C-ABI library builds, loads and exports exactly what include/opental_b200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "opental_b200.h")).read()
    return sorted(set(re.findall(r"OTAL_API\s+[\w\s\*]+?\b(otal_\w+)\s*\(", text)))


def test_header_declares_symbols():
    syms = declared_symbols()
    assert "otal_bmp_forward_f32" in syms and "otal_conv_igemm_fwd" in syms and len(syms) >= 8


def test_library_exports_every_declared_symbol():
    from opental_b200 import _lib, build
    build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing


def test_python_binding_table_matches_header():
    from opental_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_abi_version_and_error_string():
    from opental_b200 import _lib
    lib = _lib.load()
    assert lib.otal_abi_version() == 1
    assert isinstance(_lib.last_error(), str)


def test_bad_arguments_return_codes_without_gpu():
    # argument validation happens before any CUDA call, so it is testable on a CPU-only box
    from opental_b200 import _lib
    lib = _lib.load()
    assert lib.otal_bmp_forward_f32(4096, 4096, 4096, 1, 3, 4, 2, None) == -1   # odd C (pointers never dereferenced)
    assert "even" in _lib.last_error()
    assert lib.otal_bmp_forward_f32(None, None, None, 1, 4, 4, 2, None) == -1   # null pointers
    assert lib.otal_conv_igemm_fwd(None, None) == -1
    assert lib.otal_bmp_forward_f32(None, None, None, 0, 4, 4, 2, None) == 0    # empty batch is a no-op
    # descriptor-taking entry points reject a null descriptor
    for name in ("otal_conv_wgrad", "otal_conv1a_fwd", "otal_conv1a_wgrad", "otal_maxpool_fwd", "otal_msl_forward"):
        assert getattr(lib, name)(None, None) == -1, name
        assert "null descriptor" in _lib.last_error(), name
    # uint8 ingest: odd crop width, crop larger than the frame, null source
    for args in ((4096, None, None, 4096, None, 1, 4, 8, 8, 6, 5, None), (4096, None, None, 4096, None, 1, 4, 8, 8, 12, 6, None),
                 (None, None, None, 4096, None, 1, 4, 8, 8, 6, 6, None)):
        assert lib.otal_clip_ingest_u8(*args) == -1 and "clip_ingest_u8" in _lib.last_error()


def test_product_path_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from opental_b200 import ops
    with pytest.raises(RuntimeError):
        ops.bmp_forward(torch.zeros(1, 2, 4), torch.zeros(1, 1, 4))
    with pytest.raises(RuntimeError):
        ops.split_bf16(torch.zeros(8))
