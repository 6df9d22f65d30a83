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


def test_struct_mirrors_match_the_library():
    """Every descriptor struct of the header has the size of its ctypes mirror (checked again at load time by _lib.load)."""
    from opental_b200 import _lib
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "opental_b200.h")).read()
    declared = sorted(set(re.findall(r"typedef struct (otal_\w+_desc)", header)))
    assert declared == sorted(_lib.STRUCT_MIRRORS), (declared, sorted(_lib.STRUCT_MIRRORS))
    for name, mirror in _lib.STRUCT_MIRRORS.items():
        assert lib.otal_abi_sizeof(name.encode()) == ctypes.sizeof(mirror) > 0, name
    assert lib.otal_abi_sizeof(b"no_such_struct") == 0 and lib.otal_abi_sizeof(None) == 0


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


def test_staged_u8_entry_points_validate_arguments_without_gpu():
    """otal_conv1a_fwd_u8 / otal_conv1a_wgrad_u8 / otal_clip_ingest_u8_raw / otal_border_class_sums: every rejection below
    happens before the first CUDA call."""
    from opental_b200 import _lib
    lib = _lib.load()
    for name in ("otal_conv1a_fwd_u8", "otal_conv1a_wgrad_u8"):
        assert getattr(lib, name)(None, None) == -1 and "null descriptor" in _lib.last_error(), name
    P = 4096        # never dereferenced
    ok = dict(N=1, T=8, H=8, W=16, Cout=64, tT=4, tH=4, tW=8, nsplit=3, relu=1, out_cstride=64, out_coff=0, x_hi=P, x_lo=None,
              w_hi=P, w_lo=P, scale=P, shift=P, y_hi=P, y_lo=P)
    for bad, msg in ((dict(nsplit=1), "conv1a_u8"), (dict(shift=None), "conv1a_u8"), (dict(T=7), "conv1a_u8"), (dict(H=4), "conv1a_u8"),
                     (dict(w_lo=None), "null plane"), (dict(tW=4), "128 positions")):
        d = _lib.Conv1aDesc(**{**ok, **bad})
        assert lib.otal_conv1a_fwd_u8(ctypes.byref(d), None) == -1 and msg in _lib.last_error(), (bad, _lib.last_error())
    okw = dict(N=1, T=8, H=16, W=16, Cout=64, tT=1, tH=8, tW=8, nsplit=3, d_cstride=64, d_coff=0, x_hi=P, x_lo=None, d_hi=P, d_lo=P, dw=P)
    for bad, msg in ((dict(nsplit=1), "conv1a_wgrad_u8"), (dict(d_lo=None), "null pointer"), (dict(tW=4), "64 positions")):
        d = _lib.Conv1aWgradDesc(**{**okw, **bad})
        assert lib.otal_conv1a_wgrad_u8(ctypes.byref(d), None) == -1 and msg in _lib.last_error(), (bad, _lib.last_error())
    # raw ingest: odd crop width / null output;   class sums: C not a power of two, extent < 3, misaligned slice
    assert lib.otal_clip_ingest_u8_raw(P, None, None, P, 1, 4, 8, 8, 6, 5, None) == -1 and "clip_ingest_u8_raw" in _lib.last_error()
    assert lib.otal_clip_ingest_u8_raw(P, None, None, None, 1, 4, 8, 8, 6, 6, None) == -1
    assert lib.otal_clip_ingest_u8_raw(P, None, None, P, 0, 4, 8, 8, 6, 6, None) == 0          # empty batch
    for args in ((P, P, P, 1, 4, 4, 4, 48, 48, 0, None), (P, P, P, 1, 2, 4, 4, 64, 64, 0, None), (P, P, P, 1, 4, 4, 4, 64, 64, 4, None),
                 (P, P, None, 1, 4, 4, 4, 64, 64, 0, None)):
        assert lib.otal_border_class_sums(*args) == -1 and "border_class_sums" in _lib.last_error(), args
    assert lib.otal_border_class_sums(P, None, P, 0, 4, 4, 4, 64, 64, 0, None) == 0


def test_product_path_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from opental_b200 import ops
    with pytest.raises(RuntimeError):
        ops.bmp_forward(torch.zeros(1, 2, 4), torch.zeros(1, 1, 4))
    with pytest.raises(RuntimeError):
        ops.split_bf16(torch.zeros(8))


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / reference arm may
    import or execute anything under oracle/.  Static check over the product package (docstrings may cite oracle files)."""
    import ast
    import glob
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    names = {os.path.splitext(os.path.basename(f))[0] for f in glob.glob(os.path.join(root, "oracle", "*.py"))}
    assert "opental_oracle" in names
    offenders = []
    for path in glob.glob(os.path.join(root, "opental_b200", "**", "*.py"), recursive=True):
        tree = ast.parse(open(path).read())
        for node in ast.walk(tree):
            mods = []
            if isinstance(node, ast.Import):
                mods = [a.name for a in node.names]
            elif isinstance(node, ast.ImportFrom):
                mods = [node.module or ""]
            for m in mods:
                if m.split(".")[0] in names or m.split(".")[0] == "oracle":
                    offenders.append((os.path.relpath(path, root), m))
    assert offenders == [], offenders
