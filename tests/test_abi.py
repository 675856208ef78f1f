"""CPU-only: the C-ABI library is built, loads, and exports exactly the symbols include/diffmm_b200.h
declares (no compute call is made without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "diffmm_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dmm_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from diffmm_b200 import _lib
    return _lib


def test_header_and_binding_agree(lib):
    assert _declared() == sorted(lib.PROTOTYPES)


def test_library_exports_every_declared_symbol(lib):
    handle = lib.load()
    for name in _declared():
        assert hasattr(handle, name), name
    assert handle.dmm_version() >= 100


def test_epilogue_struct_layout_matches_c(lib):
    # struct dmm_gemm_epilogue: ptr, i32, f32, f32, (pad), ptr, i64, ptr, i64, ptr, ptr, i64, ptr, ptr, i64, i32, (pad)
    # ..., ptr post_bias, ptr cmax, i64 ld_cmax
    assert ctypes.sizeof(lib.GemmEpilogue) == 136
    assert lib.GemmEpilogue.residual.offset == 24 and lib.GemmEpilogue.ld_out16.offset == 72
    assert lib.GemmEpilogue.cmax.offset == 120 and lib.GemmEpilogue.ld_cmax.offset == 128


def test_no_fallback_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from diffmm_b200 import ops
    with pytest.raises(lib.DiffMMError):
        ops.spmm(None, torch.zeros(4, 64)) if False else ops._ctx(torch.zeros(1))
