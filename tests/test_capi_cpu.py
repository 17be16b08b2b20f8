"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol that
include/ivfadc.h declares, and refuses to work without a CUDA device (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "ivfadc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ivfadc_[a-z_]+)\s*\(", src)))


def test_header_and_binding_agree():
    from ivfadc_jl_b200 import _capi
    assert declared_symbols() == sorted(_capi.SYMBOLS)


def test_library_exports_every_declared_symbol():
    from ivfadc_jl_b200 import _capi
    lib = _capi.load()
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert lib.ivfadc_abi_version() == 1


def test_struct_layouts_match_header():
    from ivfadc_jl_b200 import _capi
    assert ctypes.sizeof(_capi.Config) == 12 * 4
    assert ctypes.sizeof(_capi.Stats) == 5 * 8 + 5 * 8 + 8 + 4 * 8


def test_no_cpu_fallback():
    """Without a device every entry point must fail loudly; with one this test is vacuous."""
    from ivfadc_jl_b200 import _capi
    import ivfadc_jl_b200 as iv
    lib = _capi.load()
    if lib.ivfadc_device_count() > 0:
        pytest.skip("a CUDA device is present")
    cent = np.zeros((4, 8), dtype=np.float32)
    cb = np.zeros((2, 4, 4), dtype=np.float32)
    with pytest.raises(_capi.IvfadcError) as ei:
        iv.IVFADCIndex.from_quantizers(cent, cb)
    assert ei.value.code == _capi.ERR_CUDA
    # null handle -> bad argument, never a crash
    assert lib.ivfadc_search(None, None, 0, 1, 1, None, None, None) == _capi.ERR_BAD_ARG
    assert lib.ivfadc_destroy(None) == _capi.OK


def test_product_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under the package may reference it."""
    pkg = os.path.join(ROOT, "ivfadc.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        if os.sep + "build" in dirpath:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".jl")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "liboracle" not in txt and "from oracle" not in txt and "import oracle" not in txt, f


def test_constructor_asserts_fire_before_any_device_use():
    """test/index.jl:37-40 -- AssertionError for kc=1, k>N, m>D, UInt8 ids for 300 vectors."""
    import ivfadc_jl_b200 as iv
    data = np.random.default_rng(0).random((2, 300))
    with pytest.raises(AssertionError):
        iv.IVFADCIndex(data, kc=1, k=2, m=1)
    with pytest.raises(AssertionError):
        iv.IVFADCIndex(data, kc=2, k=301, m=1)
    with pytest.raises(AssertionError):
        iv.IVFADCIndex(data, kc=2, k=300, m=3)
    with pytest.raises(AssertionError):
        iv.IVFADCIndex(data, index_type=np.uint8)
    with pytest.raises(AssertionError):
        iv.IVFADCIndex(data, kc=2, k=2, m=1, coarse_quantizer="kdtree")
