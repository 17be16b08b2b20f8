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
    assert lib.ivfadc_abi_version() == 3


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


def test_delete_from_index_id_conversion_without_a_device():
    """delete_from_index! takes 1-based integers and converts them with I.(points .- 1) (src/utils.jl:93): 0 or
    negative -> InexactError, ids beyond the id type -> InexactError; duplicates and unknown ids pass through (the
    engine ignores them).  Host logic only: the C-ABI call is captured by a stub, the integer-array path and the
    generic path must hand over the same ids."""
    import ctypes
    import ivfadc_jl_b200 as iv

    class StubLib:
        def __init__(self):
            self.calls = []

        def ivfadc_delete(self, h, ptr, n):
            self.calls.append(np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_uint64)), shape=(n,)).copy())
            return 0

    for id_type, maxid in ((np.uint8, 255), (np.uint32, 2 ** 32 - 1)):
        idx = iv.IVFADCIndex.__new__(iv.IVFADCIndex)
        idx.I, idx._h, idx._lib = np.dtype(id_type), ctypes.c_void_p(1), StubLib()
        iv.delete_from_index(idx, np.array([5, 1, 5, 200], dtype=np.int64))       # vectorised path
        iv.delete_from_index(idx, [5, 1, 5, 200])                                  # generic path
        assert [c.tolist() for c in idx._lib.calls] == [[4, 0, 4, 199]] * 2
        for bad in (0, -3, maxid + 2):
            with pytest.raises(OverflowError):
                iv.delete_from_index(idx, np.array([1, bad], dtype=np.int64))
            with pytest.raises(OverflowError):
                iv.delete_from_index(idx, [1, bad])
        iv.delete_from_index(idx, np.array([], dtype=np.int64))
        assert len(idx._lib.calls) == 2
        idx._h = None   # nothing to destroy
