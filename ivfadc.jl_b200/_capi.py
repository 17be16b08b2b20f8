"""ctypes binding of libivfadc_cuda.so -- one Python function per symbol of include/ivfadc.h.

This is the Python twin of the `ccall` stubs in julia/IVFADC/src/capi.jl.  The
library is the product; there is no fallback: if it cannot be loaded, or no CUDA device is
present, every call raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, byref, c_char_p, c_double, c_int32, c_int64, c_uint8, c_uint64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libivfadc_cuda.so")

OK = 0
ERR_BAD_ARG, ERR_CAPACITY, ERR_CUDA, ERR_OOM, ERR_UNSUPPORTED, ERR_EMPTY = -1, -2, -3, -4, -5, -6
F32, F64 = 0, 1
SQEUCLIDEAN, EUCLIDEAN, CITYBLOCK, COSINEDIST = 0, 1, 2, 3
METRICS = {"SqEuclidean": 0, "Euclidean": 1, "Cityblock": 2, "CosineDist": 3}
LAST, FIRST = 0, 1
FLAG_SCAN_LEGACY, FLAG_SCAN_QLANE, FLAG_LUT_EXACT, FLAG_LUT_MMASYNC, FLAG_SCAN_TMEM_V1, FLAG_COARSE_SCALAR = 1, 2, 4, 8, 16, 32
FLAG_TEST_MERGE_SWEEP = 64
FLAG_COARSE_FFMA, FLAG_TEST_COARSE_REDO = 128, 256

# every symbol include/ivfadc.h declares (tests check that the library exports all of them)
SYMBOLS = [
    "ivfadc_abi_version", "ivfadc_device_count", "ivfadc_create", "ivfadc_destroy", "ivfadc_last_error",
    "ivfadc_add", "ivfadc_encode", "ivfadc_coarse_search", "ivfadc_search", "ivfadc_search_device",
    "ivfadc_search_local_device", "ivfadc_coarse_search_device", "ivfadc_search_probes_local_device",
    "ivfadc_merge_device", "ivfadc_delete", "ivfadc_pop", "ivfadc_length",
    "ivfadc_list_sizes", "ivfadc_export_list", "ivfadc_import_list", "ivfadc_export_quantizers",
    "ivfadc_set_length", "ivfadc_get_stats", "ivfadc_reset_stats", "ivfadc_debug_tables", "ivfadc_set_stats_timing",
    "ivfadc_nccl_unique_id", "ivfadc_comm_init_rank", "ivfadc_comm_destroy", "ivfadc_set_graph_replay",
    "ivfadc_search_sharded", "ivfadc_search_sharded_device", "ivfadc_sharded_step_bytes", "ivfadc_set_cell_owners",
    "ivfadc_check_async", "ivfadc_group_create", "ivfadc_group_destroy", "ivfadc_group_size", "ivfadc_group_handle",
    "ivfadc_group_last_error", "ivfadc_group_set_cell_owners", "ivfadc_group_add", "ivfadc_group_search",
    "ivfadc_group_delete", "ivfadc_group_pop", "ivfadc_group_length",
    "ivfadc_add_device", "ivfadc_export_all", "ivfadc_import_all", "ivfadc_synth_uniform_device",
    "ivfadc_synth_blobs_device", "ivfadc_reserve", "ivfadc_set_centroids_device", "ivfadc_kmeanspp_device",
    "ivfadc_kmeans_accumulate_device", "ivfadc_kmeans_finish_device",
]
NCCL_ID_BYTES = 128


class Config(ctypes.Structure):
    _fields_ = [(n, c_int32) for n in (
        "dim", "kc", "m", "ksub", "dtype", "id_bytes", "metric_coarse", "metric_resid", "device",
        "shard_rank", "shard_world", "flags")]


class Stats(ctypes.Structure):
    _fields_ = [
        ("searches", c_uint64), ("queries", c_uint64), ("scanned_vectors", c_uint64),
        ("scan_code_bytes", c_uint64), ("gpu_launches", c_uint64), ("coarse_ms", c_double),
        ("plan_ms", c_double), ("scan_ms", c_double), ("merge_ms", c_double), ("encode_ms", c_double),
        ("scan_launches", c_uint64), ("last_scan_kernel", c_uint64), ("comm_ms", c_double), ("last_coarse_redo", c_uint64),
        ("reserved", c_uint64 * 1),
    ]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_ if n != "reserved"}


class IvfadcError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libivfadc_cuda error {code}: {msg}")
        self.code = code


_lib = None


def load(build_if_missing: bool = True):
    """Load the shared library (building it in-tree first when it is absent)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            raise OSError(f"{LIB_PATH} is missing; run `python -c 'import __graft_entry__ as g; g.build()'`")
        from .build import build
        build()
    lib = ctypes.CDLL(LIB_PATH)
    H = c_void_p
    lib.ivfadc_abi_version.restype = c_int32
    lib.ivfadc_device_count.restype = c_int32
    lib.ivfadc_create.argtypes = [POINTER(H), POINTER(Config), c_void_p, c_void_p, c_void_p]
    lib.ivfadc_destroy.argtypes = [H]
    lib.ivfadc_last_error.argtypes = [H]
    lib.ivfadc_last_error.restype = c_char_p
    lib.ivfadc_add.argtypes = [H, c_void_p, c_int64, c_int32, c_void_p, c_int32, c_void_p]
    lib.ivfadc_encode.argtypes = [H, c_void_p, c_int64, c_void_p, c_int32, c_void_p, c_void_p]
    lib.ivfadc_coarse_search.argtypes = [H, c_void_p, c_int64, c_int32, c_void_p, c_void_p]
    lib.ivfadc_search.argtypes = [H, c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p, c_void_p]
    lib.ivfadc_search_device.argtypes = [H, c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p,
                                         c_void_p, c_void_p]
    lib.ivfadc_search_local_device.argtypes = [H, c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p,
                                               c_void_p, c_void_p, c_void_p]
    lib.ivfadc_coarse_search_device.argtypes = [H, c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_void_p]
    lib.ivfadc_search_probes_local_device.argtypes = [H, c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p,
                                                      c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.ivfadc_merge_device.argtypes = [H, c_int32, c_int64, c_int32, c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_void_p, c_void_p, c_void_p]
    lib.ivfadc_delete.argtypes = [H, c_void_p, c_int64]
    lib.ivfadc_pop.argtypes = [H, c_int32, c_void_p, POINTER(c_int32)]
    lib.ivfadc_length.argtypes = [H, POINTER(c_int64)]
    lib.ivfadc_list_sizes.argtypes = [H, c_void_p]
    lib.ivfadc_export_list.argtypes = [H, c_int32, c_void_p, c_void_p]
    lib.ivfadc_import_list.argtypes = [H, c_int32, c_void_p, c_void_p, c_int64]
    lib.ivfadc_add_device.argtypes = [H, c_void_p, c_int64, c_int32, c_void_p, c_int32, c_void_p]
    lib.ivfadc_reserve.argtypes = [H, c_int64, c_void_p]
    lib.ivfadc_set_centroids_device.argtypes = [H, c_void_p]
    lib.ivfadc_kmeanspp_device.argtypes = [c_void_p, c_int64, c_int32, c_int32, c_int32, c_uint64, c_int32, c_void_p, c_void_p,
                                           c_void_p, c_void_p]
    lib.ivfadc_kmeans_accumulate_device.argtypes = [c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.ivfadc_kmeans_finish_device.argtypes = [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p]
    lib.ivfadc_export_all.argtypes = [H, c_void_p, c_void_p]
    lib.ivfadc_import_all.argtypes = [H, c_void_p, c_void_p, c_void_p]
    lib.ivfadc_synth_uniform_device.argtypes = [c_void_p, c_int64, c_int64, c_int32, c_uint64, c_void_p]
    lib.ivfadc_synth_blobs_device.argtypes = [c_void_p, c_void_p, c_int64, c_int64, c_int32, c_int32, c_uint64,
                                              ctypes.c_float, c_void_p, c_void_p]
    lib.ivfadc_export_quantizers.argtypes = [H, c_void_p, c_void_p, c_void_p]
    lib.ivfadc_set_length.argtypes = [H, c_int64]
    lib.ivfadc_get_stats.argtypes = [H, POINTER(Stats)]
    lib.ivfadc_reset_stats.argtypes = [H]
    lib.ivfadc_debug_tables.argtypes = [H, c_void_p]
    lib.ivfadc_set_stats_timing.argtypes = [H, c_int32]
    lib.ivfadc_nccl_unique_id.argtypes = [c_void_p]
    lib.ivfadc_comm_init_rank.argtypes = [H, c_void_p, c_int32, c_int32]
    lib.ivfadc_comm_destroy.argtypes = [H]
    lib.ivfadc_set_graph_replay.argtypes = [H, c_int32]
    lib.ivfadc_search_sharded.argtypes = [H, c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p, c_void_p]
    lib.ivfadc_search_sharded_device.argtypes = [H, c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p,
                                                 c_void_p, c_void_p]
    lib.ivfadc_sharded_step_bytes.argtypes = [H, c_int64, c_int32, c_int32, POINTER(c_int64), POINTER(c_int64),
                                              POINTER(c_int64)]
    lib.ivfadc_set_cell_owners.argtypes = [H, c_void_p]
    lib.ivfadc_check_async.argtypes = [H, c_void_p]
    lib.ivfadc_group_create.argtypes = [POINTER(H), POINTER(Config), c_int32, c_void_p, c_void_p, c_void_p, c_void_p]
    lib.ivfadc_group_destroy.argtypes = [H]
    lib.ivfadc_group_size.argtypes = [H]
    lib.ivfadc_group_handle.argtypes = [H, c_int32, POINTER(H)]
    lib.ivfadc_group_last_error.argtypes = [H]
    lib.ivfadc_group_last_error.restype = c_char_p
    lib.ivfadc_group_set_cell_owners.argtypes = [H, c_void_p]
    lib.ivfadc_group_add.argtypes = [H, c_void_p, c_int64, c_int32, c_void_p, c_int32, c_void_p]
    lib.ivfadc_group_search.argtypes = [H, c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p, c_void_p]
    lib.ivfadc_group_delete.argtypes = [H, c_void_p, c_int64]
    lib.ivfadc_group_pop.argtypes = [H, c_int32, c_void_p]
    lib.ivfadc_group_length.argtypes = [H, POINTER(c_int64)]
    for name in SYMBOLS:
        fn = getattr(lib, name)
        if name not in ("ivfadc_last_error", "ivfadc_group_last_error"):
            fn.restype = c_int32
    _lib = lib
    return lib


def check(handle, rc):
    if rc != OK:
        msg = load().ivfadc_last_error(handle) if handle else b""
        raise IvfadcError(rc, (msg or b"").decode("utf-8", "replace"))


def ptr(a):
    """Host numpy array -> void*; None -> NULL."""
    return None if a is None else a.ctypes.data_as(c_void_p)
