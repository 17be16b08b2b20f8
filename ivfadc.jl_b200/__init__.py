"""ivfadc.jl_b200 -- B200-native drop-in for the hot path of JuliaNeighbors/IVFADC.jl.

csrc/            hand-written sm_100a CUDA kernels + the C ABI (include/ivfadc.h)
_capi.py         ctypes binding of libivfadc_cuda.so
index.py         Python mirror of the reference's public API (the Julia glue's twin)
persistency.py   the reference's on-disk format
sharded.py       cell-sharded multi-GPU search over torch.distributed (one process per GPU)
training.py      quantizer training utilities (outside the parity scope)
julia/IVFADC/    the Julia glue package (ccall), reviewed but not runnable in this image

The directory name contains a dot, so import it through the repo-root shim: `import ivfadc_jl_b200`.
"""
from . import _capi
from .index import (IVFADCIndex, delete_from_index, knn_search, pop, popfirst, push, push_batch,
                    pushfirst)
from .persistency import load_ivfadc_index, save_ivfadc_index

__all__ = ["IVFADCIndex", "knn_search", "push", "pushfirst", "push_batch", "pop", "popfirst",
           "delete_from_index", "save_ivfadc_index", "load_ivfadc_index"]
