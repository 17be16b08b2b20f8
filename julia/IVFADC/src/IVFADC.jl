# IVFADC -- drop-in replacement of JuliaNeighbors/IVFADC.jl whose hot path runs in
# libivfadc_cuda (hand-written sm_100a kernels behind the C ABI of include/ivfadc.h).
#
# Same exports, same method signatures, same AssertionErrors as the reference
# (reference src/IVFADC.jl:13-20).  Training (k-means for the coarse quantizer and the PQ
# codebooks) still runs in Clustering.jl / QuantizedArrays.jl exactly as in the reference; the
# trained quantizers are uploaded once and everything after that -- encoding, list storage,
# search, mutation -- is a `ccall`.
#
# NOT EXECUTABLE IN THE BUILD IMAGE (no Julia there): reviewed against the Python twin
# ivfadc.jl_b200/index.py, which drives the same C ABI and is what the test-suite runs.
module IVFADC

using Distances
using Clustering
using QuantizedArrays
using Libdl

import Base: push!, pushfirst!, pop!, popfirst!

export IVFADCIndex,
       delete_from_index!,
       knn_search,
       save_ivfadc_index,
       load_ivfadc_index

include("defaults.jl")
include("capi.jl")
include("index.jl")
include("mutation.jl")
include("persistency.jl")

end # module
