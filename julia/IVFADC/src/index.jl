# IVFADCIndex on top of libivfadc_cuda.  The type parameters are the reference's
# (src/index.jl:39-48); the three fields of the reference struct become: the trained quantizers
# (kept on the Julia side for persistency / show) and an opaque device handle that owns the
# inverted lists (device-resident CSR of uint8 PQ codes + ids).

mutable struct IVFADCIndex{U<:Unsigned, I<:Unsigned, Dc<:Distances.PreMetric, Dr<:Distances.PreMetric,
                           T<:AbstractFloat}
    centroids::Matrix{T}                                         # D x kc   (cq.vectors)
    residual_quantizer::QuantizedArrays.OrthogonalQuantizer{U,Dr,T,2}
    coarse_kind::Symbol                                          # :naive | :hnsw (-> exact GPU search)
    handle::Handle                                               # one GPU (C_NULL when `group` is used)
    group::Group                                                 # several GPUs, lists sharded by cell (or C_NULL)
end

const _ID_BITS = Dict(UInt8 => 8, UInt16 => 16, UInt32 => 32, UInt64 => 64)

# SqEuclidean runs on the tensor-core paths; Euclidean / Cityblock / CosineDist on the exact generic kernels
_metric_code(::Distances.SqEuclidean) = IVFADC_SQEUCLIDEAN
_metric_code(::Distances.Euclidean) = Cint(1)
_metric_code(::Distances.Cityblock) = Cint(2)
_metric_code(::Distances.CosineDist) = Cint(3)
_metric_code(d) = throw(ArgumentError("libivfadc_cuda supports SqEuclidean, Euclidean, Cityblock and CosineDist, got $(typeof(d))"))

# Returns (handle, group): ENV["IVFADC_DEVICES"] with more than one device id creates a group
# (ivfadc_group_create: one handle per device, NCCL communicator inside the library).
function _upload(centroids::Matrix{T}, rq, ::Type{I}, dc, dr, kind::Symbol; device::Int=0) where {T,I}
    m = length(rq.codebooks)
    dsub, ksub = size(rq.codebooks[1].vectors)
    cbv = Array{T,3}(undef, dsub, ksub, m)
    cbc = Matrix{UInt8}(undef, ksub, m)
    for i in 1:m
        cbv[:, :, i] .= rq.codebooks[i].vectors
        cbc[:, i] .= UInt8.(rq.codebooks[i].codes)
    end
    devs = _devices()
    cfg = CConfig(size(centroids, 1), size(centroids, 2), m, ksub, _dtype(T), sizeof(I),
                  _metric_code(dc), _metric_code(dr), length(devs) == 1 ? devs[1] : device, 0, 1, 0)
    length(devs) > 1 && return (C_NULL, capi_group_create(cfg, devs, centroids, cbv, cbc))
    (capi_create(cfg, centroids, cbv, cbc), C_NULL)
end

# cell -> device by greedy bin-packing on the list lengths (longest list onto the lightest device): every GPU
# scans about the same number of code bytes per batch
function _balanced_owners(sizes::Vector{Int}, world::Int)
    owners = Vector{Cint}(undef, length(sizes)); load = zeros(Int, world); cnt = zeros(Int, world)
    for c in sortperm(sizes, rev=true, alg=MergeSort)
        r = argmin(collect(zip(load, cnt)))
        owners[c] = r - 1; load[r] += sizes[c]; cnt[r] += 1
    end
    owners
end

function _wrap(centroids::Matrix{T}, rq::QuantizedArrays.OrthogonalQuantizer{U,Dr,T,2}, ::Type{I},
               dc::Dc, kind::Symbol, hg::Tuple{Handle,Group}) where {U,I,Dc,Dr,T}
    idx = IVFADCIndex{U,I,Dc,Dr,T}(centroids, rq, kind, hg[1], hg[2])
    finalizer(idx) do x
        x.handle != C_NULL && (capi_destroy(x.handle); x.handle = C_NULL)
        x.group != C_NULL && (capi_group_destroy(x.group); x.group = C_NULL)
    end
    idx
end

# Constructor: same keyword arguments, same checks and messages as reference src/index.jl:103-125.
function IVFADCIndex(data::Matrix{T};
                     kc::Int=DEFAULT_COARSE_K, k::Int=DEFAULT_QUANTIZATION_K, m::Int=DEFAULT_QUANTIZATION_M,
                     coarse_quantizer::Symbol=DEFAULT_COARSE_QUANTIZER,
                     coarse_distance::Distances.PreMetric=DEFAULT_COARSE_DISTANCE,
                     quantization_distance::Distances.PreMetric=DEFAULT_QUANTIZATION_DISTANCE,
                     quantization_method::Symbol=DEFAULT_QUANTIZATION_METHOD,
                     coarse_maxiter::Int=DEFAULT_COARSE_MAXITER,
                     quantization_maxiter::Int=DEFAULT_QUANTIZATION_MAXITER,
                     index_type::Type{I}=UInt32) where {I<:Unsigned, T<:AbstractFloat}
    nrows, nvectors = size(data)
    bits_required = ceil(Int, log2(nvectors))
    @assert kc >= 2 "Number of coarse clusters has to be >= 2"
    @assert k <= nvectors "Number of quantization levels  has to be <= $nvectors"
    @assert m >= 1 && m <= nrows "Number of codebooks has to be between 1 and $nrows"
    @assert coarse_quantizer in [:naive, :hnsw] "Coarse quantizer can be :naive or :hnsw only"
    @assert coarse_maxiter > 0 "Number of clustering iterations has to be > 0"
    @assert quantization_maxiter > 0 "Number of clustering iterations has to be > 0"
    @assert _ID_BITS[index_type] >= bits_required "$nvectors vectors require at least $bits_required index bits"

    # training stays where the reference has it (Clustering.jl, QuantizedArrays.jl)
    cmodel = kmeans(data, kc, maxiter=coarse_maxiter, distance=coarse_distance, init=:kmpp, display=:none)
    residuals = data .- cmodel.centers[:, cmodel.assignments]
    rq = build_quantizer(residuals, k=k, m=m, method=quantization_method,
                         distance=quantization_distance, maxiter=quantization_maxiter)

    h, g = _upload(cmodel.centers, rq, I, coarse_distance, quantization_distance, coarse_quantizer)
    # index build = ONE call: residuals w.r.t. the k-means assignments, PQ encoding, CSR fill,
    # ids ascending per list (reference _build_residuals + _build_inverted_index, src/index.jl:168-194)
    if g != C_NULL
        capi_group_set_cell_owners(g, _balanced_owners(counts(cmodel), capi_group_size(g)))
        _gcheck(g, capi_group_add(g, data, IVFADC_LAST, assign=Int64.(cmodel.assignments)))
    else
        capi_reserve(h, nvectors, Int64.(counts(cmodel)))   # the cluster sizes are known: one allocation
        _check(h, capi_add(h, data, IVFADC_LAST, assign=Int64.(cmodel.assignments)))
    end
    _wrap(cmodel.centers, rq, I, coarse_distance, coarse_quantizer, (h, g))
end

_search(ivfadc::IVFADCIndex, Q, k, w) = ivfadc.group != C_NULL ? capi_group_search(ivfadc.group, Q, k, w) :
                                                                  capi_search(ivfadc.handle, Q, k, w)

Base.length(ivfadc::IVFADCIndex) = ivfadc.group != C_NULL ? capi_group_length(ivfadc.group) : capi_length(ivfadc.handle)
Base.size(ivfadc::IVFADCIndex) = (size(ivfadc.centroids, 1), length(ivfadc))
Base.size(ivfadc::IVFADCIndex, i::Int) = size(ivfadc)[i]

Base.show(io::IO, ivfadc::IVFADCIndex{U,I,Dc,Dr,T}) where {U,I,Dc,Dr,T} = begin
    nvars, nvectors = size(ivfadc)
    m = length(ivfadc.residual_quantizer.codebooks)
    cqstr = ivfadc.coarse_kind == :hnsw ? "HNSW" : "naive"
    print(io, "IVFADCIndex, $cqstr coarse quantizer, $(m * sizeof(U) + sizeof(I))-byte encoding " *
              "($(sizeof(I)) + $(sizeof(U))×$m), $nvectors $T vectors")
end

# knn_search: single query (reference src/index.jl:204-258) and batch (src/index.jl:261-273).
# The batch method is ONE library call on the packed D x nq matrix.
function knn_search(ivfadc::IVFADCIndex{U,I,Dc,Dr,T}, point::Vector{T}, k::Int; w::Int=1) where {U,I,Dc,Dr,T}
    @assert k >= 1 "Number of neighbors must be k >= 1"
    @assert w >= 1 "Number of clusters to search in must be w >= 1"
    w = min(w, size(ivfadc.centroids, 2))
    ids, dists, counts = _search(ivfadc, reshape(point, :, 1), k, w)
    n = counts[1]
    return I.(ids[1:n, 1]), dists[1:n, 1]
end

function knn_search(ivfadc::IVFADCIndex{U,I,Dc,Dr,T}, points::Vector{Vector{T}}, k::Int; w::Int=1) where {U,I,Dc,Dr,T}
    @assert k >= 1 "Number of neighbors must be k >= 1"
    @assert w >= 1 "Number of clusters to search in must be w >= 1"
    w = min(w, size(ivfadc.centroids, 2))
    Q = reduce(hcat, points)
    ids, dists, counts = _search(ivfadc, Q, k, w)
    idxs = [I.(ids[1:counts[j], j]) for j in eachindex(points)]
    ds = [dists[1:counts[j], j] for j in eachindex(points)]
    return idxs, ds
end
