# ccall stubs: one Julia function per symbol of include/ivfadc.h.
# Library lookup: ENV["LIBIVFADC_CUDA"] or the dynamic-loader path.

const LIBIVFADC = Ref{String}("libivfadc_cuda")
function __init__()
    LIBIVFADC[] = get(ENV, "LIBIVFADC_CUDA", "libivfadc_cuda")
    ver = ccall((:ivfadc_abi_version, LIBIVFADC[]), Cint, ())
    ver == 3 || error("libivfadc_cuda ABI $ver, this package binds ABI 3")
    # no CPU fallback: fail at load time if there is no device
    ccall((:ivfadc_device_count, LIBIVFADC[]), Cint, ()) > 0 ||
        error("libivfadc_cuda found no CUDA device (the engine has no CPU fallback)")
end

const IVFADC_OK = Cint(0)
const IVFADC_ERR_CAPACITY = Cint(-2)
const IVFADC_ERR_EMPTY = Cint(-6)
const IVFADC_F32, IVFADC_F64 = Cint(0), Cint(1)
const IVFADC_SQEUCLIDEAN = Cint(0)
const IVFADC_LAST, IVFADC_FIRST = Cint(0), Cint(1)

# mirrors `struct ivfadc_config` (12 x int32)
struct CConfig
    dim::Cint; kc::Cint; m::Cint; ksub::Cint; dtype::Cint; id_bytes::Cint
    metric_coarse::Cint; metric_resid::Cint; device::Cint
    shard_rank::Cint; shard_world::Cint; flags::Cint
end

const Handle = Ptr{Cvoid}

_dtype(::Type{Float32}) = IVFADC_F32
_dtype(::Type{Float64}) = IVFADC_F64
_dtype(::Type{T}) where T = throw(ArgumentError("libivfadc_cuda supports Float32 / Float64, got $T"))

function _check(h::Handle, rc::Cint)
    rc == IVFADC_OK && return nothing
    msg = h == C_NULL ? "" : unsafe_string(ccall((:ivfadc_last_error, LIBIVFADC[]), Cstring, (Handle,), h))
    error("libivfadc_cuda error $rc: $msg")
end

function capi_create(cfg::CConfig, centroids::Matrix{T}, cbvectors::Array{T,3}, cbcodes::Matrix{UInt8}) where T
    # centroids D x kc, cbvectors dsub x ksub x m, cbcodes ksub x m -- column-major Julia arrays are
    # exactly the T[kc][D], T[m][ksub][dsub], uint8[m][ksub] layouts of the header.
    out = Ref{Handle}(C_NULL)
    rc = ccall((:ivfadc_create, LIBIVFADC[]), Cint,
               (Ref{Handle}, Ref{CConfig}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{UInt8}),
               out, Ref(cfg), centroids, cbvectors, cbcodes)
    rc == IVFADC_OK || error("ivfadc_create failed with status $rc (no CUDA device, or configuration outside the hot-path scope)")
    out[]
end

capi_destroy(h::Handle) = ccall((:ivfadc_destroy, LIBIVFADC[]), Cint, (Handle,), h)

function capi_add(h::Handle, X::AbstractMatrix{T}, position::Cint;
                  assign::Union{Nothing,Vector{Int64}}=nothing) where T
    n = size(X, 2)
    rc = ccall((:ivfadc_add, LIBIVFADC[]), Cint,
               (Handle, Ptr{Cvoid}, Int64, Cint, Ptr{Int64}, Cint, Ptr{Cint}),
               h, X, n, position, assign === nothing ? C_NULL : pointer(assign), 1, C_NULL)
    rc
end

function capi_search(h::Handle, Q::Matrix{T}, k::Int, w::Int) where T
    nq = size(Q, 2)
    ids = Matrix{UInt64}(undef, k, nq); dists = Matrix{T}(undef, k, nq); counts = Vector{Cint}(undef, nq)
    rc = ccall((:ivfadc_search, LIBIVFADC[]), Cint,
               (Handle, Ptr{Cvoid}, Int64, Cint, Cint, Ptr{UInt64}, Ptr{Cvoid}, Ptr{Cint}),
               h, Q, nq, k, w, ids, dists, counts)
    _check(h, rc)
    ids, dists, counts
end

function capi_delete(h::Handle, ids0::Vector{UInt64})
    _check(h, ccall((:ivfadc_delete, LIBIVFADC[]), Cint, (Handle, Ptr{UInt64}, Int64), h, ids0, length(ids0)))
end

function capi_pop(h::Handle, position::Cint, ::Type{T}, D::Int) where T
    v = Vector{T}(undef, D)
    rc = ccall((:ivfadc_pop, LIBIVFADC[]), Cint, (Handle, Cint, Ptr{Cvoid}, Ptr{Cint}), h, position, v, C_NULL)
    rc, v
end

function capi_length(h::Handle)
    n = Ref{Int64}(0)
    ccall((:ivfadc_length, LIBIVFADC[]), Cint, (Handle, Ref{Int64}), h, n)
    Int(n[])
end

function capi_list_sizes(h::Handle, kc::Int)
    s = Vector{Int64}(undef, kc)
    _check(h, ccall((:ivfadc_list_sizes, LIBIVFADC[]), Cint, (Handle, Ptr{Int64}), h, s))
    s
end

function capi_export_list(h::Handle, cell0::Int, len::Int, m::Int)
    ids = Vector{UInt64}(undef, len); codes = Matrix{UInt8}(undef, m, len)
    _check(h, ccall((:ivfadc_export_list, LIBIVFADC[]), Cint, (Handle, Cint, Ptr{UInt64}, Ptr{UInt8}),
                    h, cell0, ids, codes))
    ids, codes
end

function capi_import_list(h::Handle, cell0::Int, ids::Vector{UInt64}, codes::Matrix{UInt8})
    _check(h, ccall((:ivfadc_import_list, LIBIVFADC[]), Cint, (Handle, Cint, Ptr{UInt64}, Ptr{UInt8}, Int64),
                    h, cell0, ids, codes, length(ids)))
end

capi_set_length(h::Handle, n::Int) =
    _check(h, ccall((:ivfadc_set_length, LIBIVFADC[]), Cint, (Handle, Int64), h, n))

# ---- bulk persistency (ivfadc_export_all / ivfadc_import_all): all lists in one device <-> host transfer ----------
function capi_export_all(h::Handle, sizes::Vector{Int64}, m::Int)
    n = Int(sum(sizes))
    ids = Vector{UInt64}(undef, n); codes = Matrix{UInt8}(undef, m, n)
    _check(h, ccall((:ivfadc_export_all, LIBIVFADC[]), Cint, (Handle, Ptr{UInt64}, Ptr{UInt8}), h, ids, codes))
    ids, codes
end

function capi_import_all(h::Handle, sizes::Vector{Int64}, ids::Vector{UInt64}, codes::Matrix{UInt8})
    _check(h, ccall((:ivfadc_import_all, LIBIVFADC[]), Cint, (Handle, Ptr{Int64}, Ptr{UInt64}, Ptr{UInt8}),
                    h, sizes, ids, codes))
end

# ---- several GPUs in one process (ivfadc_group_*): ENV["IVFADC_DEVICES"] = "0,1,2,3" --------------------------
# The lists are sharded by cell over the devices; one knn_search batch call fans out by itself (coarse slice ->
# all-gather of the probe lists -> local scans -> all-gather of the candidates -> merge, NCCL over NVLink inside
# the library) and returns what the single-GPU engine returns, bit for bit.
const Group = Ptr{Cvoid}

function _devices()
    s = strip(get(ENV, "IVFADC_DEVICES", ""))
    isempty(s) ? Cint[] : Cint.(parse.(Int, split(s, ",")))
end

function _gcheck(g::Group, rc::Cint)
    rc == IVFADC_OK && return nothing
    msg = g == C_NULL ? "" : unsafe_string(ccall((:ivfadc_group_last_error, LIBIVFADC[]), Cstring, (Group,), g))
    error("libivfadc_cuda error $rc: $msg")
end

function capi_group_create(cfg::CConfig, devices::Vector{Cint}, centroids::Matrix{T}, cbvectors::Array{T,3},
                           cbcodes::Matrix{UInt8}) where T
    out = Ref{Group}(C_NULL)
    rc = ccall((:ivfadc_group_create, LIBIVFADC[]), Cint,
               (Ref{Group}, Ref{CConfig}, Cint, Ptr{Cint}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{UInt8}),
               out, Ref(cfg), length(devices), devices, centroids, cbvectors, cbcodes)
    rc == IVFADC_OK || error("ivfadc_group_create failed with status $rc (devices $devices)")
    out[]
end

capi_group_destroy(g::Group) = ccall((:ivfadc_group_destroy, LIBIVFADC[]), Cint, (Group,), g)
capi_group_size(g::Group) = Int(ccall((:ivfadc_group_size, LIBIVFADC[]), Cint, (Group,), g))

function capi_group_handle(g::Group, i0::Int)
    out = Ref{Handle}(C_NULL)
    _gcheck(g, ccall((:ivfadc_group_handle, LIBIVFADC[]), Cint, (Group, Cint, Ref{Handle}), g, i0, out))
    out[]
end

capi_group_set_cell_owners(g::Group, owners::Vector{Cint}) =
    _gcheck(g, ccall((:ivfadc_group_set_cell_owners, LIBIVFADC[]), Cint, (Group, Ptr{Cint}), g, owners))

function capi_group_add(g::Group, X::AbstractMatrix{T}, position::Cint;
                        assign::Union{Nothing,Vector{Int64}}=nothing) where T
    ccall((:ivfadc_group_add, LIBIVFADC[]), Cint,
          (Group, Ptr{Cvoid}, Int64, Cint, Ptr{Int64}, Cint, Ptr{Cint}),
          g, X, size(X, 2), position, assign === nothing ? C_NULL : pointer(assign), 1, C_NULL)
end

function capi_group_search(g::Group, Q::Matrix{T}, k::Int, w::Int) where T
    nq = size(Q, 2)
    ids = Matrix{UInt64}(undef, k, nq); dists = Matrix{T}(undef, k, nq); counts = Vector{Cint}(undef, nq)
    _gcheck(g, ccall((:ivfadc_group_search, LIBIVFADC[]), Cint,
                     (Group, Ptr{Cvoid}, Int64, Cint, Cint, Ptr{UInt64}, Ptr{Cvoid}, Ptr{Cint}),
                     g, Q, nq, k, w, ids, dists, counts))
    ids, dists, counts
end

capi_group_delete(g::Group, ids0::Vector{UInt64}) =
    _gcheck(g, ccall((:ivfadc_group_delete, LIBIVFADC[]), Cint, (Group, Ptr{UInt64}, Int64), g, ids0, length(ids0)))

function capi_group_pop(g::Group, position::Cint, ::Type{T}, D::Int) where T
    v = Vector{T}(undef, D)
    rc = ccall((:ivfadc_group_pop, LIBIVFADC[]), Cint, (Group, Cint, Ptr{Cvoid}), g, position, v)
    rc, v
end

function capi_group_length(g::Group)
    n = Ref{Int64}(0)
    ccall((:ivfadc_group_length, LIBIVFADC[]), Cint, (Group, Ref{Int64}), g, n)
    Int(n[])
end

# capacity hint before a bulk build (ivfadc_reserve): exact per-list sizes, e.g. counts(kmeans result)
capi_reserve(h::Handle, n::Int, sizes::Vector{Int64}) =
    _check(h, ccall((:ivfadc_reserve, LIBIVFADC[]), Cint, (Handle, Int64, Ptr{Int64}), h, n, sizes))
