# save_ivfadc_index / load_ivfadc_index in the reference's on-disk format
# (reference src/persistency.jl:1-160, naive coarse quantizer; SURVEY.md Appendix B):
# 9 text header lines, then raw little-endian centroids, codebooks (codes, then the vectors
# matrix row by row), the rotation matrix, and per list: Int64 length, ids, codes.
# All lists cross the device boundary in ONE bulk call per GPU (ivfadc_export_all / ivfadc_import_all); the
# per-list records of the format are written / parsed on the host.
# Deviation (documented in INTEGRATION.md): an index built with coarse_quantizer = :hnsw is saved with a
# "NaiveQuantizer" header -- the engine's coarse search is exact either way and it holds no HNSW graph -- so the
# file loads everywhere (reference included) as a naive-quantizer index.  Files the REFERENCE saved with an
# HNSWQuantizer carry a serialised graph this loader does not parse: they are rejected, not misread.

function save_ivfadc_index(filename::AbstractString, ivfadc::IVFADCIndex{U,I,Dc,Dr,T}) where {U,I,Dc,Dr,T}
    open(filename, "w") do fid
        save_ivfadc_index(fid, ivfadc)
    end
end

function save_ivfadc_index(io::IO, ivfadc::IVFADCIndex{U,I,Dc,Dr,T}) where {U,I,Dc,Dr,T}
    rq = ivfadc.residual_quantizer
    nrows, kc = size(ivfadc.centroids)
    m = length(rq.codebooks)
    d, k = size(rq.codebooks[1].vectors)
    println(io, "$nrows $kc")
    println(io, "$(length(ivfadc)) $m $k $d")
    println(io, "NaiveQuantizer")       # :hnsw indexes are saved with the exact (naive) quantizer
    println(io, typeof(rq.quantization))
    println(io, U); println(io, I); println(io, Dc); println(io, Dr); println(io, T)
    write(io, ivfadc.centroids)
    for cb in rq.codebooks
        write(io, cb.codes)
        write(io, permutedims(cb.vectors))   # row j of the d x k matrix after row j-1
    end
    write(io, Matrix{T}(rq.rot))
    # one shard per GPU (a single handle is one shard): every cell lives on exactly one of them
    handles = ivfadc.group != C_NULL ? [capi_group_handle(ivfadc.group, i - 1) for i in 1:capi_group_size(ivfadc.group)] :
                                       [ivfadc.handle]
    sizes = [capi_list_sizes(h, kc) for h in handles]
    packed = [capi_export_all(handles[s], sizes[s], m) for s in eachindex(handles)]
    cursor = zeros(Int, length(handles))
    for c in 1:kc
        s = something(findfirst(x -> x[c] > 0, sizes), 1)
        len = Int(sizes[s][c]); o = cursor[s]
        write(io, Int64(len)); write(io, I.(view(packed[s][1], o + 1:o + len))); write(io, packed[s][2][:, o + 1:o + len])
        cursor[s] = o + len
    end
end

function load_ivfadc_index(filename::AbstractString)
    open(filename, "r") do fid
        load_ivfadc_index(fid)
    end
end

_parse_type(line) = getfield(occursin("Distances", line) || isdefined(Distances, Symbol(split(line, ".")[end])) ?
                             Distances : (isdefined(QuantizedArrays, Symbol(split(line, ".")[end])) ? QuantizedArrays : Base),
                             Symbol(split(line, ".")[end]))

function load_ivfadc_index(io::IO)
    nrows, kc = parse.(Int, split(readline(io)))
    n, m, k, d = parse.(Int, split(readline(io)))
    cq = strip(readline(io))
    split(cq, ".")[end] == "NaiveQuantizer" ||
        error("coarse quantizer $cq: files with a serialised HNSW graph (reference src/persistency.jl:214-251) are not supported")
    Qz = _parse_type(readline(io)); U = _parse_type(readline(io)); I = _parse_type(readline(io))
    Dc = _parse_type(readline(io)); Dr = _parse_type(readline(io)); T = _parse_type(readline(io))
    centroids = Matrix{T}(undef, nrows, kc); read!(io, centroids)
    cbs = Vector{QuantizedArrays.CodeBook{U,T}}(undef, m)
    for i in 1:m
        codes = Vector{U}(undef, k); read!(io, codes)
        vt = Matrix{T}(undef, k, d); read!(io, vt)
        cbs[i] = QuantizedArrays.CodeBook(codes, permutedims(vt))
    end
    rot = Matrix{T}(undef, nrows, nrows); read!(io, rot)
    rq = QuantizedArrays.ArrayQuantizer(Qz(), (nrows, n), cbs, k, Dr(), rot)
    h, g = _upload(centroids, rq, I, Dc(), Dr(), :naive)
    sizes = Vector{Int64}(undef, kc); ids = Vector{UInt64}(undef, n); codes = Matrix{UInt8}(undef, m, n)
    o = 0
    for c in 1:kc
        len = Int(read(io, Int64)); sizes[c] = len
        li = Vector{I}(undef, len); read!(io, li); ids[o + 1:o + len] .= li
        lc = Matrix{U}(undef, m, len); read!(io, lc); codes[:, o + 1:o + len] .= lc
        o += len
    end
    if g == C_NULL
        capi_import_all(h, sizes, ids, codes)
        capi_set_length(h, n)
    else
        world = capi_group_size(g)
        owners = _balanced_owners(Int.(sizes), world)
        capi_group_set_cell_owners(g, owners)
        starts = cumsum(vcat(0, sizes[1:end - 1]))
        for s in 1:world
            mine = findall(==(s - 1), owners)
            cols = reduce(vcat, [collect(starts[c] + 1:starts[c] + sizes[c]) for c in mine]; init=Int[])
            ssz = zeros(Int64, kc); ssz[mine] .= sizes[mine]
            hs = capi_group_handle(g, s - 1)
            capi_import_all(hs, ssz, ids[cols], codes[:, cols])
            capi_set_length(hs, n)
        end
    end
    _wrap(centroids, rq, I, Dc(), :naive, (h, g))
end
