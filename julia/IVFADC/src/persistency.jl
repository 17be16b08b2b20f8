# save_ivfadc_index / load_ivfadc_index in the reference's on-disk format
# (reference src/persistency.jl:1-160, naive coarse quantizer; SURVEY.md Appendix B):
# 9 text header lines, then raw little-endian centroids, codebooks (codes, then the vectors
# matrix row by row), the rotation matrix, and per list: Int64 length, ids, codes.
# Lists are exported from / imported into the device CSR one whole list per call.

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
    sizes = capi_list_sizes(ivfadc.handle, kc)
    for c in 1:kc
        ids, codes = capi_export_list(ivfadc.handle, c - 1, Int(sizes[c]), m)
        write(io, Int64(sizes[c])); write(io, I.(ids)); write(io, codes)
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
    readline(io)                                   # quantizer kind: exact GPU search either way
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
    h = _upload(centroids, rq, I, Dc(), Dr(), :naive)
    for c in 1:kc
        len = read(io, Int64)
        ids = Vector{I}(undef, len); read!(io, ids)
        codes = Matrix{U}(undef, m, len); read!(io, codes)
        capi_import_list(h, c - 1, UInt64.(ids), UInt8.(codes))
    end
    capi_set_length(h, n)
    _wrap(centroids, rq, I, Dc(), :naive, h)
end
