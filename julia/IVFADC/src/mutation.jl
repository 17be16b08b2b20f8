# push!/pushfirst!, pop!/popfirst!, delete_from_index!  (reference src/utils.jl:29-161).
# Argument checks are raised here as AssertionError BEFORE the ccall, in the reference's order
# (dimension first, capacity second: src/utils.jl:133-135), so `@test_throws AssertionError`
# tests of the reference keep passing.

_add(ivfadc::IVFADCIndex, X, position) = ivfadc.group != C_NULL ? capi_group_add(ivfadc.group, X, position) :
                                                                   capi_add(ivfadc.handle, X, position)
_rc(ivfadc::IVFADCIndex, rc) = ivfadc.group != C_NULL ? _gcheck(ivfadc.group, rc) : _check(ivfadc.handle, rc)

function _push!(ivfadc::IVFADCIndex{U,I,Dc,Dr,T}, point::Vector{T}, position::Cint) where {U,I,Dc,Dr,T}
    nrows, nvectors = size(ivfadc)
    @assert nrows == length(point) "Adding to index requires same dimensionality"
    @assert _ID_BITS[I] >= log2(nvectors + 1) "Cannot index, exceeding index capacity"
    rc = _add(ivfadc, reshape(point, :, 1), position)
    rc == IVFADC_ERR_CAPACITY && throw(AssertionError("Cannot index, exceeding index capacity"))
    _rc(ivfadc, rc)
    return nothing
end

push!(ivfadc::IVFADCIndex{U,I,Dc,Dr,T}, point::Vector{T}) where {U,I,Dc,Dr,T} = _push!(ivfadc, point, IVFADC_LAST)
pushfirst!(ivfadc::IVFADCIndex{U,I,Dc,Dr,T}, point::Vector{T}) where {U,I,Dc,Dr,T} = _push!(ivfadc, point, IVFADC_FIRST)

# Batched extension (not in the reference): n x push! in one call; column j gets id N + j - 1.
function push!(ivfadc::IVFADCIndex{U,I,Dc,Dr,T}, points::Matrix{T}) where {U,I,Dc,Dr,T}
    nrows, nvectors = size(ivfadc)
    @assert nrows == size(points, 1) "Adding to index requires same dimensionality"
    @assert _ID_BITS[I] >= log2(nvectors + size(points, 2)) "Cannot index, exceeding index capacity"
    _rc(ivfadc, _add(ivfadc, points, IVFADC_LAST))
    return nothing
end

function _pop!(ivfadc::IVFADCIndex{U,I,Dc,Dr,T}, position::Cint) where {U,I,Dc,Dr,T}
    nrows, nvectors = size(ivfadc)
    @assert nvectors > 0 "Cannot pop element from empty index"
    rc, v = ivfadc.group != C_NULL ? capi_group_pop(ivfadc.group, position, T, nrows) :
                                     capi_pop(ivfadc.handle, position, T, nrows)
    _rc(ivfadc, rc)
    return v   # centroid + decoded residual (reference src/utils.jl:58-59)
end

pop!(ivfadc::IVFADCIndex) = _pop!(ivfadc, IVFADC_LAST)
popfirst!(ivfadc::IVFADCIndex) = _pop!(ivfadc, IVFADC_FIRST)

# 1-based ids in, like the reference (src/utils.jl:90-93); I.(points .- 1) keeps its InexactError
# for ids <= 0.  Duplicates / unknown ids are ignored by the library; survivors are renumbered
# new = old - #(deleted ids < old) by one on-device compaction pass.
function delete_from_index!(ivfadc::IVFADCIndex{U,I,Dc,Dr,T}, points::Vector{<:Integer}) where {U,I,Dc,Dr,T}
    shifted = I.(points .- 1)
    ivfadc.group != C_NULL ? capi_group_delete(ivfadc.group, UInt64.(shifted)) :
                             capi_delete(ivfadc.handle, UInt64.(shifted))
    return nothing
end
