# External validation hook (SURVEY.md section 8c): pin the engine's results to REAL IVFADC.jl.
#
# The build image of this repo has no Julia, so the CPU oracle (oracle/) could only be checked against the
# reference by reading; this script closes the loop wherever Julia + JuliaNeighbors/IVFADC.jl v0.1.4 are installed:
#
#   julia julia/validate_against_reference.jl [tests/golden/validation]
#
# The directory holds a bundle written by tests/golden/make_validation_bundle.py on a GPU box:
#   index.ivfadc   an index saved by THIS engine in the reference's on-disk format (src/persistency.jl:1-80)
#   queries.bin    "nq D k w\n", then Float32 Q[nq][D], Int32 counts[nq], UInt64 ids[nq][k] (0-based),
#                  Float32 dists[nq][k] -- the engine's knn_search results (exact-table mode, IVFADC_FLAG_LUT_EXACT)
# The script loads the index with the reference's own load_ivfadc_index, runs the reference's knn_search
# (src/index.jl:204-273) on the same queries and reports: id mismatches, max relative distance error, and whether
# the distances are bit-identical (they should be: the engine's exact mode follows the reference's summation order).
using IVFADC   # the REFERENCE package, not julia/IVFADC of this repo

dir = length(ARGS) >= 1 ? ARGS[1] : joinpath(@__DIR__, "..", "tests", "golden", "validation")
index = load_ivfadc_index(joinpath(dir, "index.ivfadc"))
println(index)

open(joinpath(dir, "queries.bin"), "r") do io
    nq, D, k, w = parse.(Int, split(readline(io)))
    Q = Matrix{Float32}(undef, D, nq); read!(io, Q)
    counts = Vector{Int32}(undef, nq); read!(io, counts)
    ids = Matrix{UInt64}(undef, k, nq); read!(io, ids)
    dists = Matrix{Float32}(undef, k, nq); read!(io, dists)

    idmis = 0; cntmis = 0; bitmis = 0; maxrel = 0.0; total = 0
    for j in 1:nq
        ri, rd = knn_search(index, Q[:, j], k, w=w)
        n = Int(counts[j])
        if length(ri) != n
            cntmis += 1
            continue
        end
        for t in 1:n
            total += 1
            idmis += (UInt64(ri[t]) != ids[t, j])
            bitmis += (reinterpret(UInt32, rd[t]) != reinterpret(UInt32, dists[t, j]))
            maxrel = max(maxrel, abs(Float64(rd[t]) - Float64(dists[t, j])) / max(abs(Float64(rd[t])), 1e-30))
        end
    end
    println("queries $nq, k $k, w $w: $total results compared")
    println("  result-count mismatches : $cntmis")
    println("  id mismatches           : $idmis   (ties at equal distance may legitimately differ: see DESIGN.md section 2)")
    println("  distances not bit-equal : $bitmis")
    println("  max relative distance error: $maxrel   (north_star tolerance 1e-5)")
    ok = cntmis == 0 && maxrel <= 1e-5
    println(ok ? "VALIDATION OK" : "VALIDATION FAILED")
    exit(ok ? 0 : 1)
end
