# gen_from_reference.jl — golden vectors from the REAL reference (TEST INFRASTRUCTURE).
#
# Run by __graft_entry__.build() when a Julia toolchain and the reference are found (`which julia`, and
# /root/reference or baseline/_ref); it has never run in the image this repository was built in, which has no Julia:
# until it does, parity with the reference is pinned only by the oracle's restatement (DESIGN.md §7).
#
#     julia --project=<reference> oracle/gen_from_reference.jl <reference-dir> tests/golden/reference_vectors.bin
#
# Output (little-endian; read by tests/test_oracle.py::test_reference_vectors_pin_the_oracle when present):
#     int64  n_cases
#     per case:  int64 n | float64 logw[n] | float64 w[n] (exp_norm) | float64 lse (logsumexp) | float64 ess (ess_perc)
#                | float64 r[n] (the uniforms stratified_resample drew) | int64 idx[n] (1-based ancestors)
#                | float64 us[n] (arbitrary sorted uniforms) | int64 idx2[n] (icdf(w, us))
using Random

refdir, outpath = ARGS[1], ARGS[2]
push!(LOAD_PATH, refdir)
using WeightedSampling
const WS = WeightedSampling

# stratified_resample(weights) draws rand() once per slot from the global RNG (src/resampling.jl:39-41): seed, record
# the scalar draws, seed again, call it — the recorded r are exactly the uniforms it consumed.
function stratified_with_uniforms(w::Vector{Float64}, seed::Int)
    Random.seed!(seed)
    r = [rand() for _ in 1:length(w)]
    Random.seed!(seed)
    idx = WS.stratified_resample(w)
    return r, idx
end

cases = [(1, 0.0), (2, 1.0), (7, 0.5), (255, 2.0), (2048, 0.5), (2049, 4.0), (100_003, 2.0)]
open(outpath, "w") do io
    write(io, Int64(length(cases)))
    for (ci, (n, s)) in enumerate(cases)
        Random.seed!(1000 + ci)
        logw = s .* randn(n) .- 700.0
        w = WS.exp_norm(logw)
        lse = WS.logsumexp(logw)
        ess = WS.ess_perc(w)
        r, idx = stratified_with_uniforms(w, 2000 + ci)
        us = sort(rand(n))
        idx2 = WS.icdf(w, min.(us, prevfloat(sum(w))))      # keep the last uniform below the CDF's end (icdf would throw)
        write(io, Int64(n)); write(io, logw); write(io, w); write(io, lse); write(io, ess)
        write(io, r); write(io, Int64.(idx)); write(io, min.(us, prevfloat(sum(w)))); write(io, Int64.(idx2))
    end
end
println("reference vectors written to ", outpath)
