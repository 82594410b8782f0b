# WSB200.jl — ccall shim that makes libwsb200.so a storage/kernel backend of WeightedSampling.jl.
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no Julia.  The Python host
# (weightedsampling.jl_b200/) issues the same C-ABI call sequence and is what the tests drive; this
# file is the binding a maintainer of the reference would add (see INTEGRATION.md).
module WSB200

using WeightedSampling
import WeightedSampling: AbstractParticleStore, SMCState, ParticleTransformer, nparticles, hascol, getcol, colnames,
    broadcast_setcol!, resample!, apply!, score!, Resample, Sequence, Loop, Cond

const LIB = get(ENV, "WSB200_LIB", joinpath(@__DIR__, "..", "weightedsampling.jl_b200", "lib", "libwsb200.so"))

# ---- include/wsb200.h mirrors -------------------------------------------------------------------------
struct WsTok
    op::Int32; col::Int32; comp::Int32; reserved::Int32; val::Float64
end
struct WsExpr
    toks::Ptr{WsTok}; n::Int32; reserved::Int32
end
struct WsResampleInfo
    fired::Int32; resampled::Int32; ess_perc::Float64; log_mean_w::Float64; n_clamped::Int64
end
WsResampleInfo() = WsResampleInfo(0, 0, NaN, NaN, 0)

const TOK_CONST, TOK_PLANE, TOK_ADD, TOK_SUB, TOK_MUL, TOK_DIV, TOK_NEG, TOK_EXP, TOK_LOG, TOK_SQRT, TOK_SQUARE,
      TOK_SIN, TOK_COS, TOK_ABS, TOK_POW, TOK_RANDN, TOK_RANDU, TOK_RANDEXP, TOK_LT, TOK_LE, TOK_EQ, TOK_SELECT,
      TOK_MIN, TOK_MAX, TOK_NOT, TOK_LGAMMA, TOK_LOG1P, TOK_EXPM1, TOK_TAN, TOK_ATAN, TOK_TANH, TOK_FLOOR = Int32.(0:31)
struct WsPlaneStats
    mean::Float64; median::Float64; std::Float64; min::Float64; max::Float64; hist::NTuple{8,Float64}
end

function check(ctx, rc)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:ws_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx))
    rc == -1 ? throw(ArgumentError(msg)) : error("wsb200 error $rc: $msg")
end

# ---- storage backend (src/stores.jl:28-35) ----------------------------------------------------------
mutable struct DeviceColumnStore <: AbstractParticleStore
    ctx::Ptr{Cvoid}
    n::Int
end

function DeviceColumnStore(n::Integer; device=0, seed=0, ess_perc_min=0.5, resampler=0)
    ref = Ref{Ptr{Cvoid}}(C_NULL)
    check(C_NULL, ccall((:ws_create, LIB), Cint, (Ref{Ptr{Cvoid}}, Int64, Cint, UInt64, Cdouble, Cint),
                        ref, n, device, seed, ess_perc_min, resampler))
    s = DeviceColumnStore(ref[], Int(n))
    finalizer(s -> ccall((:ws_destroy, LIB), Cint, (Ptr{Cvoid},), s.ctx), s)
    return s
end

nparticles(s::DeviceColumnStore) = s.n

function lookup(s::DeviceColumnStore, name::Symbol)
    id = Ref{Int32}(-1); w = Ref{Int32}(0)
    check(s.ctx, ccall((:ws_col_lookup, LIB), Cint, (Ptr{Cvoid}, Cstring, Ref{Int32}, Ref{Int32}), s.ctx, String(name), id, w))
    return id[], w[]
end
hascol(s::DeviceColumnStore, name::Symbol) = lookup(s, name)[1] >= 0

function colnames(s::DeviceColumnStore)
    cnt = Ref{Int32}(0)
    check(s.ctx, ccall((:ws_col_count, LIB), Cint, (Ptr{Cvoid}, Ref{Int32}), s.ctx, cnt))
    buf = Vector{UInt8}(undef, 256); w = Ref{Int32}(0)
    map(0:cnt[]-1) do i
        check(s.ctx, ccall((:ws_col_info, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{UInt8}, Int32, Ref{Int32}), s.ctx, i, buf, 256, w))
        Symbol(unsafe_string(pointer(buf)))
    end
end

# getcol returns a host COPY (device-resident access goes through statements)
function getcol(s::DeviceColumnStore, name::Symbol)
    id, w = lookup(s, name)
    id >= 0 || throw(KeyError(name))
    out = Matrix{Float64}(undef, s.n, w)                       # plane-major == column-major n x w
    check(s.ctx, ccall((:ws_col_download, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}), s.ctx, id, out))
    return w == 1 ? vec(out) : [out[i, :] for i in 1:s.n]
end

# resample!(store, indices): particle i <- old particle indices[i] (1-based in Julia, 0-based in the ABI)
function resample!(s::DeviceColumnStore, indices::AbstractVector{<:Integer})
    idx = Int32.(indices .- 1)
    check(s.ctx, ccall((:ws_gather, LIB), Cint, (Ptr{Cvoid}, Ptr{Int32}), s.ctx, idx))
end

# broadcast_setcol! with host data (identity of an uploaded vector); device statements use the ops below
function broadcast_setcol!(s::DeviceColumnStore, name::Symbol, ::typeof(identity), args::Tuple{AbstractVector{Float64}})
    id = Ref{Int32}(-1)
    check(s.ctx, ccall((:ws_col_ensure, LIB), Cint, (Ptr{Cvoid}, Cstring, Int32, Ref{Int32}), s.ctx, String(name), 1, id))
    check(s.ctx, ccall((:ws_col_upload, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Float64}), s.ctx, id[], args[1]))
end
broadcast_setcol!(::DeviceColumnStore, name::Symbol, f, args::Tuple) =
    error("column $name: host closures cannot run on the device; the @model front-end lowers statements to device ops")

ctx(state::SMCState{DeviceColumnStore}) = state.store.ctx

# ---- expressions: what `vectorize` (src/rewrites.jl:146-219) builds, as postfix tokens -----------------
struct DeviceExpr
    toks::Vector{WsTok}
end
DeviceExpr(c::Real) = DeviceExpr([WsTok(TOK_CONST, 0, 0, 0, Float64(c))])
plane(col::Integer, comp::Integer) = DeviceExpr([WsTok(TOK_PLANE, col, comp, 0, 0.0)])
binop(op, a::DeviceExpr, b::DeviceExpr) = DeviceExpr(vcat(a.toks, b.toks, WsTok(op, 0, 0, 0, 0.0)))
Base.:+(a::DeviceExpr, b::DeviceExpr) = binop(TOK_ADD, a, b)
Base.:-(a::DeviceExpr, b::DeviceExpr) = binop(TOK_SUB, a, b)
Base.:*(a::DeviceExpr, b::DeviceExpr) = binop(TOK_MUL, a, b)
Base.:/(a::DeviceExpr, b::DeviceExpr) = binop(TOK_DIV, a, b)
unop(op, a::DeviceExpr) = DeviceExpr(vcat(a.toks, WsTok(op, 0, 0, 0, 0.0)))
Base.exp(a::DeviceExpr) = unop(TOK_EXP, a); Base.log(a::DeviceExpr) = unop(TOK_LOG, a)
Base.sqrt(a::DeviceExpr) = unop(TOK_SQRT, a); Base.cos(a::DeviceExpr) = unop(TOK_COS, a); Base.sin(a::DeviceExpr) = unop(TOK_SIN, a)
Base.:<(a::DeviceExpr, b::DeviceExpr) = binop(TOK_LT, a, b)          # Bool columns are 1.0 / 0.0 planes
Base.:<=(a::DeviceExpr, b::DeviceExpr) = binop(TOK_LE, a, b)
Base.:|(a::DeviceExpr, b::DeviceExpr) = binop(TOK_MAX, a, b)          # a || b
Base.:&(a::DeviceExpr, b::DeviceExpr) = binop(TOK_MIN, a, b)          # a && b
Base.:!(a::DeviceExpr) = unop(TOK_NOT, a)
# ifelse.(c, a, b): what `vectorize` makes of `c ? a : b` on particle variables (src/rewrites.jl:193-199)
Base.ifelse(c::DeviceExpr, a::DeviceExpr, b::DeviceExpr) = DeviceExpr(vcat(c.toks, a.toks, b.toks, WsTok(TOK_SELECT, 0, 0, 0, 0.0)))
randn_tok() = DeviceExpr([WsTok(TOK_RANDN, 0, 0, 0, 0.0)])            # fresh variates inside a sampler expression
randu_tok() = DeviceExpr([WsTok(TOK_RANDU, 0, 0, 0, 0.0)])
randexp_tok() = DeviceExpr([WsTok(TOK_RANDEXP, 0, 0, 0, 0.0)])
cexpr(e::DeviceExpr) = WsExpr(pointer(e.toks), length(e.toks), 0)

# ---- device statements: each apply! is one ccall -----------------------------------------------------------
struct DeviceAssign <: ParticleTransformer; col::Int32; comp::Int32; rhs::DeviceExpr; end
struct DeviceSampleNormal <: ParticleTransformer; col::Int32; comp::Int32; mu::DeviceExpr; sigma::DeviceExpr; end
struct DeviceObserveNormal <: ParticleTransformer; obs::DeviceExpr; mu::DeviceExpr; sigma::DeviceExpr; end

function apply!(t::DeviceAssign, state::SMCState{DeviceColumnStore})          # src/transformers.jl:28-32
    GC.@preserve t check(ctx(state), ccall((:ws_assign, LIB), Cint, (Ptr{Cvoid}, Int32, Int32, Ref{WsExpr}),
                                           ctx(state), t.col, t.comp, cexpr(t.rhs)))
    state.depth += 1
end
function apply!(t::DeviceSampleNormal, state::SMCState{DeviceColumnStore})    # src/transformers.jl:172-182
    GC.@preserve t check(ctx(state), ccall((:ws_sample_normal, LIB), Cint, (Ptr{Cvoid}, Int32, Int32, Ref{WsExpr}, Ref{WsExpr}),
                                           ctx(state), t.col, t.comp, cexpr(t.mu), cexpr(t.sigma)))
    state.depth += 1
end
function apply!(t::DeviceObserveNormal, state::SMCState{DeviceColumnStore})   # src/transformers.jl:228-235
    GC.@preserve t check(ctx(state), ccall((:ws_observe_normal, LIB), Cint, (Ptr{Cvoid}, Ref{WsExpr}, Ref{WsExpr}, Ref{WsExpr}),
                                           ctx(state), cexpr(t.obs), cexpr(t.mu), cexpr(t.sigma)))
    state.weights_changed = true
    state.depth += 1
end
function apply!(::Resample, state::SMCState{DeviceColumnStore})                # src/transformers.jl:474-498
    info = Ref(WsResampleInfo())
    check(ctx(state), ccall((:ws_resample, LIB), Cint, (Ptr{Cvoid}, Ref{WsResampleInfo}), ctx(state), info))
    if info[].fired != 0
        state.resampled = info[].resampled != 0
        state.weights_changed = false
    end
    return nothing
end
# A WeightedKernel(sampler, weighter, logpdf) (src/types.jl:226-230) whose three parts are device expressions:
# `x ~ K(args...)` is one ccall; the library samples, weights and records logpdf on its score tape.
struct DeviceSampleExpr <: ParticleTransformer
    col::Int32; comp::Int32; sampler::DeviceExpr; weighter::Union{DeviceExpr,Nothing}; logpdf::Union{DeviceExpr,Nothing}
end
function apply!(t::DeviceSampleExpr, state::SMCState{DeviceColumnStore})      # src/transformers.jl:172-182
    GC.@preserve t begin
        w = t.weighter === nothing ? C_NULL : Ref(cexpr(t.weighter))
        l = t.logpdf === nothing ? C_NULL : Ref(cexpr(t.logpdf))
        check(ctx(state), ccall((:ws_sample_expr, LIB), Cint, (Ptr{Cvoid}, Int32, Int32, Ref{WsExpr}, Ptr{WsExpr}, Ptr{WsExpr}),
                                ctx(state), t.col, t.comp, cexpr(t.sampler), w, l))
    end
    t.weighter === nothing || (state.weights_changed = true)
    state.depth += 1
end
# e.g. default_kernels.Uniform (src/default_kernels.jl:101) as device expressions:
#   sampler  (a, b)    -> a + (b - a) * randu_tok()
#   logpdf   (a, b, x) -> ifelse((a <= x) & (x <= b), -log(b - a), DeviceExpr(-Inf))

# score! of the device statements is a no-op on the host: the library records the tape itself and
# ws_move folds it (device form of the score! walk).
score!(::Union{DeviceAssign,DeviceSampleNormal,DeviceObserveNormal}, state, c) = (c.depth += 1; nothing)

# ---- analysis (src/utils.jl) ---------------------------------------------------------------------------------
function log_evidence(state::SMCState{DeviceColumnStore})
    le = Ref(0.0); ess = Ref(0.0)
    check(ctx(state), ccall((:ws_log_evidence, LIB), Cint, (Ptr{Cvoid}, Ref{Float64}, Ref{Float64}), ctx(state), le, ess))
    le[]
end
function exp_norm(state::SMCState{DeviceColumnStore})
    out = Vector{Float64}(undef, state.store.n)
    check(ctx(state), ccall((:ws_exp_norm, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}), ctx(state), out))
    out
end

# describe(state) (src/utils.jl:183-289): every statistic, including the StatsBase weighted median, is computed
# on the device; 13 numbers per plane come back
function describe_plane(state::SMCState{DeviceColumnStore}, name::Symbol, comp::Integer=0)
    id, _ = lookup(state.store, name)
    out = Ref(WsPlaneStats(0, 0, 0, 0, 0, ntuple(_ -> 0.0, 8))); ess = Ref(0.0)
    check(ctx(state), ccall((:ws_describe, LIB), Cint, (Ptr{Cvoid}, Int32, Ref{Int32}, Ref{Int32}, Ref{WsPlaneStats}, Ref{Float64}),
                            ctx(state), 1, Ref(Int32(id)), Ref(Int32(comp)), out, ess))
    out[], ess[]
end

# trajectory storage by genealogy: columns that are not read keep their order and the ancestor vectors are kept
# instead (include/wsb200.h: ws_set_genealogy); on by default, nothing to do for a model that keeps x{t}
set_genealogy!(state::SMCState{DeviceColumnStore}, on::Bool; budget_bytes::Integer=0) =
    check(ctx(state), ccall((:ws_set_genealogy, LIB), Cint, (Ptr{Cvoid}, Cint, Int64), ctx(state), on, budget_bytes))

# ---- the other default kernels of the configs: MvNormal, Exponential, Weight --------------------------------
struct DeviceSampleMvNormal <: ParticleTransformer; col::Int32; mu::Vector{DeviceExpr}; cov::Matrix{Float64}; end
struct DeviceObserveMvNormal <: ParticleTransformer; obs::Vector{DeviceExpr}; mu::Vector{DeviceExpr}; cov::Matrix{Float64}; end
struct DeviceSampleExponential <: ParticleTransformer; col::Int32; comp::Int32; theta::DeviceExpr; end
struct DeviceWeight <: ParticleTransformer; term::DeviceExpr; end
rowmajor(m::Matrix{Float64}) = collect(vec(permutedims(m)))          # the ABI takes the covariance row-major

function apply!(t::DeviceSampleMvNormal, state::SMCState{DeviceColumnStore})   # src/transformers.jl:172-182, default_kernels.jl:93
    GC.@preserve t begin
        mus = [cexpr(e) for e in t.mu]
        check(ctx(state), ccall((:ws_sample_mvnormal, LIB), Cint, (Ptr{Cvoid}, Int32, Int32, Ptr{WsExpr}, Ptr{Float64}),
                                ctx(state), t.col, length(t.mu), mus, rowmajor(t.cov)))
    end
    state.depth += 1
end
function apply!(t::DeviceObserveMvNormal, state::SMCState{DeviceColumnStore})  # src/transformers.jl:228-235
    GC.@preserve t begin
        obs = [cexpr(e) for e in t.obs]; mus = [cexpr(e) for e in t.mu]
        check(ctx(state), ccall((:ws_observe_mvnormal, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{WsExpr}, Ptr{WsExpr}, Ptr{Float64}),
                                ctx(state), length(t.mu), obs, mus, rowmajor(t.cov)))
    end
    state.weights_changed = true
    state.depth += 1
end
function apply!(t::DeviceSampleExponential, state::SMCState{DeviceColumnStore}) # default_kernels.jl:87
    GC.@preserve t check(ctx(state), ccall((:ws_sample_exponential, LIB), Cint, (Ptr{Cvoid}, Int32, Int32, Ref{WsExpr}),
                                           ctx(state), t.col, t.comp, cexpr(t.theta)))
    state.depth += 1
end
function apply!(t::DeviceWeight, state::SMCState{DeviceColumnStore})           # src/transformers.jl:283-289
    GC.@preserve t check(ctx(state), ccall((:ws_weight_expr, LIB), Cint, (Ptr{Cvoid}, Ref{WsExpr}), ctx(state), cexpr(t.term)))
    state.weights_changed = true
    state.depth += 1
end
score!(::Union{DeviceSampleMvNormal,DeviceObserveMvNormal,DeviceSampleExponential,DeviceWeight,DeviceSampleExpr}, state, c) =
    (c.depth += 1; nothing)

# ---- Move (src/transformers.jl:588-623) with the RW / autoRW proposals (src/move_kernels.jl:189-253) -----------
struct WsMoveSpec
    n_targets::Int32; col::Ptr{Int32}; comp::Ptr{Int32}; proposal::Int32; has_bounds::Int32
    lo::Ptr{Float64}; hi::Ptr{Float64}; step::Float64; diversity::Float64; target_depth::Int64
end
struct WsMoveInfo
    ran::Int32; reserved::Int32; diversity::Float64; n_accepted::Int64
end
struct DeviceMove <: ParticleTransformer
    cols::Vector{Int32}; comps::Vector{Int32}
    proposal::Int32                       # 0 RW, 1 autoRW
    step::Float64                         # RW: step_size; autoRW: min_step (1e-3)
    lo::Vector{Float64}; hi::Vector{Float64}   # empty: bounds === nothing
    diversity::Float64                    # NaN: always move
end
function apply!(t::DeviceMove, state::SMCState{DeviceColumnStore})
    info = Ref(WsMoveInfo(0, 0, NaN, 0))
    GC.@preserve t begin
        spec = WsMoveSpec(length(t.cols), pointer(t.cols), pointer(t.comps), t.proposal, isempty(t.lo) ? 0 : 1,
                          isempty(t.lo) ? C_NULL : pointer(t.lo), isempty(t.hi) ? C_NULL : pointer(t.hi),
                          t.step, t.diversity, -1)                  # -1: score up to state.depth, as Move.apply! does
        check(ctx(state), ccall((:ws_move, LIB), Cint, (Ptr{Cvoid}, Ref{WsMoveSpec}, Ref{WsMoveInfo}), ctx(state), spec, info))
    end
    return nothing                                                  # depth-neutral, weights untouched (transformers.jl:585-586)
end
score!(::DeviceMove, state, c) = nothing

function marginal_diversity(state::SMCState{DeviceColumnStore}, cols::Vector{Int32}, comps::Vector{Int32})
    out = Ref(0.0)
    check(ctx(state), ccall((:ws_marginal_diversity, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Int32}, Ptr{Int32}, Ref{Float64}),
                            ctx(state), length(cols), cols, comps, out))
    out[]
end
function score_logpdf(state::SMCState{DeviceColumnStore}, target_depth::Integer)      # src/types.jl:183-206
    out = Vector{Float64}(undef, state.store.n)
    check(ctx(state), ccall((:ws_score_logpdf, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}), ctx(state), target_depth, out))
    out
end

# ---- @E / expectation and sample(state, n) (src/utils.jl:11,45-68,102-118) ------------------------------------
function expectation(fs::Vector{DeviceExpr}, state::SMCState{DeviceColumnStore})
    out = Vector{Float64}(undef, length(fs))
    GC.@preserve fs begin
        es = [cexpr(f) for f in fs]
        check(ctx(state), ccall((:ws_expectation, LIB), Cint, (Ptr{Cvoid}, Ptr{WsExpr}, Int32, Ptr{Float64}), ctx(state), es, length(fs), out))
    end
    out
end
function sample_rows(state::SMCState{DeviceColumnStore}, name::Symbol, n::Integer; replace::Bool=true)
    idx = Vector{Int64}(undef, n)
    check(ctx(state), ccall((:ws_sample_indices, LIB), Cint, (Ptr{Cvoid}, Int64, Cint, Ptr{Int64}), ctx(state), n, replace, idx))
    id, w = lookup(state.store, name)
    rows = Matrix{Float64}(undef, n, w)
    check(ctx(state), ccall((:ws_col_download_rows, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Int64}, Int64, Ptr{Float64}),
                            ctx(state), id, idx, n, rows))
    idx .+ 1, rows
end

# ---- replayed standard variates (parity tests: SURVEY 8c consumption order) -------------------------------------
set_replay_normals(state, v::Vector{Float64}) = check(ctx(state), ccall((:ws_set_replay_normals, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64), ctx(state), v, length(v)))
set_replay_uniforms(state, v::Vector{Float64}) = check(ctx(state), ccall((:ws_set_replay_uniforms, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64), ctx(state), v, length(v)))
set_replay_exponentials(state, v::Vector{Float64}) = check(ctx(state), ccall((:ws_set_replay_exponentials, LIB), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int64), ctx(state), v, length(v)))

# ---- one filter over several GPUs: one Julia process per GPU (INTEGRATION.md) -----------------------------------
nccl_unique_id() = (buf = zeros(UInt8, 128); check(C_NULL, ccall((:ws_nccl_unique_id, LIB), Cint, (Ptr{UInt8},), buf)); buf)
function ShardedColumnStore(n_global::Integer, rank::Integer, nranks::Integer, id::Vector{UInt8}; device=rank, seed=0,
                            ess_perc_min=0.5, resampler=0)
    ref = Ref{Ptr{Cvoid}}(C_NULL)
    check(C_NULL, ccall((:ws_create_sharded, LIB), Cint, (Ref{Ptr{Cvoid}}, Int64, Cint, Cint, Ptr{UInt8}, Cint, UInt64, Cdouble, Cint),
                        ref, n_global, rank, nranks, id, device, seed, ess_perc_min, resampler))
    nl = Ref{Int64}(0); ng = Ref{Int64}(0)
    check(ref[], ccall((:ws_n_particles, LIB), Cint, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}), ref[], nl, ng))
    s = DeviceColumnStore(ref[], Int(nl[]))          # the rank's shard: global slots [rank N / R, (rank + 1) N / R)
    finalizer(s -> ccall((:ws_destroy, LIB), Cint, (Ptr{Cvoid},), s.ctx), s)
    return s
end

end # module
