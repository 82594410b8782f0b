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

end # module
