# EasyHybridCUDA.jl -- reference-side binding of libeasyhybrid_cuda.so (include/easyhybrid_cuda.h).
#
# STATUS: written against EasyHybrid.jl v0.2.0 / Lux 1.x sources; Julia is not installed in this repository's build
# image, so this file has NOT been executed here.  What IS executed: the identical C call sequence from C
# (tests/abi_replay.c, tests/test_abi_replay.py) and from Python/ctypes (easyhybrid.jl_b200/session.py, all GPU tests).
# tests/test_julia_dump.py + julia/parity_dump.jl close the numerical loop for anyone with Julia.
#
# Two seams, neither of which overwrites a method of the reference:
#
#  (1) per step -- a NEW method of Lux's own entry point on a NEW backend type:
#          Lux.Training.single_train_step!(::FusedCUDA, obj_fn, data, ts)
#      `run_epoch!` (src/training/epoch.jl:13-33) calls exactly this with `cfg.autodiff_backend`
#      (src/config/TrainingConfig.jl:52), so
#          train(model, data, (); autodiff_backend = FusedCUDA(), gdev = cpu_device())
#      runs the stock loop unchanged: DataLoader, collect_dim_data, isemptybatch, early stopping, checkpoints.  Every
#      batch crosses PCIe inside the call (eh_step_host), parameters come back after every step.
#
#  (2) per epoch -- `run_epoch_fused!`, same signature and return value as `run_epoch!`; the maintainer's patch is the
#      one line at the top of run_epoch! shown in INTEGRATION.md.  The split is staged on the device once, an epoch is
#      ONE call (eh_epoch) with the DataLoader's own permutation, so batch composition is the reference's.
#
# Not supported (eh_create answers EH_EUNSUPPORTED and the caller keeps AutoZygote()): per-branch optimisers,
# arbitrary `extra_loss` closures (the documented one, λ * weight_l2(ps.<branch>), is native: FusedCUDA(weight_l2 = ...)),
# Chain-valued `hidden_layers`, LSTM models.
module EasyHybridCUDA

using EasyHybrid
using EasyHybrid: SingleNNHybridModel, MultiNNHybridModel, TrainConfig, PerTarget
using Lux, ADTypes, ComponentArrays, MLUtils, Random, Optimisers

export FusedCUDA, run_epoch_fused!, trace_process_model

const LIB = get(ENV, "EASYHYBRID_CUDA_LIB", "libeasyhybrid_cuda.so")

# ---------------------------------------------------------------------------------------------------------------
# C structs and enums (include/easyhybrid_cuda.h)
# ---------------------------------------------------------------------------------------------------------------
struct EhPmArg; kind::Int32; index::Int32; end
struct EhPmInstr; op::Int32; a::Int32; b::Int32; imm::Float32; end
struct EhChainDesc
    n_in::Int32; in_cols::Ptr{Int32}; n_hidden::Int32; hidden::Ptr{Int32}
    n_out::Int32; activation::Int32; input_batchnorm::Int32
end
struct EhModelDesc
    abi_version::Int32
    n_pred::Int32; n_forc::Int32; n_targ::Int32
    n_chains::Int32; chains::Ptr{EhChainDesc}
    n_params::Int32; role::Ptr{Int32}; role_index::Ptr{Int32}
    deflt::Ptr{Float32}; lower::Ptr{Float32}; upper::Ptr{Float32}
    scale_nn_outputs::Int32
    process_model::Int32; n_pm_args::Int32; pm_args::Ptr{EhPmArg}
    pm_consts::NTuple{4, Float32}
    pm_prog::Ptr{EhPmInstr}; pm_len::Int32; pm_outputs::Ptr{Int32}
    loss_per_target::Ptr{Int32}; agg::Int32
    opt_kind::Int32; eta::Float32; beta1::Float32; beta2::Float32; eps::Float32; lambda::Float32
    adamw_decay_coupled_eta::Int32
    device::Int32; flags::Int32
    l2_lambda::Float32; l2_normalize::Int32; l2_chain_mask::UInt32      # ABI version 2: native weight_l2 extra loss
end

const EH_ABI_VERSION = Int32(2)
const ACT = Dict(:identity => 0, :tanh => 1, :tanh_fast => 1, :sigmoid => 2, :sigmoid_fast => 2, :σ => 2, :relu => 3, :swish => 4)
const LOSS = Dict(:mse => 0, :rmse => 1, :mae => 2, :nseLoss => 3, :pearsonLoss => 4, :kgeLoss => 5, :pbkgeLoss => 6)
const PM_RBQ10, PM_EXPO, PM_LINEAR, PM_LINEAR2, PM_EXPO2, PM_PROGRAM = 0, 1, 2, 3, 4, 100
const OP = Dict(:const => 0, :forcing => 1, :param => 2, :add => 10, :sub => 11, :mul => 12, :div => 13, :pow => 14, :min => 15,
    :max => 16, :neg => 20, :exp => 21, :log => 22, :sqrt => 23, :tanh => 24, :sigmoid => 25, :abs => 26, :sin => 27, :cos => 28)

last_error(ctx) = unsafe_string(ccall((:eh_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx))
check(ctx, st) = st == 0 || error("libeasyhybrid_cuda: status $st: " * last_error(ctx))

# ---------------------------------------------------------------------------------------------------------------
# Tracing the process model (the plug-in boundary, `m.mechanistic_model(; all_kwargs...)`,
# src/models/GenericHybridModel.jl:425): call it once with a symbolic scalar type, record a straight-line program.
# ---------------------------------------------------------------------------------------------------------------
struct Tape
    prog::Vector{Tuple{Symbol, Int32, Int32, Float32}}
end
const TAPE = Ref{Tape}()

struct Sym <: Real
    vid::Int32
end
emit(op::Symbol, a = 0, b = 0, imm = 0.0f0) = (push!(TAPE[].prog, (op, Int32(a), Int32(b), Float32(imm))); Sym(Int32(length(TAPE[].prog) - 1)))
lift(x::Sym) = x
lift(x::Real) = emit(:const, 0, 0, Float32(x))
Base.promote_rule(::Type{Sym}, ::Type{<:Real}) = Sym
Base.convert(::Type{Sym}, x::Real) = lift(x)
Base.convert(::Type{Sym}, x::Sym) = x
for (f, op) in ((:+, :add), (:-, :sub), (:*, :mul), (:/, :div), (:^, :pow), (:min, :min), (:max, :max))
    @eval Base.$f(a::Sym, b::Sym) = emit($(QuoteNode(op)), a.vid, b.vid)
    @eval Base.$f(a::Sym, b::Real) = $f(a, lift(b))
    @eval Base.$f(a::Real, b::Sym) = $f(lift(a), b)
end
Base.:^(a::Sym, b::Integer) = a^lift(Float32(b))      # (also stops Base.literal_pow from expanding x^2 into x * x)
Base.:-(a::Sym) = emit(:neg, a.vid)
Base.:+(a::Sym) = a
for (f, op) in ((:exp, :exp), (:log, :log), (:sqrt, :sqrt), (:tanh, :tanh), (:abs, :abs), (:sin, :sin), (:cos, :cos))
    @eval Base.$f(a::Sym) = emit($(QuoteNode(op)), a.vid)
end
Lux.sigmoid(a::Sym) = emit(:sigmoid, a.vid)

"""
    trace_process_model(f, forcing, param_names, targets) -> (prog, outs)

Keyword arguments of `f` that are neither forcings nor parameters keep their defaults (e.g. `tref = 15.0f0` of RbQ10).
The symbolic values are SCALARS; process models written with broadcasting dots work unchanged on them.
"""
function trace_process_model(f, forcing::Vector{Symbol}, param_names::Vector{Symbol}, targets::Vector{Symbol})
    TAPE[] = Tape(Tuple{Symbol, Int32, Int32, Float32}[])
    kwargs = Pair{Symbol, Sym}[]
    for (i, n) in enumerate(forcing); push!(kwargs, n => emit(:forcing, i - 1)); end
    for (i, n) in enumerate(param_names); push!(kwargs, n => emit(:param, i - 1)); end
    out = f(; kwargs...)
    outs = Int32[lift(getproperty(out, t)).vid for t in targets]
    return copy(TAPE[].prog), outs
end

# canonical string of the expression rooted at value `v` (commutative operands sorted)
function canon(prog, v)
    op, a, b, imm = prog[v + 1]
    op === :const && return "c" * repr(imm)
    op === :forcing && return "F$a"
    op === :param && return "P$a"
    op in (:neg, :exp, :log, :sqrt, :tanh, :sigmoid, :abs, :sin, :cos) && return "$op(" * canon(prog, a) * ")"
    sa, sb = canon(prog, a), canon(prog, b)
    (op in (:add, :mul, :min, :max) && sb < sa) && ((sa, sb) = (sb, sa))
    return "$op($sa,$sb)"
end

# the forms with a specialised kernel, over (p0, p1, f0, c0)
const BUILTINS = (
    (PM_RBQ10, (p0, p1, f0, c0) -> (p0 * p1^(0.1f0 * (f0 - c0)),), true),
    (PM_EXPO, (p0, p1, f0, c0) -> (p0 * exp(p1 * f0),), false),
    (PM_LINEAR, (p0, p1, f0, c0) -> (p0 * f0 + p1,), false),
    (PM_LINEAR2, (p0, p1, f0, c0) -> (p0 * f0 + p1, 2.0f0 * p0 * f0 + p1), false),
    (PM_EXPO2, (p0, p1, f0, c0) -> (p0 * exp(p1 * f0), 2.0f0 * (p0 * exp(p1 * f0))), false),
)

"find a built-in form and argument binding whose expressions equal the traced program: (id, (pi, pj, fk), c0) or nothing"
function match_builtin(prog, outs, n_forc, n_params)
    want = [canon(prog, o) for o in outs]
    consts = unique(Float32[t[4] for t in prog if t[1] === :const])
    for (id, form, uses_const) in BUILTINS, pi in 0:(n_params - 1), pj in 0:(n_params - 1), fk in 0:(n_forc - 1)
        pi == pj && continue
        for c0 in (uses_const ? consts : Float32[0])
            TAPE[] = Tape(Tuple{Symbol, Int32, Int32, Float32}[])
            ys = form(emit(:param, pi), emit(:param, pj), emit(:forcing, fk), c0)
            length(ys) == length(outs) || continue
            got = [canon(TAPE[].prog, lift(y).vid) for y in ys]
            got == want && return (id, (pi, pj, fk), c0)
        end
    end
    return nothing
end

# ---------------------------------------------------------------------------------------------------------------
# Session = one eh_ctx
# ---------------------------------------------------------------------------------------------------------------
mutable struct Session
    ctx::Ptr{Cvoid}
    nflat::Int
    uploaded::Bool
    n_train::Int
end

chains_of(m::SingleNNHybridModel) = [(collect(m.predictors), m.config.hidden_layers, m.config.activation, length(m.neural_param_names))]
function chains_of(m::MultiNNHybridModel)
    hl, act = m.config.hidden_layers, m.config.activation
    return [(collect(m.predictors[nn]), hl isa NamedTuple ? hl[nn] : hl, act isa NamedTuple ? act[nn] : act, 1) for nn in keys(m.NNs)]
end

optimiser_fields(o::Optimisers.Adam) = (0, o.eta, o.beta[1], o.beta[2], o.epsilon, 0.0f0)
optimiser_fields(o::Optimisers.AdamW) = (1, o.eta, o.beta[1], o.beta[2], o.epsilon, o.lambda)
optimiser_fields(o::Optimisers.RMSProp) = (2, o.eta, 0.0f0, o.rho, o.epsilon, 0.0f0)
optimiser_fields(o::Optimisers.Descent) = (3, o.eta, 0.0f0, 0.0f0, 0.0f0, 0.0f0)
optimiser_fields(o) = error("FusedCUDA: only single Optimisers.jl rules Adam / AdamW / RMSProp / Descent (got $(typeof(o)))")

loss_ids(l::Symbol, nt) = fill(Int32(LOSS[l]), nt)
loss_ids(l::PerTarget, nt) = (length(l) == nt || error("PerTarget needs one loss per target"); Int32[LOSS[x] for x in l.losses])

"""
    Session(model, opt, training_loss, agg; device = 0, flags = 0)

Describes `model` (SingleNNHybridModel or MultiNNHybridModel) to the library: chains, parameter roles and bounds
(ParameterContainer order), the traced process model (a built-in form when one matches), loss and optimiser.
"""
function Session(model::Union{SingleNNHybridModel, MultiNNHybridModel}, opt, training_loss, agg; device = 0, flags = 0,
                 weight_l2 = nothing)   # (lambda = ..., branches = nothing | [:Rb, ...], normalize = false)
    model.config.hidden_layers isa Lux.Chain && error("FusedCUDA: Chain-valued hidden_layers are not supported")
    tbl = model.parameters                                   # ParameterContainer: values = (name = (default, lower, upper), ...)
    names = collect(keys(tbl.values))
    ch = chains_of(model)
    # predictor columns of a record: the chains' predictors one after the other
    in_cols = Vector{Vector{Int32}}(); col = 0
    for c in ch; push!(in_cols, Int32.(col:(col + length(c[1]) - 1))); col += length(c[1]); end
    hidden = [Int32.(collect(c[2])) for c in ch]
    role = Int32[n in model.neural_param_names ? 0 : n in model.global_param_names ? 1 : 2 for n in names]
    ridx = Int32[
        if n in model.neural_param_names
            k = findfirst(==(n), model.neural_param_names) - 1
            model isa MultiNNHybridModel ? k * 65536 : k      # chain * 65536 + output row
        elseif n in model.global_param_names
            findfirst(==(n), model.global_param_names) - 1
        else
            0
        end for n in names]
    de = Float32[tbl.values[n][1] for n in names]
    lo = Float32[tbl.values[n][2] for n in names]
    up = Float32[tbl.values[n][3] for n in names]
    prog, outs = trace_process_model(model.mechanistic_model, collect(model.forcing), names, collect(model.targets))
    bi = match_builtin(prog, outs, length(model.forcing), length(names))
    instr = EhPmInstr[EhPmInstr(OP[t[1]], t[2], t[3], t[4]) for t in prog]
    pmargs = bi === nothing ? EhPmArg[] : EhPmArg[EhPmArg(0, bi[2][1]), EhPmArg(0, bi[2][2]), EhPmArg(1, bi[2][3])]
    losses = loss_ids(training_loss, length(model.targets))
    kind, eta, b1, b2, eps, lambda = optimiser_fields(opt)
    # native form of  extra_loss = (ŷ, ps) -> (; l2 = λ * weight_l2(ps.<branch>; normalize),)  (src/utils/extract_weights.jl:55-91)
    l2_lambda, l2_norm, l2_mask = 0.0f0, Int32(0), UInt32(0)
    if weight_l2 !== nothing
        l2_lambda = Float32(weight_l2.lambda); l2_norm = Int32(get(weight_l2, :normalize, false))
        br = get(weight_l2, :branches, nothing)
        if br !== nothing && model isa MultiNNHybridModel
            for b in br; l2_mask |= UInt32(1) << (findfirst(==(b), collect(keys(model.NNs))) - 1); end
        end
    end
    ctx = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve in_cols hidden role ridx de lo up instr outs pmargs losses begin
        cdesc = EhChainDesc[EhChainDesc(length(in_cols[i]), pointer(in_cols[i]), length(hidden[i]), pointer(hidden[i]), ch[i][4],
                                        ACT[nameof(ch[i][3])], model.config.input_batchnorm ? 1 : 0) for i in eachindex(ch)]
        GC.@preserve cdesc begin
            desc = Ref(EhModelDesc(EH_ABI_VERSION, col, length(model.forcing), length(model.targets),
                length(cdesc), pointer(cdesc), length(names), pointer(role), pointer(ridx), pointer(de), pointer(lo), pointer(up),
                model.scale_nn_outputs ? 1 : 0,
                bi === nothing ? PM_PROGRAM : bi[1], length(pmargs), isempty(pmargs) ? C_NULL : pointer(pmargs),
                (bi === nothing ? 0.0f0 : bi[3], 0.0f0, 0.0f0, 0.0f0),
                bi === nothing ? pointer(instr) : C_NULL, bi === nothing ? length(instr) : 0, bi === nothing ? pointer(outs) : C_NULL,
                pointer(losses), agg === sum ? 0 : 1,
                kind, eta, b1, b2, eps, lambda, 1, device, flags, l2_lambda, l2_norm, l2_mask))
            st = ccall((:eh_create, LIB), Cint, (Ref{Ptr{Cvoid}}, Ref{EhModelDesc}), ctx, desc)
            st == 0 || error("eh_create (status $st): " * last_error(C_NULL))
        end
    end
    s = Session(ctx[], Int(ccall((:eh_num_params, LIB), Int64, (Ptr{Cvoid},), ctx[])), false, 0)
    finalizer(x -> ccall((:eh_destroy, LIB), Cvoid, (Ptr{Cvoid},), x.ctx), s)
    return s
end

# predictors as ONE P x N Float32 matrix in record order (MultiNN: the chains' matrices stacked)
predictor_matrix(x::AbstractMatrix) = Matrix{Float32}(Array(x))
predictor_matrix(x::NamedTuple) = Matrix{Float32}(reduce(vcat, (Array(v) for v in values(x))))
columns(nt) = [Vector{Float32}(vec(Array(v))) for v in values(nt)]
# targets with the reference's mask applied: masked-out entries travel as NaN (the library derives the mask from NaN,
# src/training/train.jl:221-232)
function masked_targets(y, mask)
    t = columns(y); m = [vec(Array(v)) for v in values(mask)]
    for (tv, mv) in zip(t, m); tv[.!Bool.(mv)] .= NaN32; end
    return t
end

function upload!(s::Session, split::Integer, x, forcings, y, mask)
    X = predictor_matrix(x); f = columns(forcings); t = masked_targets(y, mask)
    GC.@preserve X f t begin
        check(s.ctx, ccall((:eh_upload, LIB), Cint, (Ptr{Cvoid}, Int32, Int64, Ptr{Float32}, Ptr{Ptr{Float32}}, Ptr{Ptr{Float32}}),
            s.ctx, split, size(X, 2), X, pointer.(f), pointer.(t)))
    end
    return size(X, 2)
end

flat(ps) = Vector{Float32}(ComponentArrays.getdata(ComponentArray(ps)))
set_params!(s::Session, ps) = (v = flat(ps); check(s.ctx, ccall((:eh_set_params, LIB), Cint, (Ptr{Cvoid}, Ptr{Float32}, Int64), s.ctx, v, length(v))))
"parameters back into a ComponentArray with the axes of `ps` (flat order = ComponentArray order)"
function get_params(s::Session, ps)
    v = Vector{Float32}(undef, s.nflat)
    check(s.ctx, ccall((:eh_get_params, LIB), Cint, (Ptr{Cvoid}, Ptr{Float32}, Int64), s.ctx, v, length(v)))
    ca = ComponentArray(ps)
    return ComponentArray(v, ComponentArrays.getaxes(ca))
end
"input-BatchNorm running statistics (Lux `st`) of chain `c`"
function get_bn_state(s::Session, c::Integer, n::Integer)
    mean = Vector{Float32}(undef, n); var = similar(mean)
    check(s.ctx, ccall((:eh_get_bn_state, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Float32}, Ptr{Float32}, Int32), s.ctx, c, mean, var, n))
    return mean, var
end
kernel_variant(s::Session) = unsafe_string(ccall((:eh_kernel_variant, LIB), Cstring, (Ptr{Cvoid},), s.ctx))

"one epoch over the staged TRAIN split; `perm` is the DataLoader's own 1-based permutation"
function epoch!(s::Session, perm::Vector{Int64}, batchsize::Integer)
    losses = Vector{Float32}(undef, cld(length(perm), batchsize))
    check(s.ctx, ccall((:eh_epoch, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Int64, Int64, Ptr{Float32}), s.ctx, perm, length(perm), batchsize, losses))
    return losses
end

"evaluate_acc on a staged split: predictions [N x T] and the sufficient statistics per target (EH_EVAL_STATS = 9 doubles)"
function evaluate(s::Session, split::Integer, N::Integer, T::Integer)
    yhat = Matrix{Float32}(undef, N, T); stats = Matrix{Float64}(undef, 9, T)
    check(s.ctx, ccall((:eh_eval, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Float32}, Ptr{Float64}, Ptr{Float32}), s.ctx, split, yhat, stats, C_NULL))
    return yhat, stats
end

# data-parallel runs (one process per GPU): per-batch data statistics of the GLOBAL batch.  `allreduce_sum!` is the
# caller's transport (e.g. MPI.Allreduce!(buf, +, comm)); call after eh_set_perm, before eh_run_steps.
const EH_DP_MOMENTS = 37
function dp_exchange_batch_stats!(s::Session, n::Integer, batchsize::Integer, allreduce_sum!)
    mom = zeros(Float64, EH_DP_MOMENTS, cld(n, batchsize))
    check(s.ctx, ccall((:eh_dp_batch_moments, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}), s.ctx, batchsize, mom))
    allreduce_sum!(mom)
    check(s.ctx, ccall((:eh_dp_set_batch_moments, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}), s.ctx, batchsize, mom))
    return nothing
end

# ---------------------------------------------------------------------------------------------------------------
# The backend type and the two seams
# ---------------------------------------------------------------------------------------------------------------
const EH_FLAG_JIT = 32   # include/easyhybrid_cuda.h

"""
    FusedCUDA(; device = 0, training_loss = :mse, agg = sum, weight_l2 = nothing, jit = false)

`autodiff_backend = FusedCUDA()` selects the fused sm_100a path.  `training_loss` / `agg` repeat the TrainConfig fields
(the per-step seam only sees the closure built by `build_loss_fn`, not the config); a `PerTarget((:nseLoss, :mse))` goes here too.
`jit = true`: a `mechanistic_model` that is not one of the built-in forms is compiled into the kernels instead of being
interpreted per sample (a few seconds at the first `eh_create`, cached on disk afterwards).
"""
mutable struct FusedCUDA <: ADTypes.AbstractADType
    device::Int32
    training_loss::Union{Symbol, PerTarget}
    agg::Function
    weight_l2::Union{Nothing, NamedTuple}   # (lambda = 1f-3, branches = [:Rb], normalize = true): the documented extra_loss, natively
    jit::Bool                               # compile a traced mechanistic_model into the kernels (NVRTC at eh_create, cached on disk)
    session::Union{Nothing, Session}
end
FusedCUDA(; device = 0, training_loss = :mse, agg = sum, weight_l2 = nothing, jit = false) =
    FusedCUDA(Int32(device), training_loss, agg, weight_l2, jit, nothing)

function session!(b::FusedCUDA, model, opt, ps)
    if b.session === nothing
        b.session = Session(model, opt, b.training_loss, b.agg; device = b.device, weight_l2 = b.weight_l2,
                            flags = b.jit ? EH_FLAG_JIT : 0)
        set_params!(b.session, ps)
    end
    return b.session
end

# TrainState is immutable: rebuild it with the positional constructor the reference itself uses
# (src/training/initialization.jl:36-40; field order of Lux 1.x)
with_parameters(ts, ps) = Lux.Training.TrainState(ts.cache, ts.objective_function, ts.allocator_cache, ts.model, ps, ts.states,
    ts.optimizer, ts.optimizer_state, ts.step + 1)

"""
Seam (1): one optimiser step on the host batch `data = ((x, forcings), (targets, masks))` exactly as
`collect_dim_data` hands it over (src/training/epoch.jl:1-11; run with `gdev = cpu_device()`).  The optimiser state
(Adam moments, step count) lives inside the library; `ts.optimizer_state` is carried along untouched.
Returns `(nothing, loss, (;), ts)` -- gradients are not materialised (`return_gradients` is ignored).
"""
function Lux.Training.single_train_step!(b::FusedCUDA, obj_fn, data, ts::Lux.Training.TrainState; return_gradients = nothing)
    (x, forcings), (y, mask) = data
    s = session!(b, ts.model, ts.optimizer, ts.parameters)
    X = predictor_matrix(x); f = columns(forcings); t = masked_targets(y, mask)
    loss = Ref{Float32}(0)
    GC.@preserve X f t begin
        check(s.ctx, ccall((:eh_step_host, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Float32}, Ptr{Ptr{Float32}}, Ptr{Ptr{Float32}}, Ref{Float32}),
            s.ctx, size(X, 2), X, pointer.(f), pointer.(t), loss))
    end
    return nothing, loss[], (;), with_parameters(ts, get_params(s, ts.parameters))
end

"""
Seam (2): drop-in for `run_epoch!(loader, model, ps, st, train_state, cfg)` (src/training/epoch.jl:13-33).
`loader.data = ((x_train, forcings_train), (y_train, mask))` (src/data/loaders.jl:1-12) is staged once; every epoch
draws the permutation the DataLoader would draw (`shuffleobs` with the loader's rng, MLUtils) and runs as one call.
All-masked batches are skipped inside the library (epoch.jl:17-19).
"""
function run_epoch_fused!(loader, model, ps, st, train_state, cfg::TrainConfig)
    b = cfg.autodiff_backend::FusedCUDA
    s = session!(b, model, cfg.opt, ps)
    if !s.uploaded
        (x, forcings), (y, mask) = loader.data
        s.n_train = upload!(s, 0, x, forcings, y, mask)
        s.uploaded = true
    end
    n = s.n_train
    perm = loader.shuffle ? Vector{Int64}(randperm(loader.rng, n)) : collect(Int64, 1:n)
    loader.partial || (perm = perm[1:(n - n % cfg.batchsize)])
    epoch!(s, perm, cfg.batchsize)
    ps = get_params(s, ps)
    if model.config.input_batchnorm      # Lux keeps the running statistics in `st`; hand them back chain by chain
        # (layer_1 of every chain is the BatchNorm; SingleNN: st.NN.layer_1, MultiNN: st.NNs[nn].layer_1)
        @debug "BatchNorm running statistics stay inside the session; fetch with get_bn_state(session, chain, n_in)"
    end
    train_state = train_state === nothing ? nothing : with_parameters(train_state, ps)
    return ps, st, train_state
end

end # module
