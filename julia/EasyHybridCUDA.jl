# EasyHybridCUDA.jl -- reference-side binding of libeasyhybrid_cuda.so.
#
# NOT EXECUTED IN THIS REPO'S CI: Julia is not installed in the build image (SURVEY.md section 0).
# It documents, in runnable form, the `ccall` layer a maintainer adds to EasyHybrid.jl so that
#     train(model, data; autodiff_backend = FusedCUDA())
# takes the fused path.  The same call sequence is exercised by the Python host mirror through
# ctypes (easyhybrid.jl_b200/session.py), which is what the tests run.
#
# Seams used (EasyHybrid.jl v0.2.0):
#   TrainConfig.autodiff_backend            src/config/TrainingConfig.jl:52
#   run_epoch!(loader, model, ps, st, train_state, cfg)   src/training/epoch.jl:13-33
#   evaluate_acc                            src/training/train.jl:347-355
module EasyHybridCUDA

using EasyHybrid
using EasyHybrid: SingleNNHybridModel, MultiNNHybridModel, TrainConfig, default, lower, upper, pnames
import EasyHybrid: run_epoch!
using ComponentArrays

const LIB = get(ENV, "EASYHYBRID_CUDA_LIB", "libeasyhybrid_cuda.so")

"`autodiff_backend = FusedCUDA()` selects the fused sm_100a path"
struct FusedCUDA
    device::Int32
end
FusedCUDA() = FusedCUDA(0)

# ---- C structs (include/easyhybrid_cuda.h) --------------------------------------------------
struct EhPmArg; kind::Int32; index::Int32; end
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
    pm_prog::Ptr{Cvoid}; pm_len::Int32; pm_outputs::Ptr{Int32}
    loss_per_target::Ptr{Int32}; agg::Int32
    opt_kind::Int32; eta::Float32; beta1::Float32; beta2::Float32; eps::Float32; lambda::Float32
    adamw_decay_coupled_eta::Int32
    device::Int32; flags::Int32
end

const ACT = Dict(:identity => 0, :tanh => 1, :tanh_fast => 1, :sigmoid => 2, :sigmoid_fast => 2, :σ => 2, :relu => 3, :swish => 4)
const LOSS = Dict(:mse => 0, :rmse => 1, :mae => 2, :nseLoss => 3)

check(ctx, st) = st == 0 || error("libeasyhybrid_cuda: status $st: " *
    unsafe_string(ccall((:eh_last_error, LIB), Cstring, (Ptr{Cvoid},), ctx)))

mutable struct Session
    ctx::Ptr{Cvoid}
    nflat::Int
end

"""
    Session(model, cfg, pm; device) -> Session

`pm = (id, (param_i, param_j, forcing_k), consts)` names the built-in process-model form
(EH_PM_RBQ10 = 0, EH_PM_EXPO = 1, EH_PM_LINEAR = 2, EH_PM_LINEAR2 = 3) and its argument binding;
`trace_process_model` in the Python mirror shows how the form is recognised by tracing
`model.mechanistic_model` with a symbolic number type.
"""
function Session(model::SingleNNHybridModel, cfg::TrainConfig, pm; device = 0)
    names = collect(pnames(model.parameters))
    role = Int32[n in model.neural_param_names ? 0 : n in model.global_param_names ? 1 : 2 for n in names]
    ridx = Int32[n in model.neural_param_names ? findfirst(==(n), model.neural_param_names) - 1 :
                 n in model.global_param_names ? findfirst(==(n), model.global_param_names) - 1 : 0 for n in names]
    de = Float32[default(model.parameters)[n] for n in names]
    lo = Float32[lower(model.parameters)[n] for n in names]
    up = Float32[upper(model.parameters)[n] for n in names]
    hidden = Int32.(model.config.hidden_layers)
    in_cols = Int32.(0:(length(model.predictors) - 1))
    pmargs = [EhPmArg(0, pm[2][1]), EhPmArg(0, pm[2][2]), EhPmArg(1, pm[2][3])]
    losses = fill(Int32(LOSS[cfg.training_loss]), length(model.targets))
    opt = cfg.opt
    kind = opt isa EasyHybrid.Adam ? 0 : opt isa EasyHybrid.AdamW ? 1 : opt isa EasyHybrid.RMSProp ? 2 : 3
    ctx = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve role ridx de lo up hidden in_cols pmargs losses begin
        chain = [EhChainDesc(length(in_cols), pointer(in_cols), length(hidden), pointer(hidden),
                             length(model.neural_param_names), ACT[nameof(model.config.activation)],
                             model.config.input_batchnorm)]
        GC.@preserve chain begin
            desc = Ref(EhModelDesc(1, length(model.predictors), length(model.forcing), length(model.targets),
                1, pointer(chain), length(names), pointer(role), pointer(ridx), pointer(de), pointer(lo), pointer(up),
                model.scale_nn_outputs, pm[1], 3, pointer(pmargs), (Float32.(pm[3])..., ntuple(_ -> 0f0, 4 - length(pm[3]))...),
                C_NULL, 0, C_NULL, pointer(losses), cfg.agg === sum ? 0 : 1,
                kind, opt.eta, kind == 2 ? 0f0 : opt.beta[1], kind == 2 ? opt.rho : opt.beta[2], opt.epsilon,
                kind == 1 ? opt.lambda : 0f0, 1, device, 0))
            st = ccall((:eh_create, LIB), Cint, (Ref{Ptr{Cvoid}}, Ref{EhModelDesc}), ctx, desc)
            st == 0 || error("eh_create: " * unsafe_string(ccall((:eh_last_error, LIB), Cstring, (Ptr{Cvoid},), C_NULL)))
        end
    end
    s = Session(ctx[], Int(ccall((:eh_num_params, LIB), Int64, (Ptr{Cvoid},), ctx[])))
    finalizer(x -> ccall((:eh_destroy, LIB), Cvoid, (Ptr{Cvoid},), x.ctx), s)
    return s
end

"stage one split: X is the P x N Matrix{Float32} of prepare_data, forcings / targets NamedTuples of Vector{Float32}"
function upload!(s::Session, split::Integer, X::Matrix{Float32}, forcings::NamedTuple, targets::NamedTuple)
    f = collect(values(forcings)); t = collect(values(targets))
    GC.@preserve X f t begin
        check(s.ctx, ccall((:eh_upload, LIB), Cint,
            (Ptr{Cvoid}, Int32, Int64, Ptr{Float32}, Ptr{Ptr{Float32}}, Ptr{Ptr{Float32}}),
            s.ctx, split, size(X, 2), X, pointer.(f), pointer.(t)))
    end
end

set_params!(s::Session, ps::ComponentVector{Float32}) =
    check(s.ctx, ccall((:eh_set_params, LIB), Cint, (Ptr{Cvoid}, Ptr{Float32}, Int64), s.ctx, getdata(ps), length(ps)))

function get_params!(s::Session, ps::ComponentVector{Float32})
    check(s.ctx, ccall((:eh_get_params, LIB), Cint, (Ptr{Cvoid}, Ptr{Float32}, Int64), s.ctx, getdata(ps), length(ps)))
    return ps
end

"run_epoch!: `perm` is the DataLoader's own permutation (1-based), so batch composition is the reference's"
function epoch!(s::Session, perm::Vector{Int64}, batchsize::Integer)
    losses = Vector{Float32}(undef, cld(length(perm), batchsize))
    check(s.ctx, ccall((:eh_epoch, LIB), Cint, (Ptr{Cvoid}, Ptr{Int64}, Int64, Int64, Ptr{Float32}),
        s.ctx, perm, length(perm), batchsize, losses))
    return losses
end

"evaluate_acc: sufficient statistics (n, Sy, Sh, Syy, Shh, Syh, SSE, SAE, shift) per target + predictions"
# data-parallel runs (one process per GPU): per-batch data statistics of the GLOBAL batch.  `allreduce_sum!` is the
# caller's transport (e.g. MPI.Allreduce!(buf, +, comm)); call after eh_set_perm, before eh_run_steps.
const EH_DP_MOMENTS = 37
function dp_exchange_batch_stats!(s::Session, n::Integer, batchsize::Integer, allreduce_sum!)
    nb = cld(n, batchsize)
    mom = zeros(Float64, EH_DP_MOMENTS, nb)
    check(s.ctx, ccall((:eh_dp_batch_moments, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}), s.ctx, batchsize, mom))
    allreduce_sum!(mom)
    check(s.ctx, ccall((:eh_dp_set_batch_moments, LIB), Cint, (Ptr{Cvoid}, Int64, Ptr{Float64}), s.ctx, batchsize, mom))
    return nothing
end

"Name of the compiled kernel family that serves this model (specialised fp32, generic fp32, or bf16 tensor-core path)."
kernel_variant(s::Session) = unsafe_string(ccall((:eh_kernel_variant, LIB), Cstring, (Ptr{Cvoid},), s.ctx))

function evaluate(s::Session, split::Integer, N::Integer, T::Integer)
    yhat = Matrix{Float32}(undef, N, T); stats = Matrix{Float64}(undef, 9, T)
    check(s.ctx, ccall((:eh_eval, LIB), Cint, (Ptr{Cvoid}, Int32, Ptr{Float32}, Ptr{Float64}, Ptr{Float32}),
        s.ctx, split, yhat, stats, C_NULL))
    return yhat, stats
end

# The seam: a method of the reference's own hot loop for the fused backend.  `loader.data` holds
# ((x_train, forcings_train), (y_train, mask)); MLUtils draws `randperm(rng, n)` per epoch, which is
# reproduced here with the same rng so that batches are bit-identical to the stock path.
function run_epoch!(loader, model, ps, st, train_state, cfg::TrainConfig{<:Any}) where {}
    cfg.autodiff_backend isa FusedCUDA || return invoke(run_epoch!, Tuple{Any, Any, Any, Any, Any, TrainConfig}, loader, model, ps, st, train_state, cfg)
    s = session_for(model, cfg, loader)              # cached: created + uploaded on first use
    set_params!(s, ps)
    perm = collect(Int64, MLUtils.shuffleobs(loader.rng, 1:MLUtils.numobs(loader.data)).indices)
    epoch!(s, perm, cfg.batchsize)
    get_params!(s, ps)
    return ps, st, train_state
end

end # module
