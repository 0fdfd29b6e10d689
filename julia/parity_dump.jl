# parity_dump.jl -- run the REAL EasyHybrid.jl train step on the CPU and dump what this repo's oracle
# restates, so that the "parity unpinned" items of DESIGN.md section 3 can be pinned by anyone with Julia:
#   julia --project=/path/to/EasyHybrid.jl julia/parity_dump.jl out.bin
# Layout of out.bin (little endian): Int64 n, B, nsteps, nflat; Float32 X[2,n], ta[n], reco[n];
# Float32 ps0[nflat]; Int64 perm[nsteps*B]; then per step: Float32 loss, Float32 grad[nflat], Float32 ps[nflat].
# tests/test_julia_dump.py (skipped when the file is absent) replays it through the oracle and the GPU.
using EasyHybrid, Lux, Random, Zygote, ComponentArrays, Optimisers, Statistics

function RbQ10(; ta, Q10, rb, tref = 15.0f0)
    reco = rb .* Q10 .^ (0.1f0 .* (ta .- tref))
    return (; reco, Q10, rb)
end

function main(path)
    rng = MersenneTwister(42); n = 4096; B = 512; nsteps = 8
    ta = Float32.(10 .+ 10 .* randn(rng, n)); sw = Float32.(abs.(50 .+ 20 .* randn(rng, n)))
    dsw = Float32.(vcat(0.0, diff(sw)))
    reco = Float32.((3 .+ 0.02 .* (sw .- mean(sw))) .* 2 .^ (0.1 .* (ta .- 15)) .+ 0.1 .* randn(rng, n))
    model = constructHybridModel([:sw_pot, :dsw_pot], [:ta], [:reco], RbQ10,
        (rb = (3.0f0, 0.0f0, 13.0f0), Q10 = (2.0f0, 1.0f0, 4.0f0)), [:rb], [:Q10];
        hidden_layers = [16, 16], activation = tanh, scale_nn_outputs = true)
    ps, st = LuxCore.setup(rng, model); ps = ComponentArray(ps)
    X = permutedims(hcat(sw, dsw)); perm = randperm(rng, n)[1:(nsteps * B)]
    ts = Lux.Training.TrainState(model, ps, st, Adam(0.01f0))
    logging = EasyHybrid.LoggingLoss(train_mode = true, training_loss = :mse, loss_types = [:mse], agg = sum)
    lossf = (m, p, s, d) -> EasyHybrid.compute_loss(m, p, s, d; logging)
    open(path, "w") do io
        write(io, Int64[n, B, nsteps, length(ps)]); write(io, X); write(io, ta); write(io, reco)
        write(io, collect(ps)); write(io, Int64.(perm))
        for k in 1:nsteps
            idx = perm[((k - 1) * B + 1):(k * B)]
            d = ((X[:, idx], (; ta = ta[idx])), ((; reco = reco[idx]), (; reco = .!isnan.(reco[idx]))))
            g, l, _, ts = Lux.Training.single_train_step!(AutoZygote(), lossf, d, ts)
            write(io, Float32(l)); write(io, Float32.(collect(g))); write(io, Float32.(collect(ts.parameters)))
        end
    end
end
main(ARGS[1])
