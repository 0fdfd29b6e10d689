# bench_reference.jl -- time the real reference train() on host cores for the C3-shaped workload
# (fill in BASELINE.md with the number; this repo can only time its C restatement):
#   julia -t auto --project=/path/to/EasyHybrid.jl julia/bench_reference.jl
using EasyHybrid, Lux, Random, DataFrames, Statistics
include("parity_dump.jl") # RbQ10
n = 2^20; rng = MersenneTwister(42)
ta = Float32.(10 .+ 10 .* randn(rng, n)); sw = Float32.(abs.(50 .+ 20 .* randn(rng, n)))
df = DataFrame(; ta, sw_pot = sw, dsw_pot = Float32.(vcat(0.0, diff(sw))),
               reco = Float32.((3 .+ 0.02 .* (sw .- mean(sw))) .* 2 .^ (0.1 .* (ta .- 15)) .+ 0.1 .* randn(rng, n)))
model = constructHybridModel([:sw_pot, :dsw_pot], [:ta], [:reco], RbQ10,
    (rb = (3.0f0, 0.0f0, 13.0f0), Q10 = (2.0f0, 1.0f0, 4.0f0)), [:rb], [:Q10];
    hidden_layers = [16, 16], activation = tanh, scale_nn_outputs = true)
kw = (; nepochs = 1, batchsize = 65536, plotting = false, show_progress = false, save_training = false,
      keep_history = false, gdev = cpu_device())
train(model, df, (); kw...)                                   # compile
t = @elapsed train(model, df, (); kw..., nepochs = 3)
println("reference CPU train(): ", 3 * 0.8 * n / t, " samples/s on ", Threads.nthreads(), " threads (includes per-epoch evaluation)")
