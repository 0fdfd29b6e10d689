// eh_step_kernel.cuh -- K1: the fused hybrid training-step kernel, one launch per minibatch.
//
// Dense-chain forward -> parameter squashing -> process model -> masked residual ->
// hand-derived backward -> per-CTA partial sums of the (already loss-scaled) gradient and of the
// loss statistics.  K2 (eh_update_kernel.cuh) reduces the partials in fixed order and applies
// the optimiser.  Replaces, per step, the call
//   Lux.Training.single_train_step!(backend, loss_fn, batch, train_state)
// at src/training/epoch.jl:20-26 (forward: src/models/GenericHybridModel.jl:370-431,
// loss: src/losses/compute_loss.jl:20-35, src/losses/loss_fn.jl:58-81).
// The per-chunk device code lives in eh_chunk.cuh; the persistent multi-step form of the same
// computation is eh_epoch_kernel.cuh.
#pragma once
#include "eh_engine_ffma.cuh"
#include "eh_engine_mma.cuh"
#include "eh_engine_tc.cuh"

namespace eh {

struct StepArgs {
    const float4* rec;        // packed records, canonical order [x(P) | f(F) | y(T) | pad], R4 floats each
    const int* idx;           // 0-based sample ids of this batch (NULL: records rec_base .. rec_base+B)
    long long rec_base;
    int B;                    // samples in this batch
    const float* pblock;      // parameter block: flat theta/phi (reference ComponentArray order) + tail
    int nflat;
    const int* wsrc;          // [nweights] flat index feeding each smem weight cell, -1 = zero pad
    const float* bscal;       // per-batch scalar row (BS_* layout) of this batch
    float* partial;           // [gridDim.x][npart] per-CTA partial sums
    int npart;
    int loss_kind[MAXT];
    PSlot slot[MAXPS];
    float pmc[4];             // process-model constants
    int use_bn;
    const PmProgData* prog;   // traced process model (PmProgram variants), device memory
    int scale_rt;             // PmProgram variants: scale_nn_outputs
    unsigned pass_mask[3];    // PmProgram variants: pass-through units per hidden layer (PmCtx::pass)
};

// per-CTA work region (floats): staging tiles of all warps, reused as the [nwarps][NPART] reduction scratch
template <class E>
__host__ __device__ constexpr int work_floats(int nwarps)
{
    return nwarps * (E::STAGE_FLOATS > E::NPART ? E::STAGE_FLOATS : E::NPART);
}

template <class E>
__global__ void __launch_bounds__(E::MAX_WARPS * 32, 1) k_step(const StepArgs a)
{
    using C = typename E::Cfg;
    extern __shared__ float4 smem4[];
    float* sW = reinterpret_cast<float*>(smem4);
    float* sS = sW + rup4(C::NW);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    float* work = sS + SS_FLOATS;
    float* stage = work + warp * E::STAGE_FLOATS;

    // prologue part 1 (independent of the previous step's update): fetch my samples, constant rows
    const int GW = gridDim.x * nwarps;
    const int nchunks = (a.B + E::CHUNK - 1) / E::CHUNK;
    int chunk = blockIdx.x * nwarps + warp;
    typename E::State st;
    const FetchArgs fa{a.rec, a.idx, a.rec_base, a.B, nchunks, nullptr, 0};
    E::fetch(st, fa, chunk, lane);
    E::init_warp(st, stage, lane);

    // wait for the previous update kernel (PDL), then pull weights and scalars
    pdl_wait();
    load_weights_and_scalars<C>(a.pblock, a.nflat, a.wsrc, a.bscal, a.use_bn, sW, sS);
    __syncthreads();
    pdl_launch_dependents();

    PmCtx cx;
    cx.pms = sS + SS_PMS;
    cx.c = a.pmc;
    cx.prog = a.prog;
    cx.scale_rt = a.scale_rt;
    for (int l = 0; l < 3; l++) cx.pass[l] = a.pass_mask[l];
    cx.uniform_mask = 0;
    cx.phi_flag = nullptr;
    cx.phi_want = 0;
#pragma unroll
    for (int s = 0; s < MAXPS; s++)
        if (s >= C::NPS || a.slot[s].role != ROLE_NEURAL) cx.uniform_mask |= 1u << s;

    E::step_begin(st, sW, lane);
    for (; chunk < nchunks; chunk += GW) E::chunk(st, fa, chunk + GW, sW, sS, stage, lane, a.slot, a.loss_kind, cx);
    if (E::SCRATCH_ALIASES_STAGE) __syncthreads();  // every warp is done with its staging tile; the work region becomes [nwarps][NPART]
    E::reduce_prepare(st, work);
    __syncthreads();
    float* out = a.partial + (size_t)blockIdx.x * a.npart;
    for (int p = threadIdx.x; p < E::NPART; p += blockDim.x) __stcg(out + p, E::reduce_sum_at(work, nwarps, p));
}

}  // namespace eh
