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
#include "eh_chunk.cuh"

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
};

template <class C>
__host__ __device__ constexpr int step_smem_floats(int nwarps)
{
    return rup4(C::NW) + SS_FLOATS + nwarps * C::STAGE_FLOATS;
}

template <class C>
__device__ __forceinline__ void fetch_record(const float4* rec, const int* idx, long long rec_base, int B, int chunk,
                                             int nchunks, int lane, float4* r, bool& valid)
{
    const int s = chunk * CHUNK + lane;
    valid = chunk < nchunks && s < B;
    long long i = rec_base + s;
    if (idx && valid) i = idx[s];
#pragma unroll
    for (int q = 0; q < C::R4 / 4; q++)
        r[q] = valid ? __ldg(rec + i * (C::R4 / 4) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
}

template <class C>
__global__ void __launch_bounds__(512, 1) k_step(const StepArgs a)
{
    extern __shared__ float4 smem4[];
    float* sW = reinterpret_cast<float*>(smem4);
    float* sS = sW + rup4(C::NW);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    float* stage = sS + SS_FLOATS + warp * C::STAGE_FLOATS;

    // prologue part 1 (independent of the previous step's update): fetch my sample, constant rows
    const int GW = gridDim.x * nwarps;
    const int nchunks = (a.B + CHUNK - 1) / CHUNK;
    int chunk = blockIdx.x * nwarps + warp;
    float4 r[C::R4 / 4];
    bool valid;
    fetch_record<C>(a.rec, a.idx, a.rec_base, a.B, chunk, nchunks, lane, r, valid);
    init_stage_rows<C>(stage, lane);
    int rowD[C::NBI], rowA[C::NBI];
    tile_rows<C>(lane, rowD, rowA);

    // wait for the previous update kernel (PDL), then pull weights and scalars
    pdl_wait();
    load_weights_and_scalars<C>(a.pblock, a.nflat, a.wsrc, a.bscal, a.use_bn, sW, sS);
    __syncthreads();
    pdl_launch_dependents();

    PmCtx cx;
    cx.pms = sS + SS_PMS;
    cx.c = a.pmc;
    cx.uniform_mask = 0;
#pragma unroll
    for (int s = 0; s < MAXPS; s++)
        if (s >= C::NPS || a.slot[s].role != ROLE_NEURAL) cx.uniform_mask |= 1u << s;

    float2 acc[C::NBI][16];
#pragma unroll
    for (int i = 0; i < C::NBI; i++)
#pragma unroll
        for (int e = 0; e < 16; e++) acc[i][e] = f2s(0.f);
    ChunkStats st;
#pragma unroll
    for (int t = 0; t < MAXT; t++) st.loss[t] = 0.f;
#pragma unroll
    for (int s = 0; s < MAXPS; s++) st.gphi[s] = 0.f;
    LastAcc<C> la;
    la.zero();

    for (; chunk < nchunks; chunk += GW) {
        float rec[C::R4];
#pragma unroll
        for (int q = 0; q < C::R4 / 4; q++) {
            rec[4 * q] = r[q].x; rec[4 * q + 1] = r[q].y; rec[4 * q + 2] = r[q].z; rec[4 * q + 3] = r[q].w;
        }
        const bool v = valid;
        fetch_record<C>(a.rec, a.idx, a.rec_base, a.B, chunk + GW, nchunks, lane, r, valid);  // prefetch
        chunk_sample_phase<C>(rec, v, sW, sS, stage, lane, a.slot, a.loss_kind, cx, st, la);
        __syncwarp();
        chunk_dw_phase<C>(stage, lane, rowD, rowA, acc);
        __syncwarp();
    }
    __syncthreads();  // every warp is done with its staging tile; reuse it as [nwarps][NPART]
    cta_reduce<C>(acc, st, la, sS + SS_FLOATS, a.partial + (size_t)blockIdx.x * a.npart, 1);
}

}  // namespace eh
