// eh_epoch_kernel.cuh -- the persistent form of the training loop: MANY optimiser steps in one
// cooperative launch (one CTA per SM), replacing run_epoch! (src/training/epoch.jl:13-33) as a
// whole.  Per step every CTA
//   1. runs the fused forward/backward on its chunks (eh_chunk.cuh) with the weights it keeps in
//      shared memory,
//   2. publishes its partial vector, passes a grid barrier,
//   3. reduce-scatter: CTA c reduces slice c of the vector over all CTAs in a fixed order and
//      applies the optimiser to the parameters that live in that slice (their Adam moments stay in
//      the owner's shared memory for the whole launch),
//   4. publishes the new parameter values, passes a second grid barrier, reloads its weight image.
// Compared with one launch per step this removes two kernel launches (~2 x 5000 cycles of launch
// ramp), the dependent-load prologue and the single-CTA second pass from every step.
// No atomics on data: the only atomic is the barrier's arrival counter.
#pragma once
#include "eh_step_kernel.cuh"

namespace eh {

struct EpochArgs {
    const float4* rec;
    const int* idx;            // resident index stream
    long long n;               // its length
    int B;                     // nominal batch size
    long long first_step;
    int nsteps;
    int nb;                    // batches per pass = ceil(n / B)
    float* pblock;             // in/out: parameter block (flat + tail)
    int nflat, ntheta;
    float* m;                  // in/out optimiser moments
    float* v;
    OptState* ost;             // in/out
    const int* wsrc;           // [NW]
    const int* inv;            // [NPART] partial index -> flat parameter or -1
    const float* pspan;        // [nflat]
    const int* slot_of_flat;   // [nflat] phi entries: canonical slot (for the tail), -1 otherwise
    const float* bscal;        // [nb][BS_STRIDE]
    float* pbuf;               // [2][gridDim.x][npartp] published partial vectors
    float* pub;                // [2][nflat + PARAM_TAIL] published parameter blocks
    unsigned* counter;         // grid barrier arrival counter (zeroed before launch)
    float* stats_out;          // [nsteps][MAXT] reduced loss sums
    int npartp;                // padded partial length (multiple of 4)
    int SL;                    // slice of the partial vector owned by one CTA
    int T;
    int loss_kind[MAXT];
    PSlot slot[MAXPS];
    float pmc[4];
    int use_bn;
    int pm_id;
    int opt_kind, adamw_coupled;
    float eta, beta1, beta2, eps, lambda;
};

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p)
{
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// arrive + wait on a monotonically increasing counter; all threads of the CTA call it
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned target)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        while (ld_acquire_gpu(counter) < target) { }
        __threadfence();
    }
    __syncthreads();
}

template <class C>
__global__ void __launch_bounds__(512, 1) k_epoch(const EpochArgs a)
{
    extern __shared__ float4 smem4[];
    float* sW = reinterpret_cast<float*>(smem4);
    float* sS = sW + rup4(C::NW);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    float* stage0 = sS + SS_FLOATS;
    float* stage = stage0 + warp * C::STAGE_FLOATS;
    // owner state behind the staging tiles: [SL] theta, m, v, flat index
    float* own = stage0 + nwarps * C::STAGE_FLOATS;
    float* own_th = own;
    float* own_m = own + a.SL;
    float* own_v = own + 2 * a.SL;
    int* own_p = reinterpret_cast<int*>(own + 3 * a.SL);
    float* red = own + 4 * a.SL;  // [SL] reduced slice
    const int G = gridDim.x;
    const int q0 = blockIdx.x * a.SL;

    for (int j = threadIdx.x; j < a.SL; j += blockDim.x) {
        int q = q0 + j;
        int p = (q < C::NPART) ? a.inv[q] : -1;
        own_p[j] = p;
        own_th[j] = p >= 0 ? a.pblock[p] : 0.f;
        own_m[j] = p >= 0 ? a.m[p] : 0.f;
        own_v[j] = p >= 0 ? a.v[p] : 0.f;
    }
    float b1t = a.ost->b1t, b2t = a.ost->b2t;
    long long tdone = 0, tskip = 0;

    init_stage_rows<C>(stage, lane);
    int rowD[C::NBI], rowA[C::NBI];
    tile_rows<C>(lane, rowD, rowA);
    load_weights_and_scalars<C>(a.pblock, a.nflat, a.wsrc, nullptr, 0, sW, sS);

    PmCtx cx;
    cx.pms = sS + SS_PMS;
    cx.c = a.pmc;
    cx.uniform_mask = 0;
#pragma unroll
    for (int s = 0; s < MAXPS; s++)
        if (s >= C::NPS || a.slot[s].role != ROLE_NEURAL) cx.uniform_mask |= 1u << s;

    const int GW = G * nwarps;
    const int gw = blockIdx.x * nwarps + warp;
    float4 r[C::R4 / 4];
    bool valid;
    {
        long long b = a.first_step % a.nb;
        long long rem = a.n - b * a.B;
        int Bk = (int)(rem < a.B ? rem : a.B);
        fetch_record<C>(a.rec, a.idx + b * a.B, 0, Bk, gw, (Bk + CHUNK - 1) / CHUNK, lane, r, valid);
    }
    unsigned bar = 0;

    for (int s = 0; s < a.nsteps; s++) {
        const long long b = (a.first_step + s) % a.nb;
        const long long rem = a.n - b * a.B;
        const int Bk = (int)(rem < a.B ? rem : a.B);
        const int nchunks = (Bk + CHUNK - 1) / CHUNK;
        const float* bs = a.bscal + (size_t)b * BS_STRIDE;
        const int par = s & 1;
        // per-batch scalars
        if (threadIdx.x < MAXT) sS[SS_C + threadIdx.x] = bs[BS_C + threadIdx.x];
        if (threadIdx.x < 2 * C::P)
            sS[SS_BN + threadIdx.x] = a.use_bn ? bs[BS_BN + threadIdx.x] : ((threadIdx.x & 1) ? 1.f : 0.f);
        __syncthreads();

        float2 acc[C::NBI][16];
#pragma unroll
        for (int i = 0; i < C::NBI; i++)
#pragma unroll
            for (int e = 0; e < 16; e++) acc[i][e] = f2s(0.f);
        ChunkStats st;
#pragma unroll
        for (int t = 0; t < MAXT; t++) st.loss[t] = 0.f;
#pragma unroll
        for (int t = 0; t < MAXPS; t++) st.gphi[t] = 0.f;

        for (int chunk = gw; chunk < nchunks; chunk += GW) {
            float rec[C::R4];
#pragma unroll
            for (int q = 0; q < C::R4 / 4; q++) {
                rec[4 * q] = r[q].x; rec[4 * q + 1] = r[q].y; rec[4 * q + 2] = r[q].z; rec[4 * q + 3] = r[q].w;
            }
            const bool v = valid;
            fetch_record<C>(a.rec, a.idx + b * a.B, 0, Bk, chunk + GW, nchunks, lane, r, valid);
            chunk_sample_phase<C>(rec, v, sW, sS, stage, lane, a.slot, a.loss_kind, cx, st);
            __syncwarp();
            chunk_dw_phase<C>(stage, lane, rowD, rowA, acc);
            __syncwarp();
        }
        // prefetch my first sample of the next step: its latency hides behind the exchange below
        if (s + 1 < a.nsteps) {
            long long b2 = (a.first_step + s + 1) % a.nb;
            long long rem2 = a.n - b2 * a.B;
            int Bk2 = (int)(rem2 < a.B ? rem2 : a.B);
            fetch_record<C>(a.rec, a.idx + b2 * a.B, 0, Bk2, gw, (Bk2 + CHUNK - 1) / CHUNK, lane, r, valid);
        }
        __syncthreads();
        // the scratch of cta_reduce aliases the staging tiles, whose constant rows are rewritten below
        cta_reduce<C>(acc, st, stage0, a.pbuf + ((size_t)par * G + blockIdx.x) * a.npartp, 1);
        bar += (unsigned)G;
        grid_barrier(a.counter, bar);

        // ---- reduce-scatter + optimiser on my slice (fixed order: lane-strided partial sums, xor tree)
        float ntot = 0.f;
        for (int t = 0; t < a.T; t++) ntot += bs[BS_N + t];
        const bool skip = (ntot == 0.f);  // all-masked batch: epoch.jl:17-19
        for (int j = warp; j < a.SL; j += nwarps) {
            const int q = q0 + j;
            float sum = 0.f;
            if (q < C::NPART)
                for (int g = lane; g < G; g += 32) sum += __ldcg(a.pbuf + ((size_t)par * G + g) * a.npartp + q);
            sum = warp_sum(sum);
            if (lane == 0) red[j] = sum;
        }
        __syncthreads();
        for (int j = threadIdx.x; j < a.SL; j += blockDim.x) {
            const int q = q0 + j;
            if (q >= C::NPART) continue;
            float g = red[j];
            if (q >= C::D.npart_dw() && q < C::D.npart_dw() + MAXT) a.stats_out[(size_t)s * MAXT + (q - C::D.npart_dw())] = g;
            const int p = own_p[j];
            if (p < 0) continue;
            float th = own_th[j];
            if (!skip) {
                if (p >= a.ntheta) {
                    float sg = 1.f / (1.f + expf(-th));
                    g *= a.pspan[p] * sg * (1.f - sg);
                }
                float dx;
                if (a.opt_kind == OPT_ADAM || a.opt_kind == OPT_ADAMW) {
                    float mt = a.beta1 * own_m[j] + (1.f - a.beta1) * g;
                    float vt = a.beta2 * own_v[j] + (1.f - a.beta2) * g * g;
                    own_m[j] = mt;
                    own_v[j] = vt;
                    dx = mt / (1.f - b1t) / (sqrtf(vt / (1.f - b2t)) + a.eps) * a.eta;
                    if (a.opt_kind == OPT_ADAMW) dx += (a.adamw_coupled ? a.eta * a.lambda : a.lambda) * th;
                } else if (a.opt_kind == OPT_RMSPROP) {
                    float qv = a.beta2 * own_v[j] + (1.f - a.beta2) * g * g;
                    own_v[j] = qv;
                    dx = g * a.eta / (sqrtf(qv) + a.eps);
                } else {
                    dx = a.eta * g;
                }
                th -= dx;
                own_th[j] = th;
            }
            float* pubp = a.pub + (size_t)par * (a.nflat + PARAM_TAIL);
            __stcg(pubp + p, th);
            if (p >= a.ntheta) {
                const int sl = a.slot_of_flat[p];
                if (sl >= 0) {
                    const PSlot ps = a.slot[sl];
                    float val = ps.lo + ps.span * (1.f / (1.f + expf(-th)));
                    float o4[4];
                    pm_prep_slot(a.pm_id, sl, val, o4);
                    __stcg(pubp + a.nflat + sl, val);
                    for (int i = 0; i < 4; i++) __stcg(pubp + a.nflat + MAXPS + sl * PMS_PER_SLOT + i, o4[i]);
                }
            }
        }
        if (skip) tskip++;
        else { b1t *= a.beta1; b2t *= a.beta2; tdone++; }
        init_stage_rows<C>(stage, lane);  // constant rows were overwritten by the reduction scratch
        bar += (unsigned)G;
        grid_barrier(a.counter, bar);
        load_weights_and_scalars<C>(a.pub + (size_t)par * (a.nflat + PARAM_TAIL), a.nflat, a.wsrc, nullptr, 0, sW, sS);
    }

    // write back the state owned by this CTA
    __syncthreads();
    for (int j = threadIdx.x; j < a.SL; j += blockDim.x) {
        const int p = own_p[j];
        if (p < 0) continue;
        a.pblock[p] = own_th[j];
        a.m[p] = own_m[j];
        a.v[p] = own_v[j];
        if (p >= a.ntheta) {
            const int sl = a.slot_of_flat[p];
            if (sl >= 0) {
                const PSlot ps = a.slot[sl];
                float val = ps.lo + ps.span * (1.f / (1.f + expf(-own_th[j])));
                float o4[4];
                pm_prep_slot(a.pm_id, sl, val, o4);
                a.pblock[a.nflat + sl] = val;
                for (int i = 0; i < 4; i++) a.pblock[a.nflat + MAXPS + sl * PMS_PER_SLOT + i] = o4[i];
            }
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        a.ost->b1t = b1t;
        a.ost->b2t = b2t;
        a.ost->t += tdone;
        a.ost->skipped += tskip;
    }
}

}  // namespace eh
