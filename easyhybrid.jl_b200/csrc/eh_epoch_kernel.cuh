// eh_epoch_kernel.cuh -- the persistent form of the training loop: MANY optimiser steps in one
// launch (one CTA per SM, thread-block clusters), replacing run_epoch!
// (src/training/epoch.jl:13-33) as a whole.  Per step every CTA
//   1. runs the fused forward/backward on its chunks (eh_chunk.cuh) with the weight image it keeps
//      in shared memory,
//   2. reduces its lane tiles to one partial vector in shared memory,
//   3. takes part in a three-hop grid-wide sum (cluster leader over DSMEM -> one vector per cluster
//      in L2 -> every CTA sums a share of them -> shares exchanged over DSMEM); every slot is an
//      8-byte {value, step tag} pair that validates itself, so there is no grid barrier, no fence
//      and no atomic anywhere: readers poll exactly the slots they need,
//   4. applies the optimiser REDUNDANTLY to its own copy of theta / m / v in shared memory,
//      patching its weight image in place.
// All CTAs execute the same float operations in the same order, so the replicas stay bit-identical
// and nothing has to be broadcast back.  Compared with one launch per step this removes two kernel
// launches, the dependent-load prologue and the single-CTA second pass from every step.
#pragma once
#include "eh_step_kernel.cuh"

namespace eh {

struct EpochArgs {
    const float4* rec;
    const int* idx;            // resident index stream (NULL: batch b = records b*B .. of `rec`)
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
    const int* wsrc;           // [NW] image cell -> flat index
    const int* pmap;           // [nflat] flat -> index into the partial vector
    const int* cells;          // [2*nflat] flat -> up to two image cells (-1: none)
    const float* pspan;        // [nflat]
    const int* slot_of_flat;   // [nflat] phi entries: canonical slot (for the tail), -1 otherwise
    const float* bscal;        // [nb][BS_STRIDE]
    uint2* pbuf;               // [2][nclusters][npartp] published cluster vectors, {value bits, step tag} slots
    unsigned tag_base;         // steps run by earlier launches of this ctx (tags are absolute, never reused)
    float* stats_out;          // [nsteps][MAXT] reduced loss sums
    int npartp;                // padded partial length (multiple of 4)
    int work_floats;           // size of the per-CTA work region (staging / scratch / vector landing zone)
    int csize;                 // cluster size (1, 2, 4, 8)
    int T, agg_mean;
    int loss_kind[MAXT];
    PSlot slot[MAXPS];
    float pmc[4];
    int use_bn;
    const PmProgData* prog;    // traced process model (PmProgram variants), device memory
    int scale_rt;              // PmProgram variants: scale_nn_outputs
    int pm_id;
    int opt_kind, adamw_coupled;
    float eta, beta1, beta2, eps, lambda;
    long long* dbg;            // optional [nsteps][gridDim.x][32] SM-clock timestamps (EH_EPOCH_DEBUG)
    // ---- data parallel: one process per GPU, peer memory mapped with CUDA IPC over NVLink ----
    int world, rank;
    unsigned step_base;        // steps exchanged by earlier launches (flags carry absolute step tags)
    uint2* inbox_peer[EH_MAX_WORLD];      // rank r's inbox [2][world][npartp] of {value bits, step tag}, as mapped here
    unsigned* err;             // set to 1 when a bounded spin gives up (peer / CTA never arrived)
};

// 8-byte accesses are single-copy atomic: value and tag always travel together
__device__ __forceinline__ void st_volatile_v2(uint2* p, unsigned x, unsigned y)
{
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ uint2 ld_volatile_v2(const uint2* p)
{
    uint2 v;
    asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    return v;
}
constexpr unsigned EH_SPIN_LIMIT = 1u << 26;  // ~seconds; then give up loudly instead of hanging the GPU

__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// shared-memory window address of `local` in CTA `rank` of this cluster
__device__ __forceinline__ unsigned dsmem_addr(const void* local, unsigned rank)
{
    unsigned a = (unsigned)__cvta_generic_to_shared(local), ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rank));
    return ra;
}
__device__ __forceinline__ void st_dsmem_v2(unsigned addr, unsigned x, unsigned y)
{
    asm volatile("st.relaxed.cluster.shared::cluster.v2.u32 [%0], {%1, %2};" ::"r"(addr), "r"(x), "r"(y) : "memory");
}
__device__ __forceinline__ uint2 ld_shared_v2(const uint2* p)
{
    uint2 v;
    unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("ld.relaxed.cluster.shared::cta.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(a) : "memory");
    return v;
}

// ---- grid-wide sum of the CTA partials: three self-validating hops, no barrier, no fence, no atomic.
// Every slot travels as an 8-byte {value, step tag} pair (single-copy atomic), so a reader simply polls the
// slot it needs until the tag of this step shows up.  Thread p owns element p on every hop.
//   hop 1  non-leader CTAs push their partial into the cluster leader's shared memory (DSMEM store);
//          the leader adds the ranks in order and publishes ONE vector per cluster in L2;
//   hop 2  CTA rank r of every cluster sums its contiguous share of the NC published vectors out of L2;
//   hop 3  the cs shares are pushed to every CTA of the cluster (DSMEM) and added in rank order.
// Every cluster uses the same grouping and order, so all CTAs of the grid hold bit-identical totals.
// Kept out of line: its registers must not weigh on the allocation of the compute phase.
static __device__ __noinline__ void grid_sum(const float* cpart, float* red, uint2* box1, uint2* box3, uint2* pbuf_par, int npartp,
                                      int npart, int cs, int NC, unsigned tag, unsigned* err, long long* dbg)
{
    if (gridDim.x == 1) {   // small batches run in ONE CTA: its partial is the total
        for (int p = threadIdx.x; p < npart; p += blockDim.x) red[p] = cpart[p];
        return;
    }
    const unsigned crank = (unsigned)(blockIdx.x % cs);
    const int cid = blockIdx.x / cs;
    {
        uint2* pub = pbuf_par + (size_t)cid * npartp;
        if (crank != 0) {
            const unsigned dst = dsmem_addr(box1 + (size_t)(crank - 1) * npartp, 0u);
            for (int p = threadIdx.x; p < npart; p += blockDim.x) st_dsmem_v2(dst + 8u * (unsigned)p, __float_as_uint(cpart[p]), tag);
        } else {
            for (int p = threadIdx.x; p < npart; p += blockDim.x) {
                float sum = cpart[p];
                for (int rk = 1; rk < cs; rk++) {
                    const uint2* slot = box1 + (size_t)(rk - 1) * npartp + p;
                    uint2 v = ld_shared_v2(slot);
                    unsigned spins = 0;
                    while (v.y != tag) {
                        if (++spins > EH_SPIN_LIMIT) { *err = 1; break; }
                        v = ld_shared_v2(slot);
                    }
                    sum += __uint_as_float(v.x);
                }
                st_volatile_v2(pub + p, __float_as_uint(sum), tag);
            }
        }
    }
    if (dbg && threadIdx.x == 0) dbg[4] = clock64();
    {
        const uint2* src = pbuf_par;
        const int per = (NC + cs - 1) / cs;
        const int c0 = (int)crank * per;
        const int c1 = c0 + per < NC ? c0 + per : NC;
        constexpr int U = 12;
        for (int p = threadIdx.x; p < npart; p += blockDim.x) {
            float sum = 0.f;
            for (int c = c0; c < c1; c += U) {
                uint2 t[U];
#pragma unroll
                for (int u = 0; u < U; u++)
                    if (c + u < c1) t[u] = ld_volatile_v2(src + (size_t)(c + u) * npartp + p);
#pragma unroll
                for (int u = 0; u < U; u++)
                    if (c + u < c1) {
                        unsigned spins = 0;
                        while (t[u].y != tag) {
                            if (++spins > EH_SPIN_LIMIT) { *err = 1; break; }
                            t[u] = ld_volatile_v2(src + (size_t)(c + u) * npartp + p);
                        }
                        sum += __uint_as_float(t[u].x);
                    }
            }
            if (cs > 1) {
                for (int rk = 0; rk < cs; rk++)
                    st_dsmem_v2(dsmem_addr(box3 + (size_t)crank * npartp + p, (unsigned)rk), __float_as_uint(sum), tag);
            } else {
                red[p] = sum;
            }
        }
    }
    if (dbg && threadIdx.x == 0) dbg[5] = clock64();
    if (cs > 1) {
        for (int p = threadIdx.x; p < npart; p += blockDim.x) {
            float sum = 0.f;
            for (int rk = 0; rk < cs; rk++) {
                const uint2* slot = box3 + (size_t)rk * npartp + p;
                uint2 v = ld_shared_v2(slot);
                unsigned spins = 0;
                while (v.y != tag) {
                    if (++spins > EH_SPIN_LIMIT) { *err = 1; break; }
                    v = ld_shared_v2(slot);
                }
                sum += __uint_as_float(v.x);
            }
            red[p] = sum;
        }
    }
}

// floats of shared memory besides weights / scalars / work region: cpart + red (padded partial vectors)
// + the two {value, tag} inboxes (cs-1 and cs vectors of 8-byte slots) + theta, m, v copies
// + tables (pmap, 2 cells, span, slot)
__host__ __device__ constexpr int epoch_extra_floats(int npartp, int nflat, int cs)
{
    return (2 + 2 * (2 * cs - 1)) * npartp + 8 * rup4(nflat);
}

template <class E>
__global__ void __launch_bounds__(E::MAX_WARPS * 32, 1) k_epoch(const EpochArgs a)
{
    using C = typename E::Cfg;
    extern __shared__ float4 smem4[];
    float* sW = reinterpret_cast<float*>(smem4);
    float* sS = sW + rup4(C::NW);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    float* stage0 = sS + SS_FLOATS;                      // work region: staging tiles / reduction scratch / vectors
    float* stage = stage0 + warp * E::STAGE_FLOATS;
    float* cpart = stage0 + a.work_floats;               // [npartp] this CTA's partial vector
    float* red = cpart + rup4(E::NPART);                 // [npartp] fully reduced vector
    uint2* box1 = reinterpret_cast<uint2*>(red + rup4(E::NPART));  // [cs-1][npartp] {value, tag}: partials pushed to a cluster leader
    uint2* box3 = box1 + (size_t)(a.csize - 1) * rup4(E::NPART);   // [cs][npartp] {value, tag}: the cluster's shares of the grid-wide sum
    float* s_th = reinterpret_cast<float*>(box3 + (size_t)a.csize * rup4(E::NPART));  // [nflat] replicated parameters
    float* s_m = s_th + rup4(a.nflat);
    float* s_v = s_m + rup4(a.nflat);
    int* t_pmap = reinterpret_cast<int*>(s_v + rup4(a.nflat));
    int* t_cell0 = t_pmap + rup4(a.nflat);
    int* t_cell1 = t_cell0 + rup4(a.nflat);
    int* t_slot = t_cell1 + rup4(a.nflat);
    float* t_span = reinterpret_cast<float*>(t_slot + rup4(a.nflat));
    const int G = gridDim.x;
    const int cs = a.csize;
    const int NC = G / cs;                               // clusters = published vectors per step

    for (int p = threadIdx.x; p < a.nflat; p += blockDim.x) {
        s_th[p] = a.pblock[p];
        s_m[p] = a.m[p];
        s_v[p] = a.v[p];
        t_pmap[p] = a.pmap[p];
        t_cell0[p] = a.cells[2 * p];
        t_cell1[p] = a.cells[2 * p + 1];
        t_slot[p] = a.slot_of_flat[p];
        t_span[p] = a.pspan[p];
        if (p >= a.ntheta && p - a.ntheta < MAXPS) {
            // d(squashed phi)/d(raw phi), reused by the gradient of the next step (recomputed after every update)
            const float sg0 = 1.f / (1.f + expf(-a.pblock[p]));
            sS[SS_SGD + p - a.ntheta] = a.pspan[p] * sg0 * (1.f - sg0);
        }
    }
    float b1t = a.ost->b1t, b2t = a.ost->b2t;
    long long tdone = 0, tskip = 0;

    // inboxes start with tag 0 (never a valid step tag); nobody pushes before everybody has cleared
    for (int i = threadIdx.x; i < (2 * cs - 1) * rup4(E::NPART); i += blockDim.x) box1[i] = make_uint2(0u, 0u);
    if (cs > 1) cluster_sync_all(); else __syncthreads();

    typename E::State st;
    E::init_warp(st, stage, lane);
    load_weights_and_scalars<C>(a.pblock, a.nflat, a.wsrc, nullptr, 0, sW, sS);

    PmCtx cx;
    cx.pms = sS + SS_PMS;
    cx.c = a.pmc;
    cx.prog = a.prog;
    cx.scale_rt = a.scale_rt;
    cx.uniform_mask = 0;
#pragma unroll
    for (int s = 0; s < MAXPS; s++)
        if (s >= C::NPS || a.slot[s].role != ROLE_NEURAL) cx.uniform_mask |= 1u << s;

    const int GW = G * nwarps;
    // chunk -> warp assignment interleaves CTAs so that a ragged chunk count spreads over all SMs
    const int gw = warp * G + blockIdx.x;
    // batch of the current step, tracked incrementally (no 64-bit modulo in the loop)
    int bcur = (int)(a.first_step % a.nb);
    const int Blast = (int)(a.n - (long long)(a.nb - 1) * a.B);  // size of the (possibly partial) last batch
    {
        const long long b = bcur;
        const int Bk = bcur == a.nb - 1 ? Blast : a.B;
        E::fetch(st, a.rec, a.idx ? a.idx + b * a.B : nullptr, a.idx ? 0 : b * a.B, Bk, gw, (Bk + E::CHUNK - 1) / E::CHUNK, lane);
    }
    // per-batch scalars of the coming step, prefetched one step ahead (a dependent global load otherwise)
    float pre_bs = 0.f;
    auto prefetch_bs = [&](int batch) {
        const float* bs = a.bscal + (size_t)batch * BS_STRIDE;
        if (threadIdx.x < MAXT) pre_bs = bs[BS_C + threadIdx.x];
        else if (threadIdx.x < MAXT + 2 * C::P)
            pre_bs = a.use_bn ? bs[BS_BN + threadIdx.x - MAXT] : (((threadIdx.x - MAXT) & 1) ? 1.f : 0.f);
        else if (threadIdx.x >= 28 && threadIdx.x < 28 + MAXT) pre_bs = bs[BS_N + threadIdx.x - 28];
    };
    prefetch_bs(bcur);
#define EH_STAMP(slot)                                                                      \
    if (a.dbg && threadIdx.x == 0) a.dbg[((size_t)s * G + blockIdx.x) * 32 + (slot)] = clock64();

    for (int s = 0; s < a.nsteps; s++) {
        EH_STAMP(0)
        const long long b = bcur;
        const int bnext = bcur + 1 == a.nb ? 0 : bcur + 1;
        const int Bk = bcur == a.nb - 1 ? Blast : a.B;
        const int nchunks = (Bk + E::CHUNK - 1) / E::CHUNK;
        const int par = s & 1;
        if (threadIdx.x < MAXT) sS[SS_C + threadIdx.x] = pre_bs;
        else if (threadIdx.x < MAXT + 2 * C::P) sS[SS_BN + threadIdx.x - MAXT] = pre_bs;
        else if (threadIdx.x >= 28 && threadIdx.x < 28 + MAXT) {
            sS[SS_NV + threadIdx.x - 28] = pre_bs;
            // second copy by step parity for the update phase: sS[SS_NV] is rewritten at the top of the next
            // step, which a fast warp may reach while a slow one is still in this step's update
            sS[SS_NV2 + par * MAXT + threadIdx.x - 28] = pre_bs;
        }
        __syncthreads();
        if (s + 1 < a.nsteps) prefetch_bs(bnext);
        EH_STAMP(1)

        E::step_begin(st, sW, lane);
        const FetchArgs fa{a.rec, a.idx ? a.idx + b * a.B : nullptr, a.idx ? 0 : b * a.B, Bk, nchunks};
        // the records this warp needs first in the NEXT step are prefetched while it computes its last chunk of this
        // step (index load + dependent record gather = two DRAM latencies, hidden behind ~12 000 cycles of compute
        // instead of the much shorter exchange)
        const long long b2 = bnext;
        const int Bk2 = bnext == a.nb - 1 ? Blast : a.B;
        const FetchArgs fn{a.rec, a.idx ? a.idx + b2 * a.B : nullptr, a.idx ? 0 : b2 * a.B, Bk2,
                           s + 1 < a.nsteps ? (Bk2 + E::CHUNK - 1) / E::CHUNK : 0};
        for (int chunk = gw; chunk < nchunks; chunk += GW) {
            const bool last = chunk + GW >= nchunks;
            E::chunk(st, last ? fn : fa, last ? gw : chunk + GW, sW, sS, stage, lane, a.slot, a.loss_kind, cx);
        }
        if (a.dbg && lane == 0) a.dbg[((size_t)s * G + blockIdx.x) * 32 + 8 + warp] = clock64();
        // a warp without a chunk in this step still has to fetch its first sample of the next one
        if (gw >= nchunks) E::fetch(st, fn.rec, fn.idx, fn.rec_base, fn.B, gw, fn.nchunks, lane);
        __syncthreads();
        EH_STAMP(2)
        // CTA partial -> cpart (the scratch of cta_reduce aliases the staging tiles)
        E::reduce(st, stage0, cpart, 0);
        __syncthreads();
        EH_STAMP(3)
        // grid-wide sum of the CTA partials (grid_sum above): cpart -> red, identical bits in every CTA
        grid_sum(cpart, red, box1, box3, a.pbuf + (size_t)par * NC * a.npartp, a.npartp, E::NPART, cs, NC, a.tag_base + (unsigned)s + 1u,
                 a.err, a.dbg ? a.dbg + ((size_t)s * G + blockIdx.x) * 32 : nullptr);
        __syncthreads();
        EH_STAMP(6)
        if (a.dbg && threadIdx.x == 0) a.dbg[((size_t)s * G + blockIdx.x) * 32 + 27] = clock64();
        if (a.world > 1) {
            // ---- fused exchange over NVLink peer memory, LL style: every 8-byte store carries {value, step tag},
            // so a slot validates itself -- no fences, no separate flags, one-way NVLink latency.  CTA 0 pushes this
            // GPU's reduced vector into every rank's inbox; every CTA of every GPU polls its own GPU's inbox and sums
            // the rank vectors in rank order (bitwise identical everywhere).
            const unsigned tag = a.step_base + (unsigned)s + 1u;
            if (blockIdx.x == 0) {
                for (int r = 0; r < a.world; r++) {
                    uint2* dst = a.inbox_peer[r] + ((size_t)par * a.world + a.rank) * a.npartp;
                    for (int p = threadIdx.x; p < E::NPART; p += blockDim.x) st_volatile_v2(dst + p, __float_as_uint(red[p]), tag);
                }
                __syncthreads();  // red is about to be overwritten below
            }
            const uint2* inbox = a.inbox_peer[a.rank] + (size_t)par * a.world * a.npartp;
            for (int p = threadIdx.x; p < E::NPART; p += blockDim.x) {
                float sum = 0.f;
                for (int r = 0; r < a.world; r++) {
                    uint2 v = ld_volatile_v2(inbox + (size_t)r * a.npartp + p);
                    unsigned spins = 0;
                    while (v.y != tag) {
                        if (++spins > EH_SPIN_LIMIT) { *a.err = 1; break; }
                        v = ld_volatile_v2(inbox + (size_t)r * a.npartp + p);
                    }
                    sum += __uint_as_float(v.x);
                }
                red[p] = sum;
            }
            __syncthreads();
        }
        if (a.dbg && threadIdx.x == 0) a.dbg[((size_t)s * G + blockIdx.x) * 32 + 28] = clock64();
        // every thread derives the batch scalars itself (identical arithmetic everywhere): no serial section
        float ntot = 0.f, post = 1.f;
        for (int t = 0; t < a.T; t++) {
            const float nv = sS[SS_NV2 + par * MAXT + t];
            ntot += nv;
            if (a.loss_kind[t] == LOSS_RMSE) post = 1.f / (2.f * sqrtf(red[E::OFF_STATS + t] / nv));
        }
        const bool skip = ntot == 0.f;  // all-masked batch: epoch.jl:17-19
        if (blockIdx.x == 0 && threadIdx.x < MAXT) a.stats_out[(size_t)s * MAXT + threadIdx.x] = red[E::OFF_STATS + threadIdx.x];
        EH_STAMP(29)
        if (!skip) {
            for (int p = threadIdx.x; p < a.nflat; p += blockDim.x) {
                float g = red[t_pmap[p]] * post;
                float th = s_th[p];
                const bool phi_cached = p >= a.ntheta && p - a.ntheta < MAXPS;
                if (phi_cached) {
                    g *= sS[SS_SGD + p - a.ntheta];
                } else if (p >= a.ntheta) {
                    float sg = 1.f / (1.f + expf(-th));
                    g *= t_span[p] * sg * (1.f - sg);
                }
                float dx;
                if (a.opt_kind == OPT_ADAM || a.opt_kind == OPT_ADAMW) {
                    float mt = a.beta1 * s_m[p] + (1.f - a.beta1) * g;
                    float vt = a.beta2 * s_v[p] + (1.f - a.beta2) * g * g;
                    s_m[p] = mt;
                    s_v[p] = vt;
                    dx = mt / (1.f - b1t) / (sqrtf(vt / (1.f - b2t)) + a.eps) * a.eta;
                    if (a.opt_kind == OPT_ADAMW) dx += (a.adamw_coupled ? a.eta * a.lambda : a.lambda) * th;
                } else if (a.opt_kind == OPT_RMSPROP) {
                    float qv = a.beta2 * s_v[p] + (1.f - a.beta2) * g * g;
                    s_v[p] = qv;
                    dx = g * a.eta / (sqrtf(qv) + a.eps);
                } else {
                    dx = a.eta * g;
                }
                th -= dx;
                s_th[p] = th;
                // patch my weight image in place
                const int c0 = t_cell0[p], c1 = t_cell1[p];
                if (c0 >= 0) sW[c0] = th;
                if (c1 >= 0) sW[c1] = th;
                if (p >= a.ntheta) {
                    const float sgn = 1.f / (1.f + expf(-th));
                    if (phi_cached) sS[SS_SGD + p - a.ntheta] = t_span[p] * sgn * (1.f - sgn);
                    const int sl = t_slot[p];
                    if (sl >= 0) {
                        const PSlot ps = a.slot[sl];
                        float val = ps.lo + ps.span * sgn;
                        float o4[4];
                        pm_prep_slot_fast(a.pm_id, sl, val, o4);
                        sS[SS_SLOT + sl] = val;
                        for (int i = 0; i < 4; i++) sS[SS_PMS + sl * PMS_PER_SLOT + i] = o4[i];
                    }
                }
            }
            EH_STAMP(30)
            b1t *= a.beta1;
            b2t *= a.beta2;
            tdone++;
        } else {
            tskip++;
        }
        E::after_reduce(st, stage, lane);  // constant staging rows were overwritten by the scratch / vectors
        bcur = bnext;
        EH_STAMP(7)
    }
#undef EH_STAMP

    // write back (CTA 0 holds the same state as everybody else)
    __syncthreads();
    if (blockIdx.x == 0) {
        for (int p = threadIdx.x; p < a.nflat; p += blockDim.x) {
            a.pblock[p] = s_th[p];
            a.m[p] = s_m[p];
            a.v[p] = s_v[p];
        }
        if (threadIdx.x < MAXPS) a.pblock[a.nflat + threadIdx.x] = sS[SS_SLOT + threadIdx.x];
        if (threadIdx.x < MAXPS * PMS_PER_SLOT) a.pblock[a.nflat + MAXPS + threadIdx.x] = sS[SS_PMS + threadIdx.x];
        if (threadIdx.x == 0) {
            a.ost->b1t = b1t;
            a.ost->b2t = b2t;
            a.ost->t += tdone;
            a.ost->skipped += tskip;
        }
    }
    if (cs > 1) cluster_sync_all();  // nobody leaves while a leader may still read its shared memory
}

}  // namespace eh
