// eh_epoch_kernel.cuh -- the persistent form of the training loop: MANY optimiser steps in one
// launch (one CTA per SM, cooperative launch), replacing run_epoch! (src/training/epoch.jl:13-33) as a whole.
//
// CTA layout: `wcomp` compute warps + ONE service warp (the last one).  Per step
//   1. the service warp prefetches the CTA's records of the NEXT step with one cp.async.bulk (TMA) into a
//      double-buffered shared-memory tile (batch-contiguous records: the staged epoch stream or host-batch slots;
//      collect_dim_data, epoch.jl:1-11); with an index stream (gather mode) lanes prefetch their records themselves;
//   2. the compute warps run the fused forward/backward on the CTA's contiguous chunk range (eh_chunk.cuh) with the
//      weight image the CTA keeps in shared memory;
//   3. grid-wide sum of the CTA partials as a reduce-scatter + all-gather through L2, no barrier, no fence, no atomic:
//      every slot is an 8-byte {value, step tag} pair that validates itself, readers poll exactly the slots they need.
//        A  every CTA publishes its partial vector (476 slots for [2-16-16-1]);
//        B  the vector is cut into slices of 4 elements, slice j belongs to CTA j mod G: the owner sums the slice over
//           all CTAs in a fixed order (lane groups + shuffle tree) -- with several GPUs it also swaps the slice sums with
//           its peers over NVLink here (rank-ordered sum) -- and publishes the totals;
//        C  every thread that owns a parameter polls the one total it needs;
//      each total is computed by exactly ONE CTA, so all CTAs (and all GPUs: same rank order) see identical bits;
//   4. every CTA applies the optimiser REDUNDANTLY to its own copy of theta / m / v in shared memory and patches its
//      weight image in place: nothing is broadcast back.  The global physical parameters (phi) live on the service
//      warp: their long tail (sigmoid, squashing, double-precision log2) runs while the compute warps have already
//      started the next step; consumers wait on a shared-memory flag right before they need those scalars.
// Two CTA barriers per step (after the compute phase, after the optimiser).
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
    uint2* pbuf;               // [2][pbuf_rows][npartp] {value bits, step tag}: rows 0..G-1 CTA partials, the last EH_TOT_REPL rows totals
    int pbuf_rows;
    unsigned tag_base;         // steps run by earlier launches of this ctx (tags are absolute, never reused)
    float* stats_out;          // [nsteps][MAXT] reduced loss sums
    int npartp;                // padded partial length (multiple of 4)
    int work_floats;           // size of the per-CTA work region (staging tiles / reduction rows)
    int wcomp;                 // compute warps (blockDim.x / 32 - 1)
    int pg_log2;               // phase B: lanes per element = 1 << pg_log2
    int tile_floats;           // floats per record-tile buffer (0: no TMA staging)
    int T, agg_mean;
    int loss_kind[MAXT];
    PSlot slot[MAXPS];
    float pmc[4];
    int use_bn;
    const PmProgData* prog;    // traced process model (PmProgram variants), device memory
    int scale_rt;              // PmProgram variants: scale_nn_outputs
    unsigned pass_mask[3];     // PmProgram variants: pass-through units per hidden layer (PmCtx::pass)
    int pm_id;
    int opt_kind, adamw_coupled;
    float eta, beta1, beta2, eps, lambda;
    long long* dbg;            // optional [nsteps][gridDim.x][32] SM-clock timestamps (EH_EPOCH_DEBUG)
    // ---- consumer mode (host-batch stream, eh_step_host_async): the kernel is launched with the first batch of a burst
    // and trains on ring slot (first_step + s) mod nb as soon as the packer kernel has published that slot; it leaves when
    // the host has announced the number of steps of the burst and they are done.  nsteps is then only an upper bound.
    const unsigned* ready;     // [nb] per-slot tags (NULL: not a stream); slot of step s carries ready_base + s + 1
    unsigned ready_base;
    unsigned* done;            // steps retired (absolute count = ready_base + s + 1), written by CTA 0; packers wait on it
    const int* host_total;     // host-mapped: number of steps of this burst once the host knows it (0: still open)
    float* loss_stream;        // [nsteps] host-mapped loss cells, written as the steps retire
    long long batch_stride;    // records between consecutive batches of `rec` (0: B)
    // ---- data parallel: one process per GPU, peer memory mapped with CUDA IPC over NVLink ----
    int world, rank;
    unsigned step_base;        // steps exchanged by earlier launches (flags carry absolute step tags)
    uint2* inbox_peer[EH_MAX_WORLD];      // rank r's inbox [2][world][npartp] of {value bits, step tag}, as mapped here
    unsigned* err;             // set to 1 when a bounded spin gives up (peer / CTA never arrived)
    int stagger_ns;            // tile engines: tile slot u of a CTA starts its step u * stagger_ns later (see the chunk loop)
    int eng_off;               // byte offset of the engine's own shared-memory region (tensor engine: tf32 weight images,
                               // mbarriers, TMEM slot); 0: none
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
#ifndef EH_TOT_REPL_N
#define EH_TOT_REPL_N 8
#endif
constexpr int EH_TOT_REPL = EH_TOT_REPL_N;   // replicas of the published totals (the last rows of EpochArgs::pbuf; at most 32)
constexpr int EH_PBUF_EXTRA_ROWS = 32;       // rows of pbuf behind the CTA partials
constexpr unsigned EH_SPIN_LIMIT = 1u << 26;  // ~seconds; then give up loudly instead of hanging the GPU

// Slot pairs that live in this GPU's L2 use WEAK cache-global accesses (measured, tools/ubench_exchange.cu: volatile /
// relaxed accesses are handled one transaction at a time per SM -- 9.8 us per exchange of 148 CTAs against 5.1 us with
// ld/st.global.cg, 3.0 us when every thread issues exactly one load).  L2 is the point of coherence, the asm is volatile
// (never hoisted or merged), and every slot validates itself through its tag.
__device__ __forceinline__ void st_cg_v4(uint2* p, unsigned x0, unsigned t0, unsigned x1, unsigned t1)
{
    asm volatile("st.global.cg.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(x0), "r"(t0), "r"(x1), "r"(t1) : "memory");
}
__device__ __forceinline__ uint4 ld_cg_v4(const uint2* p)
{
    uint4 v;
    asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
// poll a pair of slots (weak loads) until both carry the tag of this step (a __nanosleep back-off of 20 / 100 ns between
// attempts was measured: 10.75 -> 10.77 / 10.81 us per step, i.e. the retries are not what loads L2)
__device__ __forceinline__ uint4 poll_pair_cg(const uint2* p, unsigned tag, unsigned* err)
{
    uint4 v = ld_cg_v4(p);
    unsigned spins = 0;
    while (v.y != tag || v.w != tag) {
        if (++spins > EH_SPIN_LIMIT) { *err = 1; break; }
        v = ld_cg_v4(p);
    }
    return v;
}
// two neighbouring slots at once (16-byte aligned): {value0, tag0, value1, tag1}
__device__ __forceinline__ void st_volatile_v4(uint2* p, unsigned x0, unsigned t0, unsigned x1, unsigned t1)
{
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(x0), "r"(t0), "r"(x1), "r"(t1) : "memory");
}
__device__ __forceinline__ uint4 ld_volatile_v4(const uint2* p)
{
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}

// poll a pair of slots until both carry the tag of this step
__device__ __forceinline__ uint4 poll_pair(const uint2* p, uint4 v, unsigned tag, unsigned* err)
{
    unsigned spins = 0;
    while (v.y != tag || v.w != tag) {
        if (++spins > EH_SPIN_LIMIT) { *err = 1; break; }
        v = ld_volatile_v4(p);
    }
    return v;
}

// poll a {value, tag} slot until the tag of this step shows up
__device__ __forceinline__ float poll_slot(const uint2* p, unsigned tag, unsigned* err)
{
    uint2 v = ld_volatile_v2(p);
    unsigned spins = 0;
    while (v.y != tag) {
        if (++spins > EH_SPIN_LIMIT) { *err = 1; break; }
        v = ld_volatile_v2(p);
    }
    return __uint_as_float(v.x);
}

// named barriers: the compute warps wait (sync), the service warp only announces itself (arrive)
__device__ __forceinline__ void bar_sync_id(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }
__device__ __forceinline__ void bar_arrive_id(int id, int nthreads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

// ---- mbarrier + bulk copy (TMA) for the record tiles ----
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned parity)
{
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// bounded: a bulk copy that never lands raises the error flag instead of hanging the GPU
// (consumer mode waits for batches the host has not even submitted yet: no bound there)
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity, unsigned* err, bool unbounded = false)
{
    unsigned spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (!unbounded && ++spins > (1u << 22)) { *err = 1; break; }
    }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

// contiguous chunk range of CTA `bid`: chunks [c0, c0 + cnt) of a batch with nchunks chunks
__device__ __forceinline__ void cta_chunk_range(int nchunks, int G, int bid, int& c0, int& cnt)
{
    const int base = nchunks / G, rem = nchunks - base * G;
    cnt = base + (bid < rem ? 1 : 0);
    c0 = bid * base + (bid < rem ? bid : rem);
}

// floats of shared memory besides weights / scalars / work region: red (padded partial vector; single-CTA grids only),
// theta, m, v copies + tables (pmap, 2 cells, span, slot), two record tiles, two mbarriers + the phi flag
// + xbuf: the values an owner CTA collects for its slices, [row][half][G] float2 (at most NSL + G rows x peers)
__host__ __device__ constexpr int epoch_xbuf_floats(int npartp, int G) { return 4 * (npartp / 4 + G) + 8; }
__host__ __device__ constexpr int epoch_extra_floats(int npartp, int nflat, int tile_floats, int G)
{
    return npartp + 8 * rup4(nflat) + 2 * rup4(tile_floats) + 8 + epoch_xbuf_floats(npartp, G);
}

// warps that share one chunk (tile): 1 for the lane-per-sample engines, 4 for the tensor engine
template <class E, class = void>
struct EngWpc { static constexpr int value = 1; };
template <class E>
struct EngWpc<E, decltype((void)E::WPC)> { static constexpr int value = E::WPC; };
// threads of the persistent CTA: the engine's compute warps + the service warp; 512 (128 registers each) unless the
// engine needs 16 compute warps (tensor engine: 544 threads of at most 120 registers)
template <class E>
constexpr int epoch_threads() { return EngWpc<E>::value > 1 ? (E::MAX_WARPS + 1) * 32 : ((E::MAX_WARPS + 1) * 32 > 512 ? 512 : (E::MAX_WARPS + 1) * 32); }

template <class E>
__global__ void __launch_bounds__(epoch_threads<E>(), 1) k_epoch(const EpochArgs a)
{
    constexpr int WPC = EngWpc<E>::value;
    using C = typename E::Cfg;
    extern __shared__ float4 smem4[];
    float* sW = reinterpret_cast<float*>(smem4);
    float* sS = sW + rup4(C::NW);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int wcomp = a.wcomp;
    const bool service = warp == nwarps - 1;
    float* stage0 = sS + SS_FLOATS;                      // work region: staging tiles / reduction rows
    float* stage = stage0 + (service ? 0 : warp) * E::STAGE_FLOATS;
    float* red = stage0 + a.work_floats;                 // [npartp] fully reduced vector (gridDim.x == 1 only)
    float* s_th = red + a.npartp;                        // [nflat] replicated parameters
    float* s_m = s_th + rup4(a.nflat);
    float* s_v = s_m + rup4(a.nflat);
    int* t_pmap = reinterpret_cast<int*>(s_v + rup4(a.nflat));
    int* t_cell0 = t_pmap + rup4(a.nflat);
    int* t_cell1 = t_cell0 + rup4(a.nflat);
    int* t_slot = t_cell1 + rup4(a.nflat);
    float* t_span = reinterpret_cast<float*>(t_slot + rup4(a.nflat));
    float4* tile0 = reinterpret_cast<float4*>(t_span + rup4(a.nflat));   // [2][tile_floats]
    unsigned long long* mbar = reinterpret_cast<unsigned long long*>(reinterpret_cast<float*>(tile0) + 2 * rup4(a.tile_floats));  // [2]
    unsigned* phi_flag = reinterpret_cast<unsigned*>(mbar + 2);
    float2* xbuf = reinterpret_cast<float2*>(phi_flag + 4);   // [nown][2][G] values collected by this CTA as a slice owner (sized for G <= number of SMs)
    const int G = gridDim.x, bid = blockIdx.x;
    unsigned char* eng = reinterpret_cast<unsigned char*>(smem4) + a.eng_off;
    const int unit = warp / WPC, nunits = wcomp / WPC;   // this warp's tile slot, tile slots of the CTA
    if constexpr (WPC > 1) E::cta_init(eng);
    const bool tiles = a.tile_floats > 0;
    constexpr int R44 = C::R4 / 4;

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
    if (threadIdx.x == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        *phi_flag = a.tag_base;   // the scalars loaded below belong to "step tag_base"
        phi_flag[1] = 0u;         // stop flag of the consumer mode
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }

    typename E::State st;
    if constexpr (WPC == 1) {
        if (!service) E::init_warp(st, stage, lane);
    }
    load_weights_and_scalars<C>(a.pblock, a.nflat, a.wsrc, nullptr, 0, sW, sS);
    __syncthreads();   // the batch-scalar cells are rewritten by other threads below (store_bs)
    if constexpr (WPC > 1) {
        E::load_w2(eng, sW);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the tensor core reads these images
        if (!service) E::init_warp(st, stage, lane, eng);
    }

    PmCtx cx;
    cx.pms = sS + SS_PMS;
    cx.c = a.pmc;
    cx.prog = a.prog;
    cx.scale_rt = a.scale_rt;
    for (int l = 0; l < 3; l++) cx.pass[l] = a.pass_mask[l];
    cx.uniform_mask = 0;
    cx.phi_flag = phi_flag;
    cx.phi_want = a.tag_base;
#pragma unroll
    for (int s = 0; s < MAXPS; s++)
        if (s >= C::NPS || a.slot[s].role != ROLE_NEURAL) cx.uniform_mask |= 1u << s;

    // batch of the current step, tracked incrementally (no 64-bit modulo in the loop)
    int bcur = (int)(a.first_step % a.nb);
    const int Blast = (int)(a.n - (long long)(a.nb - 1) * a.B);  // size of the (possibly partial) last batch
    auto batch_size = [&](int b) { return b == a.nb - 1 ? Blast : a.B; };
    const long long bstride = a.batch_stride ? a.batch_stride : (long long)a.B;
    const bool stream = a.ready != nullptr;
    // the service warp's lane 0 stages the CTA's records of batch b into tile buffer `buf` (one bulk copy; an empty
    // range still completes the mbarrier phase so that the parities keep counting steps)
    auto issue_tile = [&](int b, int buf) {
        const int Bk = batch_size(b);
        int c0, cnt;
        cta_chunk_range((Bk + E::CHUNK - 1) / E::CHUNK, G, bid, c0, cnt);
        const long long s0 = (long long)c0 * E::CHUNK;
        long long s1 = s0 + (long long)cnt * E::CHUNK;
        if (s1 > Bk) s1 = Bk;
        if (s1 > s0) {
            const unsigned bytes = (unsigned)(s1 - s0) * (unsigned)(C::R4 * 4);
            mbar_expect_tx(&mbar[buf], bytes);
            bulk_g2s(tile0 + (size_t)buf * (rup4(a.tile_floats) / 4), a.rec + ((long long)b * bstride + s0) * R44, bytes, &mbar[buf]);
        } else {
            mbar_arrive(&mbar[buf]);
        }
    };
    auto fetch_args = [&](int b, int buf) {
        FetchArgs fa;
        const int Bk = batch_size(b);
        fa.rec = a.rec;
        fa.idx = a.idx ? a.idx + (long long)b * a.B : nullptr;
        fa.rec_base = a.idx ? 0 : (long long)b * bstride;
        fa.B = Bk;
        fa.nchunks = (Bk + E::CHUNK - 1) / E::CHUNK;
        int c0, cnt;
        cta_chunk_range(fa.nchunks, G, bid, c0, cnt);
        fa.tile = tiles ? tile0 + (size_t)buf * (rup4(a.tile_floats) / 4) : nullptr;
        fa.tile_s0 = c0 * E::CHUNK;
        return fa;
    };
    // Per-batch scalars (seed scales c_t, BatchNorm rows, valid counts): 32 values, one per lane of the SERVICE warp, which
    // fetches them one step ahead and puts them into shared memory between the compute phases.
    auto load_bs = [&](int batch) {
        const float* bs = a.bscal + (size_t)batch * BS_STRIDE;
        float v = 0.f;
        if (lane < MAXT) v = bs[BS_C + lane];
        else if (lane < MAXT + 2 * C::P) v = a.use_bn ? bs[BS_BN + lane - MAXT] : (((lane - MAXT) & 1) ? 1.f : 0.f);
        else if (lane >= 28 && lane < 28 + MAXT) v = bs[BS_N + lane - 28];
        return v;
    };
    auto store_bs = [&](float v, int s) {   // scalars of step s (valid counts double-buffered: the optimiser of step s - 1 may still read its own)
        if (lane < MAXT) sS[SS_C + lane] = v;
        else if (lane < MAXT + 2 * C::P) sS[SS_BN + lane - MAXT] = v;
        else if (lane >= 28 && lane < 28 + MAXT) sS[SS_NV2 + (s & 1) * MAXT + lane - 28] = v;
    };
    // consumer mode: has the packer published the slot of step s?  (all lanes read the same word)
    auto slot_ready = [&](int s, int batch) { return *reinterpret_cast<const volatile unsigned*>(a.ready + batch) == a.ready_base + (unsigned)s + 1u; };
    // ... wait for it; false: the host has closed the burst before step s
    auto wait_slot = [&](int s, int batch) {
        for (;;) {
            if (slot_ready(s, batch)) return true;
            const int tot = *reinterpret_cast<const volatile int*>(a.host_total);
            if (tot > 0 && s >= tot) return false;
        }
    };
    volatile unsigned* stop_flag = phi_flag + 1;   // set by the service warp when the burst ends before a.nsteps (cleared above)
    float pre = 0.f;
    bool have_next = false;
    if (service) {
        bool go = true;
        if (stream) go = wait_slot(0, bcur);
        if (go) {
            store_bs(load_bs(bcur), 0);
        } else if (lane == 0) {
            *stop_flag = 1u;
        }
    }
    __syncthreads();   // mbarriers initialised, weights / scalars / tables in place
    if (tiles && service && lane == 0 && !*stop_flag) issue_tile(bcur, 0);
    if (!tiles && !service) {
        const FetchArgs f0 = fetch_args(bcur, 0);
        int c0, cnt;
        cta_chunk_range(f0.nchunks, G, bid, c0, cnt);
        E::fetch(st, f0, unit < cnt ? c0 + unit : f0.nchunks, lane);
    }
#define EH_STAMP(slot)                                                                      \
    if (a.dbg && threadIdx.x == 0) a.dbg[((size_t)s * G + bid) * 32 + (slot)] = clock64();
#define EH_STAMP_SVC(slot)                                                                  \
    if (a.dbg && service && lane == 0) a.dbg[((size_t)s * G + bid) * 32 + (slot)] = clock64();

    const int NSL = a.npartp / 4;                        // slices of the (padded) partial vector
    const int nown = (NSL - bid + G - 1) / G;            // slices owned by this CTA (bid, bid + G, ...)
    const int ncomp_threads = wcomp * 32;
    const int Gp = (G + 3) & ~3;

    int s_end = a.nsteps;
    for (int s = 0; s < a.nsteps; s++) {
        if (*stop_flag) { s_end = s; break; }
        EH_STAMP(0)
        const int bnext = bcur + 1 == a.nb ? 0 : bcur + 1;
        const int par = s & 1;
        const unsigned tag = a.tag_base + (unsigned)s + 1u;
        uint2* part = a.pbuf + (size_t)par * a.pbuf_rows * a.npartp;
        // the totals are published EH_TOT_REPL times (rows pbuf_rows - R ..): every CTA polls replica bid % R, so that at most
        // G / R CTAs poll the same L2 lines
        uint2* tot = part + (size_t)(a.pbuf_rows - EH_TOT_REPL) * a.npartp;
        const uint2* mytot = tot + (size_t)(bid % EH_TOT_REPL) * a.npartp;

        // ---- compute phase ----
        if (service) {
            // the next step's scalars and record tile (its buffer was last read in the compute phase of step s - 1); in
            // consumer mode only if its batch has already landed -- otherwise after this step (below)
            have_next = s + 1 < a.nsteps && (!stream || slot_ready(s + 1, bnext));
            if (have_next) {
                pre = load_bs(bnext);
                if (tiles && lane == 0) issue_tile(bnext, par ^ 1);
            }
        } else {
            const FetchArgs fa = fetch_args(bcur, par);
            const FetchArgs fn = fetch_args(bnext, par ^ 1);
            int c0, cnt, n0, ncnt;
            cta_chunk_range(fa.nchunks, G, bid, c0, cnt);
            cta_chunk_range(fn.nchunks, G, bid, n0, ncnt);
            // first chunk of this warp in the next step (gather mode prefetches it during the last chunk of this step:
            // index load + dependent record gather = two DRAM latencies, hidden behind the compute)
            const int nfirst = (s + 1 < a.nsteps && unit < ncnt) ? n0 + unit : fn.nchunks;
            E::step_begin(st, sW, lane);
            if (tiles) {
                mbar_wait(&mbar[par], (unsigned)(s >> 1) & 1u, a.err, stream);
                if (*stop_flag) { s_end = s; break; }   // consumer mode: the burst ended, the barrier was released without a tile
                E::fetch(st, fa, unit < cnt ? c0 + unit : fa.nchunks, lane);
            }
            cx.phi_want = a.tag_base + (unsigned)s;
            if constexpr (WPC > 1) {
                // The tile slots of a CTA would otherwise run in lockstep: all 16 warps in the MUFU-bound activation phases
                // together, then all in the tensor-core round trips, then all in the HMMA / LDS-bound weight-gradient phase --
                // every pipe saturated in its phase and idle in the others.  Starting slot u a little later than slot u - 1
                // lets the phases of different tiles overlap (measured: DESIGN.md section 5.3).
                if (a.stagger_ns > 0 && unit > 0) __nanosleep((unsigned)(unit * a.stagger_ns));
            }
            for (int chunk = c0 + unit; chunk < c0 + cnt; chunk += nunits) {
                const bool last = chunk + nunits >= c0 + cnt;
                if constexpr (WPC > 1) {
                    if (last && !tiles) E::chunk(st, fn, nfirst, sW, sS, stage, lane, a.slot, a.loss_kind, cx, a.err);
                    else E::chunk(st, fa, last ? fa.nchunks : chunk + nunits, sW, sS, stage, lane, a.slot, a.loss_kind, cx, a.err);
                } else {
                    if (last && !tiles) E::chunk(st, fn, nfirst, sW, sS, stage, lane, a.slot, a.loss_kind, cx);
                    else E::chunk(st, fa, last ? fa.nchunks : chunk + nunits, sW, sS, stage, lane, a.slot, a.loss_kind, cx);
                }
            }
            if (a.dbg && lane == 0) a.dbg[((size_t)s * G + bid) * 32 + 8 + warp] = clock64();
            // a warp without a chunk in this step still has to fetch its first sample of the next one
            if (!tiles && unit >= cnt) E::fetch(st, fn, nfirst, lane);
            if constexpr (WPC > 1) E::reduce_prepare(st, stage0, a.err);
            else if (!E::SCRATCH_ALIASES_STAGE) E::reduce_prepare(st, stage0);
        }
        EH_STAMP(1)
        if constexpr (WPC == 1) {
            if (E::SCRATCH_ALIASES_STAGE) {
                __syncthreads();   // every warp is done with its staging tile; the work region becomes the reduction rows
                if (!service) E::reduce_prepare(st, stage0);
            }
        }
        __syncthreads();
        EH_STAMP(2)
        if (service && have_next) store_bs(pre, s + 1);   // the compute phase that read the cells of step s has ended

        // ---- A: CTA partial: ONE element per thread (every row fetched at once), neighbours paired with a shuffle and published
        // as 16-byte {value, tag, value, tag} pairs.  Slice-major layout [slice][CTA][4 slots]: the owner of a slice finds the
        // contributions of all CTAs in ONE contiguous run of 32 * Gp bytes, Gp = G rounded up to a multiple of 4 (whole
        // 128-byte lines; coalesced loads in phase B) ----
        for (int p0 = 0; p0 < a.npartp; p0 += blockDim.x) {
            const int p = p0 + threadIdx.x;
            const float v0 = p < E::NPART ? E::reduce_sum_at(stage0, wcomp, p) : 0.f;   // (padding slots carry zeros)
            const float v1 = __shfl_down_sync(0xffffffffu, v0, 1);
            if (p < a.npartp && !(p & 1)) {
                const int k = p >> 1;
                if (G == 1) *reinterpret_cast<float2*>(red + p) = make_float2(v0, v1);
                else st_cg_v4(part + ((size_t)(k >> 1) * Gp + bid) * 4 + (k & 1) * 2, __float_as_uint(v0), tag, __float_as_uint(v1), tag);
            }
        }
        EH_STAMP(3)
        // bias corrections are step constants: their reciprocals are taken here, before the sums arrive (IEEE division and
        // square root throughout: MUFU-based 2-ulp forms were measured, 1.72k -> 1.55k cycles in the optimiser phase, no
        // difference in the step time beyond run-to-run noise)
        const float rb1 = 1.f / (1.f - b1t), rb2 = 1.f / (1.f - b2t);
        if (G > 1) {
            // ---- B: slice owners.  Slice j (4 elements) belongs to CTA j mod G.  Every thread fetches ONE 16-byte pair
            // {2 elements of one slice of one peer CTA}; consecutive threads read consecutive pairs of the slice's run:
            // item i -> (row, peer, half) = (i / 2G, (i % 2G) / 2, i & 1); xbuf keeps the [row][half][peer] order of the sums.
            const int nitems = 2 * G * nown;
            for (int i = threadIdx.x; i < nitems; i += blockDim.x) {
                const int row = i / (2 * G), rem = i - row * 2 * G, half = rem & 1, peer = rem >> 1;
                const uint4 v = poll_pair_cg(part + (size_t)(bid + row * G) * Gp * 4 + rem * 2, tag, a.err);
                xbuf[(row * 2 + half) * G + peer] = make_float2(__uint_as_float(v.x), __uint_as_float(v.z));
            }
            __syncthreads();
            // one warp per (row, half): lanes stride over the peers, butterfly, lane 0 publishes the two totals
            {
                const unsigned dtag = a.step_base + (unsigned)s + 1u;
                const int dpar = (int)((a.step_base + (unsigned)s) & 1u);
                for (int pr = nwarps - 1 - warp; pr < 2 * nown; pr += nwarps) {
                    EH_STAMP_SVC(24)
                    const float2* src = xbuf + (size_t)pr * G;
                    float acc0 = 0.f, acc1 = 0.f;
                    for (int c = lane; c < G; c += 32) { const float2 v = src[c]; acc0 += v.x; acc1 += v.y; }
                    for (int o = 16; o > 0; o >>= 1) {
                        acc0 += __shfl_xor_sync(0xffffffffu, acc0, o);
                        acc1 += __shfl_xor_sync(0xffffffffu, acc1, o);
                    }
                    const int ge = (bid + (pr >> 1) * G) * 4 + (pr & 1) * 2;
                    if (a.world > 1) {
                        // ---- fused exchange over NVLink peer memory, LL style: every slot carries {value, step tag}.
                        // Lane r pushes this GPU's sums into rank r's inbox and polls the slot rank r wrote here; the
                        // ranks' values are added in RANK ORDER (identical bits on every GPU whatever its grid looks like).
                        float v0 = 0.f, v1 = 0.f;
                        if (lane < a.world) {
                            st_volatile_v4(a.inbox_peer[lane] + ((size_t)dpar * a.world + a.rank) * a.npartp + ge, __float_as_uint(acc0), dtag,
                                           __float_as_uint(acc1), dtag);
                            const uint2* in = a.inbox_peer[a.rank] + ((size_t)dpar * a.world + lane) * a.npartp + ge;
                            const uint4 v = poll_pair(in, ld_volatile_v4(in), dtag, a.err);
                            v0 = __uint_as_float(v.x);
                            v1 = __uint_as_float(v.z);
                        }
                        acc0 = 0.f; acc1 = 0.f;
                        for (int r = 0; r < a.world; r++) {
                            acc0 += __shfl_sync(0xffffffffu, v0, r);
                            acc1 += __shfl_sync(0xffffffffu, v1, r);
                        }
                    }
                    if (lane < EH_TOT_REPL) st_cg_v4(tot + (size_t)lane * a.npartp + ge, __float_as_uint(acc0), tag, __float_as_uint(acc1), tag);
                    EH_STAMP_SVC(25)
                }
            }
            // ---- C: every thread fetches ONE pair of totals into shared memory ----
            for (int k = threadIdx.x; k < a.npartp / 2; k += blockDim.x) {
                const uint4 v = poll_pair_cg(mytot + 2 * k, tag, a.err);
                *reinterpret_cast<float2*>(red + 2 * k) = make_float2(__uint_as_float(v.x), __uint_as_float(v.z));
            }
        }
        __syncthreads();   // red holds the grid-wide sums
        // the service warp owns no weight-image cell: it announces itself at the end-of-step barrier right away, the
        // compute warps do not wait for its tail (consumers of the global parameters' scalars poll a flag instead)
        if (service) bar_arrive_id(1, blockDim.x);
        EH_STAMP(4)

        // ---- optimiser: every thread owns parameters; red holds the grid-wide sums ----
        // every thread derives the batch scalars itself (identical arithmetic everywhere): no serial section
        float ntot = 0.f, post = 1.f;
        for (int t = 0; t < a.T; t++) {
            const float nv = sS[SS_NV2 + par * MAXT + t];
            ntot += nv;
            if (a.loss_kind[t] == LOSS_RMSE) post = 1.f / (2.f * sqrtf(red[E::OFF_STATS + t] / nv));
        }
        const bool skip = ntot == 0.f;  // all-masked batch: epoch.jl:17-19
        if (bid == 0 && service) {
            if (lane < MAXT) a.stats_out[(size_t)s * MAXT + lane] = red[E::OFF_STATS + lane];
            if (stream && lane == 0) {
                // consumer mode: the loss of this step straight into the host's page-locked cell (loss_fn.jl:58-66; agg
                // over the targets, compute_loss.jl:50-53), and the retired-step count the packers wait on
                float L = 0.f;
                for (int t = 0; t < a.T; t++) {
                    const float nv = sS[SS_NV2 + par * MAXT + t], acc = red[E::OFF_STATS + t];
                    L += a.loss_kind[t] == LOSS_RMSE ? sqrtf(acc / nv) : acc / nv;
                }
                if (a.agg_mean) L /= (float)a.T;
                a.loss_stream[s] = skip ? __int_as_float(0x7fc00000) : L;
                *reinterpret_cast<volatile unsigned*>(a.done) = a.ready_base + (unsigned)s + 1u;
            }
        }
        // compute warps own theta (entry p on thread p, ...), the service warp owns phi
        const int pbeg = service ? a.ntheta + lane : threadIdx.x;
        const int pend = service ? a.nflat : a.ntheta;
        const int pstep = service ? 32 : ncomp_threads;
        if (!skip) {
            for (int p = pbeg; p < pend; p += pstep) {
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
                    dx = mt * rb1 / (sqrtf(vt * rb2) + a.eps) * a.eta;
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
                if constexpr (WPC > 1) {   // the tensor engine's tf32 images of the hidden-layer weights
                    if (c0 >= 0) E::patch_cell(eng, c0, th);
                    if (c1 >= 0) E::patch_cell(eng, c1, th);
                }
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
            b1t *= a.beta1;
            b2t *= a.beta2;
            tdone++;
        } else {
            tskip++;
        }
        EH_STAMP(5)
        if (service) {
            // the global parameters' scalars of step s + 1 are in place: publish (consumers: PmCtx::wait_phi)
            __syncwarp();
            __threadfence_block();
            if (lane == 0) asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(phi_flag)), "r"(tag) : "memory");
            EH_STAMP_SVC(27)
            if (stream && !have_next && s + 1 < a.nsteps) {
                // consumer mode, the next batch had not landed when this step began: wait for it now (the compute warps are
                // parked on the tile's mbarrier), or learn that the burst is over
                if (wait_slot(s + 1, bnext)) {
                    store_bs(load_bs(bnext), s + 1);
                    __syncwarp();   // the scalars of all lanes before lane 0's arrive (release) on the tile barrier
                    if (lane == 0) issue_tile(bnext, par ^ 1);
                } else {
                    if (lane == 0) {
                        *stop_flag = 1u;
                        __threadfence_block();
                        mbar_arrive(&mbar[par ^ 1]);   // releases the compute warps, which then see the flag
                    }
                    __syncwarp();
                }
            }
        }
        // end-of-step barrier of the compute warps: all theta patches of the weight image and the next step's batch
        // scalars are visible
        if (!service) {
            if constexpr (WPC > 1) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // patched tf32 images -> tensor core
            bar_sync_id(1, blockDim.x);
            E::after_reduce(st, stage, lane);  // constant staging rows were overwritten by the scratch / vectors
        }
        bcur = bnext;
        EH_STAMP(6)
    }
#undef EH_STAMP
#undef EH_STAMP_SVC

    // write back (CTA 0 holds the same state as everybody else)
    __syncthreads();
    if constexpr (WPC > 1) E::cta_exit(eng);
    if (bid == 0) {
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
}

}  // namespace eh
