// eh_engine_ffma.cuh -- compute engine 0: exact-fp32 FFMA2 path (eh_chunk.cuh) behind the engine
// interface used by k_step / k_epoch.  A lane owns one sample; weights are broadcast from shared memory.
#pragma once
#include "eh_chunk.cuh"

namespace eh {

template <class C>
struct EngFfma {
    using Cfg = C;
    static constexpr int ENGINE = C::SPL == 2 ? 2 : 0;
    // per-warp shared memory: the staging tile, then the warp's row of the CTA reduction (dW tile sums between
    // chunks + its statistics); separate regions, so nothing has to be re-initialised after a reduction
    static constexpr int STAGE_FLOATS = C::STAGE_FLOATS + rup4(C::SCR);
    static constexpr int NPART = C::NPART;
    static constexpr int OFF_STATS = C::D.npart_dw();
    static constexpr int CHUNK = C::CHUNKS;           // samples per warp pass
    static constexpr int MAX_WARPS = C::SPL == 2 ? 10 : 16;
    // weight-gradient outer products on the tensor pipe (3xTF32) where the shape allows; the FFMA2 register tiles otherwise
#ifdef EH_DW_FFMA
    static constexpr bool DW_MMA = false;
#else
    static constexpr bool DW_MMA = C::LR == 1 && C::H % 16 == 0 && C::SPL == 1 && C::D.ka(C::D.nlt()) <= 48;
#endif

    struct State {
        int nacc;  // chunks accumulated into this warp's reduction row this step
        ChunkStats st;
        LastAcc<C> la;
        int rowD[C::NBI], rowA[C::NBI];
        float4 r[C::SPL][C::R4 / 4];
        bool valid[C::SPL];
    };

    __device__ __forceinline__ static void init_warp(State& s, float* stage, int lane)
    {
        init_stage_rows<C>(stage, lane);
        tile_rows<C>(lane, s.rowD, s.rowA);
    }
    __device__ __forceinline__ static void after_reduce(State&, float*, int) {}
    __device__ __forceinline__ static void step_begin(State& s, const float* sW, int lane)
    {
        s.nacc = 0;
#pragma unroll
        for (int t = 0; t < MAXT; t++) s.st.loss[t] = 0.f;
#pragma unroll
        for (int t = 0; t < MAXPS; t++) s.st.gphi[t] = 0.f;
        s.la.zero();
    }
    // prefetch the records of chunk `chunk`: lane l owns samples SPL*l .. SPL*l + SPL - 1 of the chunk
    __device__ __forceinline__ static void fetch(State& s, const FetchArgs& fa, int chunk, int lane)
    {
#pragma unroll
        for (int sp = 0; sp < C::SPL; sp++) {
            const int smp = chunk * CHUNK + C::SPL * lane + sp;
            const bool v = chunk < fa.nchunks && smp < fa.B;
            s.valid[sp] = v;
#pragma unroll
            for (int q = 0; q < C::R4 / 4; q++)
                s.r[sp][q] = v ? fetch_rec4<C::R4 / 4>(fa, smp, q) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    // process the prefetched chunk; `next` (chunk index, may be past the end) is prefetched right after the
    // current records have been copied out, so its latency hides behind the compute
    __device__ __forceinline__ static void chunk(State& s, const FetchArgs& fa, int next, const float* sW, const float* sS,
                                                 float* stage, int lane, const PSlot* slot, const int* loss_kind,
                                                 const PmCtx& cx)
    {
        float rec[C::SPL][C::R4];
        bool valid[C::SPL];
#pragma unroll
        for (int sp = 0; sp < C::SPL; sp++) {
#pragma unroll
            for (int q = 0; q < C::R4 / 4; q++) {
                rec[sp][4 * q] = s.r[sp][q].x; rec[sp][4 * q + 1] = s.r[sp][q].y;
                rec[sp][4 * q + 2] = s.r[sp][q].z; rec[sp][4 * q + 3] = s.r[sp][q].w;
            }
            valid[sp] = s.valid[sp];
        }
        fetch(s, fa, next, lane);
        chunk_sample_phase<C>(rec, valid, sW, sS, stage, lane, slot, loss_kind, cx, s.st, s.la);
        __syncwarp();
        if constexpr (DW_MMA) chunk_dw_phase_mma<C>(stage, lane, stage + C::STAGE_FLOATS, s.nacc == 0);
        else chunk_dw_phase<C>(stage, lane, s.rowD, s.rowA, stage + C::STAGE_FLOATS, s.nacc == 0);
        s.nacc++;
        __syncwarp();
    }
    // the warps' reduction rows live next to (not inside) their staging tiles
    static constexpr bool SCRATCH_ALIASES_STAGE = false;
    // per warp, before the CTA barrier
    __device__ __forceinline__ static void reduce_prepare(State& s, float* work)
    {
        cta_reduce_prepare<C>(s.nacc, s.st, s.la, work + C::STAGE_FLOATS, STAGE_FLOATS);
    }
    // after the barrier: position p of the partial vector summed over the first nw warps
    __device__ __forceinline__ static float reduce_sum_at(const float* work, int nw, int p)
    {
        return cta_reduce_sum_at<C>(work + C::STAGE_FLOATS, STAGE_FLOATS, nw, p);
    }
};

}  // namespace eh
