// eh_engine_tc.cuh -- compute engine 4: the two contractions of the 16-wide hidden layer that are per-sample products
// (forward Z = a1 W2^T, backward data E = delta2 W2) on the 5th-generation tensor cores (tcgen05, operands and
// accumulators in TMEM), fp32-level accuracy through 3xTF32 splitting.  Persistent kernel only (eh_epoch_kernel.cuh).
//
// Why: with one sample per lane (EngFfma) each of these products costs 64 broadcast LDS.128 + 128 FFMA2 per 32 samples
// and the step is bound by the shared-memory pipe.  Here a GROUP of four warps forms a
// 128-sample tile; a thread still owns one sample for everything that is per-sample (layer 1, activations, output layer,
// process model, loss, seeds) and for the staging rows of the weight-gradient phase (unchanged: HMMA 3xTF32 over the
// warp's feature-major tile, chunk_dw_phase_mma), but
//   * it writes its activation row (hi and lo tf32 halves, 2 x 16 columns) straight into TENSOR MEMORY with tcgen05.st:
//     lane = sample, column = feature is exactly the A-operand layout of an M = 128 MMA, so A never touches shared memory;
//   * one thread of the group issues 6 tcgen05.mma kind::tf32 (M = 128, N = 16, K = 8; D += Al Bh + Ah Bl + Ah Bh, two
//     k-steps) against the tf32 images of W2 in shared memory (512 bytes per MMA: unswizzled K-major planes
//     [k / 4][row][4 floats], LBO = plane, SBO = 128), commits to an mbarrier;
//   * every thread reads its row of the accumulator back with tcgen05.ld (16 columns) and goes on in registers.
//   Measured round trip (tools/tc_proto.cu): 537 cycles for split + st + 6 MMAs + commit + ld, hidden behind the other
//   three groups of the CTA.  Operands from shared memory instead (SS form) cost 4 KB of operand fetch per MMA, i.e.
//   ~50 cycles for each of these tiny MMAs (same prototype) -- that form lost to the FFMA2 engine.
//
// Numerics: same bar as the FFMA2 engine (tests/test_gpu_parity.py, test_gpu_baseline_sizes.py: 1e-5 against the float64 reference restatement).
#pragma once
#include <cstdint>
#include "eh_engine_ffma.cuh"

namespace eh {

namespace tc {
constexpr int WPS = 256;                       // bytes per weight plane: 16 rows x 16 bytes
// engine region of the CTA (EpochArgs::eng_off): forward image [k / 4][j][4] hi, lo; backward image [j / 4][k][4] hi, lo;
// 8 mbarriers (forward, backward per group); TMEM slot
constexpr int ENG_WF_H = 0, ENG_WF_L = 4 * WPS, ENG_WB_H = 8 * WPS, ENG_WB_L = 12 * WPS, ENG_BAR = 16 * WPS, ENG_TMEM = ENG_BAR + 8 * 8,
              ENG_BYTES = ENG_TMEM + 16;
constexpr int COL_AH = 0, COL_AL = 16, COL_Z = 32, COLS_PER_GROUP = 64;

__device__ __forceinline__ uint32_t saddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// shared-memory matrix descriptor, no swizzle (sm_100 descriptor version 1)
__device__ __forceinline__ uint64_t desc(uint32_t sa, uint32_t lbo, uint32_t sbo)
{
    return (uint64_t)((sa & 0x3FFFF) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | ((uint64_t)1 << 46);
}
__host__ __device__ constexpr uint32_t idesc(int m, int n)   // kind::tf32, fp32 accumulate, A and B K-major
{
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// D[tmem] (+)= A[tmem] * B[smem]^T
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t id, uint32_t acc)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(id), "r"(acc) : "memory");
}
__device__ __forceinline__ void commit(uint32_t bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void ld16(uint32_t taddr, float (&r)[16])
{
    uint32_t u[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]),
          "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; i++) r[i] = __uint_as_float(u[i]);
}
__device__ __forceinline__ void st16(uint32_t taddr, const float (&r)[16])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
        "r"(__float_as_uint(r[0])), "r"(__float_as_uint(r[1])), "r"(__float_as_uint(r[2])), "r"(__float_as_uint(r[3])),
        "r"(__float_as_uint(r[4])), "r"(__float_as_uint(r[5])), "r"(__float_as_uint(r[6])), "r"(__float_as_uint(r[7])),
        "r"(__float_as_uint(r[8])), "r"(__float_as_uint(r[9])), "r"(__float_as_uint(r[10])), "r"(__float_as_uint(r[11])),
        "r"(__float_as_uint(r[12])), "r"(__float_as_uint(r[13])), "r"(__float_as_uint(r[14])), "r"(__float_as_uint(r[15]))
        : "memory");
}
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ bool bar_try(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// bounded: a tensor-core operation that never completes raises the error flag instead of hanging the GPU
__device__ __forceinline__ void bar_wait(uint32_t bar, uint32_t parity, unsigned* err)
{
    unsigned spins = 0;
    while (!bar_try(bar, parity))
        if (++spins > (1u << 24)) { *err = 1; break; }
}
// this thread's row of 16 values into the A-operand columns of the group: hi = the bits the tensor core reads (tf32:
// sign, exponent, 10 mantissa bits), lo = the exact remainder
__device__ __forceinline__ void st8(uint32_t taddr, const float (&r)[8])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(__float_as_uint(r[0])),
                 "r"(__float_as_uint(r[1])), "r"(__float_as_uint(r[2])), "r"(__float_as_uint(r[3])), "r"(__float_as_uint(r[4])),
                 "r"(__float_as_uint(r[5])), "r"(__float_as_uint(r[6])), "r"(__float_as_uint(r[7]))
                 : "memory");
}
// (eight values at a time, hi before lo: at most 16 staging registers are live next to the row itself)
__device__ __forceinline__ void put_row(uint32_t tm, const float2 (&v)[8])
{
#pragma unroll
    for (int q = 0; q < 2; q++) {
        float h[8], l[8];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            h[2 * k] = __uint_as_float(__float_as_uint(v[4 * q + k].x) & 0xffffe000u);
            h[2 * k + 1] = __uint_as_float(__float_as_uint(v[4 * q + k].y) & 0xffffe000u);
        }
        st8(tm + COL_AH + 8 * q, h);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            l[2 * k] = v[4 * q + k].x - h[2 * k];
            l[2 * k + 1] = v[4 * q + k].y - h[2 * k + 1];
        }
        st8(tm + COL_AL + 8 * q, l);
    }
    wait_st();
}
// the group's issuing thread: Z = A B^T with B = the weight image at `wh` / `wl` (hi / lo planes), two k-steps of 8
__device__ __forceinline__ void issue(uint32_t tmg, uint32_t wh, uint32_t wl, uint32_t bar)
{
    constexpr uint32_t id = idesc(128, 16);
#pragma unroll
    for (int ks = 0; ks < 2; ks++) {
        const uint64_t bh = desc(wh + ks * 2 * WPS, WPS, 128), bl = desc(wl + ks * 2 * WPS, WPS, 128);
        mma_ts(tmg + COL_Z, tmg + COL_AL + ks * 8, bh, id, ks ? 1u : 0u);   // small terms first
        mma_ts(tmg + COL_Z, tmg + COL_AH + ks * 8, bl, id, 1u);
        mma_ts(tmg + COL_Z, tmg + COL_AH + ks * 8, bh, id, 1u);
    }
    commit(bar);
}
__device__ __forceinline__ void group_sync(int g) { asm volatile("bar.sync %0, 128;" ::"r"(2 + g) : "memory"); }
}  // namespace tc

template <class C>
struct EngTc {
    using Cfg = C;
    using F = EngFfma<C>;
    static_assert(C::NH == 2 && C::H == 16 && C::LR == 1 && C::SPL == 1 && C::ACT != ACT_SWISH && F::DW_MMA,
                  "tensor engine: two hidden layers of 16, output layer in registers, tensor-pipe weight gradient");
    static constexpr int ENGINE = 4;
    static constexpr int WPC = 4;                          // warps per tile
    static constexpr int CHUNK = 128;                      // samples per tile
    static constexpr int MAX_WARPS = 16;
    static constexpr int STAGE_FLOATS = F::STAGE_FLOATS;   // per warp: feature-major staging tile + reduction row, as EngFfma
    static constexpr int NPART = C::NPART;
    static constexpr int OFF_STATS = C::D.npart_dw();
    static constexpr int ENG_FLOATS = tc::ENG_BYTES / 4;
    static constexpr bool SCRATCH_ALIASES_STAGE = false;

    struct State {
        int nacc;        // tiles accumulated into this warp's reduction row this step
        unsigned tiles;  // tiles this group has run since the launch (mbarrier phase)
        ChunkStats st;
        LastAcc<C> la;
        float4 r[C::R4 / 4];
        bool valid;
        uint32_t tm;     // TMEM address of the group's columns, lanes of this warp
        uint32_t bar;    // shared-memory address of the group's two mbarriers (forward, backward)
        uint32_t wf, wb; // shared-memory addresses of the forward / backward weight images (hi; lo follows 4 planes later)
    };

    // ---- CTA-level set-up (all threads; the caller synchronises afterwards) ----
    __device__ __forceinline__ static void cta_init(unsigned char* eng)
    {
        if (threadIdx.x == 0) {
#pragma unroll
            for (int i = 0; i < 8; i++)
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tc::saddr(eng + tc::ENG_BAR + 8 * i)) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        if (threadIdx.x < 32) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc::saddr(eng + tc::ENG_TMEM)), "r"(256u) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        tc::fence_before();
    }
    __device__ __forceinline__ static void cta_exit(unsigned char* eng)
    {
        tc::fence_before();
        __syncthreads();
        if (threadIdx.x < 32) {
            const uint32_t tm = *reinterpret_cast<const uint32_t*>(eng + tc::ENG_TMEM);
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(256u) : "memory");
        }
    }
    // tf32 images of W2[j][k] (j = output unit, k = input unit), hi and lo: forward B operand rows = j, K = k;
    // backward B operand rows = k, K = j
    __device__ __forceinline__ static void put_w2(unsigned char* eng, int k, int j, float w)
    {
        const float hi = __uint_as_float(__float_as_uint(w) & 0xffffe000u), lo = w - hi;
        const int of = (k >> 2) * tc::WPS + j * 16 + (k & 3) * 4, ob = (j >> 2) * tc::WPS + k * 16 + (j & 3) * 4;
        *reinterpret_cast<float*>(eng + tc::ENG_WF_H + of) = hi;
        *reinterpret_cast<float*>(eng + tc::ENG_WF_L + of) = lo;
        *reinterpret_cast<float*>(eng + tc::ENG_WB_H + ob) = hi;
        *reinterpret_cast<float*>(eng + tc::ENG_WB_L + ob) = lo;
    }
    __device__ __forceinline__ static void load_w2(unsigned char* eng, const float* sW)
    {
        for (int i = threadIdx.x; i < C::H * C::H; i += blockDim.x) put_w2(eng, i / C::H, i % C::H, sW[C::D.off_wf(2) + i]);
    }
    // the optimiser patched weight-image cell `cell` (k-major image of W2: cell = off_wf(2) + k H + j): keep the tf32 images in step
    __device__ __forceinline__ static void patch_cell(unsigned char* eng, int cell, float w)
    {
        const int i = cell - C::D.off_wf(2);
        if (i >= 0 && i < C::H * C::H) put_w2(eng, i / C::H, i % C::H, w);
    }

    __device__ __forceinline__ static void init_warp(State& s, float* stage, int lane, unsigned char* eng)
    {
        const int warp = threadIdx.x >> 5, g = warp >> 2;
        init_stage_rows<C>(stage, lane);
        tc::fence_after();
        const uint32_t tm = *reinterpret_cast<const volatile uint32_t*>(eng + tc::ENG_TMEM);
        s.tm = tm + (uint32_t)(g * tc::COLS_PER_GROUP) + ((uint32_t)((warp & 3) * 32) << 16);
        s.bar = tc::saddr(eng + tc::ENG_BAR + 16 * g);
        s.wf = tc::saddr(eng + tc::ENG_WF_H);
        s.wb = tc::saddr(eng + tc::ENG_WB_H);
        s.tiles = 0;
        s.nacc = 0;
    }
    __device__ __forceinline__ static void after_reduce(State&, float*, int) {}
    __device__ __forceinline__ static void step_begin(State& s, const float*, int)
    {
        s.nacc = 0;
#pragma unroll
        for (int t = 0; t < MAXT; t++) s.st.loss[t] = 0.f;
#pragma unroll
        for (int t = 0; t < MAXPS; t++) s.st.gphi[t] = 0.f;
        s.la.zero();
    }
    // record of this thread's sample of tile `chunk` (row = 32 * warp-in-group + lane)
    __device__ __forceinline__ static void fetch(State& s, const FetchArgs& fa, int chunk, int lane)
    {
        const int smp = chunk * CHUNK + (((threadIdx.x >> 5) & 3) << 5) + lane;
        const bool v = chunk < fa.nchunks && smp < fa.B;
        s.valid = v;
#pragma unroll
        for (int q = 0; q < C::R4 / 4; q++) s.r[q] = v ? fetch_rec4<C::R4 / 4>(fa, smp, q) : make_float4(0.f, 0.f, 0.f, 0.f);
    }

    // one 128-sample tile; `stage` = this warp's staging tile (+ reduction row)
    __device__ __forceinline__ static void chunk(State& s, const FetchArgs& fa, int next, const float* sW, const float* sS,
                                                 float* stage, int lane, const PSlot* slot, const int* loss_kind,
                                                 const PmCtx& cx, unsigned* err)
    {
        constexpr ShapeDims D = C::D;
        constexpr int P = C::P, H = C::H, NOUT = C::NOUT, T = C::T, NF = C::F, NPS = C::NPS, HP = C::H / 2, RS = C::RS;
        using PM = typename C::PM;
        const int warp = threadIdx.x >> 5, wq = warp & 3, g = warp >> 2;
        const bool leader = (wq == (g & 3)) && lane == 0;   // the groups' issuing threads sit on different schedulers
        const uint32_t ph = s.tiles & 1u;
        const uint32_t tmg = s.tm & 0x0000ffffu;            // lane 0 of the group's columns

        float rec[C::R4];
#pragma unroll
        for (int q = 0; q < C::R4 / 4; q++) {
            rec[4 * q] = s.r[q].x; rec[4 * q + 1] = s.r[q].y; rec[4 * q + 2] = s.r[q].z; rec[4 * q + 3] = s.r[q].w;
        }
        const bool valid = s.valid;
        fetch(s, fa, next, lane);

        // ---- layer 1 in fp32 (input BatchNorm(affine=false): (x - mu) * rstd with per-batch statistics) ----
        float2 hp[HP];
        {
            float x[P];
#pragma unroll
            for (int k = 0; k < P; k++) {
                x[k] = (rec[k] - sS[SS_BN + 2 * k]) * sS[SS_BN + 2 * k + 1];
                stage[goff<RS>(D.gA(1), k) + lane] = x[k];
            }
            const float4* b4 = reinterpret_cast<const float4*>(sW + D.off_b1());
#pragma unroll
            for (int j = 0; j < HP; j += 2) {
                const float4 b = b4[j >> 1];
                hp[j] = f2(b.x, b.y);
                hp[j + 1] = f2(b.z, b.w);
            }
#pragma unroll
            for (int k = 0; k < P; k++) {
                const float4* w4 = reinterpret_cast<const float4*>(sW + D.off_w1f() + k * H);
#pragma unroll
                for (int j = 0; j < HP; j += 2) {
                    const float4 w = w4[j >> 1];
                    hp[j] = fma2s(f2(w.x, w.y), x[k], hp[j]);
                    hp[j + 1] = fma2s(f2(w.z, w.w), x[k], hp[j + 1]);
                }
            }
            float2 aux;
#pragma unroll
            for (int j = 0; j < HP; j++) {
                hp[j] = act_fwd2<C::ACT>(hp[j], aux);
                stage[goff<RS>(D.gA(2), 2 * j) + lane] = hp[j].x;      // a1: input rows of the layer-2 weight gradient
                stage[goff<RS>(D.gA(2), 2 * j + 1) + lane] = hp[j].y;
            }
        }
        // ---- layer 2 on the tensor core: Z = a1 W2^T ----
        tc::put_row(s.tm, hp);
        tc::fence_before();
        tc::group_sync(g);
        if (leader) {
            tc::fence_after();
            tc::issue(tmg, s.wf, s.wf + 4 * tc::WPS, s.bar);
        }
        {
            float z[16];
            tc::bar_wait(s.bar, ph, err);
            tc::fence_after();
            tc::ld16(s.tm + tc::COL_Z, z);
            const float4* b4 = reinterpret_cast<const float4*>(sW + D.off_b(2));
            float2 aux;
#pragma unroll
            for (int j = 0; j < HP; j += 2) {
                const float4 b = b4[j >> 1];
                hp[j] = act_fwd2<C::ACT>(add2(f2(z[2 * j], z[2 * j + 1]), f2(b.x, b.y)), aux);
                hp[j + 1] = act_fwd2<C::ACT>(add2(f2(z[2 * j + 2], z[2 * j + 3]), f2(b.z, b.w)), aux);
            }
        }
        // linear output layer, dot form (hp = a2)
        float zo[NOUT];
#pragma unroll
        for (int o = 0; o < NOUT; o++) {
            const float4* w4 = reinterpret_cast<const float4*>(sW + D.off_wo() + o * H);
            float2 s0 = f2(sW[D.off_bo() + o], 0.f), s1 = f2s(0.f);
#pragma unroll
            for (int k = 0; k < HP; k += 2) {
                const float4 w = w4[k >> 1];
                s0 = fma2(f2(w.x, w.y), hp[k], s0);
                s1 = fma2(f2(w.z, w.w), hp[k + 1], s1);
            }
            const float2 t = add2(s0, s1);
            zo[o] = t.x + t.y;
        }
        // ---- process parameters, physics, masked residual, seeds (as chunk_sample_phase) ----
        cx.wait_phi();
        float dz[NOUT];
        {
            float f[NF > 0 ? NF : 1], y[T], pv[NPS], sg[NPS], yh[T], sv[PM::NSV], gy[T], gp[NPS];
#pragma unroll
            for (int k = 0; k < NF; k++) f[k] = rec[P + k];
#pragma unroll
            for (int k = 0; k < T; k++) y[k] = rec[P + NF + k];
            resolve_params<C>(slot, sS, zo, pv, sg, cx);
            PM::fwd(pv, f, cx, yh, sv);
#pragma unroll
            for (int t = 0; t < T; t++) {
                const bool m = valid && (y[t] == y[t]);
                const float r = m ? yh[t] - y[t] : 0.f;
                const float c = sS[SS_C + t];
                if (loss_kind[t] == LOSS_MAE) {
                    s.st.loss[t] += fabsf(r);
                    gy[t] = r > 0.f ? c : (r < 0.f ? -c : 0.f);
                } else {
                    s.st.loss[t] = fmaf(r, r, s.st.loss[t]);
                    gy[t] = 2.f * c * r;
                }
            }
            PM::bwd(pv, f, cx, yh, sv, gy, gp);
#pragma unroll
            for (int o = 0; o < NOUT; o++) dz[o] = 0.f;
#pragma unroll
            for (int q = 0; q < NPS; q++) {
                const PSlot sl = slot[q];
                if (sl.role == ROLE_NEURAL) {
                    float gq = gp[q];
                    if (C::PM::DYNAMIC ? (cx.scale_rt != 0) : C::SCALE) gq *= sl.span * sg[q] * (1.f - sg[q]);
#pragma unroll
                    for (int o = 0; o < NOUT; o++)
                        if (sl.idx == o) dz[o] += gq;
                } else if (sl.role == ROLE_GLOBAL) {
                    s.st.gphi[q] += gp[q];
                }
            }
        }
        // output-layer gradient in registers; delta2 = (Wo^T dz) .* act'(a2)
        float2 d[HP];
#pragma unroll
        for (int k = 0; k < HP; k++) d[k] = f2s(0.f);
#pragma unroll
        for (int o = 0; o < NOUT; o++) {
            s.la.b[o] += dz[o];
            const float4* w4 = reinterpret_cast<const float4*>(sW + D.off_wo() + o * H);
#pragma unroll
            for (int k = 0; k < HP; k += 2) {
                const float4 w = w4[k >> 1];
                s.la.w[o][k] = fma2s(hp[k], dz[o], s.la.w[o][k]);
                s.la.w[o][k + 1] = fma2s(hp[k + 1], dz[o], s.la.w[o][k + 1]);
                d[k] = fma2s(f2(w.x, w.y), dz[o], d[k]);
                d[k + 1] = fma2s(f2(w.z, w.w), dz[o], d[k + 1]);
            }
        }
#pragma unroll
        for (int k = 0; k < HP; k++) {
            d[k] = mul2(d[k], act_bwd2<C::ACT>(hp[k], f2s(0.f)));
            stage[goff<RS>(D.gD(2), 2 * k) + lane] = d[k].x;
            stage[goff<RS>(D.gD(2), 2 * k + 1) + lane] = d[k].y;
        }
        // ---- backward data pass on the tensor core: E = delta2 W2, delta1 = E .* act'(a1) ----
        tc::put_row(s.tm, d);
        tc::fence_before();
        tc::group_sync(g);
        if (leader) {
            tc::fence_after();
            tc::issue(tmg, s.wb, s.wb + 4 * tc::WPS, s.bar + 8);
        }
        {
            float e[16];
            tc::bar_wait(s.bar + 8, ph, err);
            tc::fence_after();
            tc::ld16(s.tm + tc::COL_Z, e);
#pragma unroll
            for (int k = 0; k < HP; k++) {
                const float2 a1 = f2(stage[goff<RS>(D.gA(2), 2 * k) + lane], stage[goff<RS>(D.gA(2), 2 * k + 1) + lane]);
                const float2 d1 = mul2(f2(e[2 * k], e[2 * k + 1]), act_bwd2<C::ACT>(a1, f2s(0.f)));
                stage[goff<RS>(D.gD(1), 2 * k) + lane] = d1.x;
                stage[goff<RS>(D.gD(1), 2 * k + 1) + lane] = d1.y;
            }
        }
        tc::fence_before();   // (orders this thread's tcgen05.ld before the next tile's MMAs, which follow a group barrier)
        __syncwarp();
        // ---- weight-gradient outer products over the warp's 32 samples (tensor pipe, 3xTF32; as EngFfma) ----
        chunk_dw_phase_mma<C>(stage, lane, stage + C::STAGE_FLOATS, s.nacc == 0);
        s.nacc++;
        s.tiles++;
        __syncwarp();
    }

    // the warps' reduction rows live next to their staging tiles, exactly as EngFfma's
    __device__ __forceinline__ static void reduce_prepare(State& s, float* work, unsigned*)
    {
        cta_reduce_prepare<C>(s.nacc, s.st, s.la, work + C::STAGE_FLOATS, STAGE_FLOATS);
    }
    __device__ __forceinline__ static float reduce_sum_at(const float* work, int nw, int p)
    {
        return cta_reduce_sum_at<C>(work + C::STAGE_FLOATS, STAGE_FLOATS, nw, p);
    }
};

}  // namespace eh
