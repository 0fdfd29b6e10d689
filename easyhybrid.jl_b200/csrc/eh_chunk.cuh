// eh_chunk.cuh -- the per-chunk device code shared by the fused step kernel (K1), the
// persistent epoch kernel and the eval kernel (sm_100a).
//
// Mapping (why): the small-MLP step is FP32-issue bound, not HBM bound,
// and at the benchmark batch (65 536 samples on 148 SMs) it is short of thread-level
// parallelism, so the layout maximises resident warps at a fixed instruction count:
//   * a lane owns ONE sample, a warp a 32-sample chunk end to end (2 048 chunks per batch);
//   * every multiply-add is a packed FFMA2 over a PAIR OF NEURONS: the weight pair comes
//     straight out of one LDS.128 (two pairs) of the reference's column-major weight image in
//     shared memory, the activation is the scalar-broadcast operand
//     (`FFMA2 Rd, Rw.F32x2, Ra.F32, Rc.F32x2`), so no register shuffling is needed;
//   * activations / deltas are staged feature-major in a warp-private shared-memory tile; the
//     weight-gradient outer products then run as lane-owned 4x4 register tiles whose operands
//     are LDS.128 over 4 consecutive samples (FFMA2 over sample pairs);
//   * no atomics: lane tiles -> per-warp -> per-CTA partial -> fixed-order reduction.
#pragma once
#include "eh_pm.cuh"

namespace eh {

// per-batch scalar row (floats): seed scale c_t, n_valid_t, SS_tot_t, then (mu, rstd) per chain input
constexpr int MAXP = 12;  // chain inputs
constexpr int EH_MAX_WORLD = 8;  // GPUs of one NVSwitch box
// ... then, for LOSS_AFFINE targets: seed coefficients sa, sb, sc and the loss value, written by the pre-pass (k_stat_seeds)
constexpr int BS_C = 0, BS_N = MAXT, BS_SS = 2 * MAXT, BS_BN = 3 * MAXT, BS_AFF = 3 * MAXT + 2 * MAXP, BS_STRIDE = BS_AFF + 4 * MAXT;

enum : int { OPT_ADAM = 0, OPT_ADAMW = 1, OPT_RMSPROP = 2, OPT_DESCENT = 3 };
enum : int { UPD_FULL = 0, UPD_REDUCE_ONLY = 1, UPD_FROM_VECTOR = 2 };

struct OptState {   // device-resident scalars
    float b1t, b2t; // running beta^t products, as Optimisers keeps them (start at beta)
    long long t;    // completed steps
    long long skipped;
};

struct PSlot {
    int role;      // ROLE_*
    int idx;       // NEURAL: chain output row; GLOBAL: index g into phi; FIXED: unused
    float lo;      // lower bound
    float span;    // upper - lower
    float fixedv;  // FIXED: default value
};

// SPL = samples per lane of the FFMA2 engine: 1 -> 32-sample chunks (most warps), 2 -> 64-sample chunks
// (every weight fetched from shared memory feeds two samples: half the LDS traffic per sample)
template <int P_, int NH_, int H_, int NOUT_, int ACT_, bool SCALE_, class PM_, int SPL_ = 1>
struct StepCfg {
    static constexpr int P = P_, NH = NH_, H = H_, NOUT = NOUT_, ACT = ACT_;
    static constexpr bool SCALE = SCALE_;
    static constexpr int SPL = SPL_;
    static constexpr int CHUNKS = 32 * SPL_;        // samples per warp pass
    static constexpr int RS = CHUNKS + 4;           // floats per staging row (bank skew)
    using PM = PM_;
    static constexpr int F = PM::NF, T = PM::NT, NPS = PM::NPS;
    // output-layer dW kept in registers when all per-lane scalars fit one 32-value transpose-reduce
    static constexpr int LR = (NOUT_ * (H_ + 1) + PM_::NT + PM_::NPS <= 32) ? 1 : 0;
    static constexpr ShapeDims D{P_, NH_, H_, NOUT_, LR};
    static constexpr int NLAST = NOUT_ * (H_ + 1);
    static constexpr int R4 = rup4(P_ + PM::NF + PM::NT);  // floats per record
    static constexpr int NB = D.nblocks();
    static constexpr int NBI = (NB + 31) / 32;             // dW tiles per lane
    // the warps' reduction rows keep the dW cells element-major with an ODD block stride (cell e * NBP + b): conflict-free for
    // the lanes of the dW phase (consecutive b) and nearly so for the CTA-level sum, which walks the partial vector in its
    // own tile-major order (consecutive e) -- with the even stride NB that pass ran into 4-way bank conflicts
    static constexpr int NBP = NB | 1;
    static constexpr int SCR_PAD = (NBP - NB) * 16;        // floats the row is longer than the partial vector
    static constexpr int SCR = D.npart() + SCR_PAD;
    static constexpr int GS = 4 * RS + 4;                  // floats per 4-row staging group (bank skew, see goff)
    static constexpr int AUXOFF = D.ngroups() * GS;        // swish sigma rows start here
    static constexpr int STAGE_FLOATS = AUXOFF + (ACT_ == ACT_SWISH ? NH_ * H_ * RS : 0);  // per warp
    static constexpr int NW = D.nweights();
    static constexpr int NPART = D.npart();
    static_assert(P_ <= MAXP && H_ % 4 == 0 && NOUT_ <= 4 && (SPL_ == 1 || SPL_ == 2), "shape limits");
};

// shared memory carve-up (floats): [weights NW pad4][scalars 128][per-warp stage ...]
//   scalars: [0..8) uniform slot values, [16..48) per-slot derived scalars, [48..52) c_t, [56..80) BN (mu, rstd),
//   [96..104) persistent kernel: span * sigma'(phi_g) of the global parameters, kept from their last update
//   [104..116) seed coefficients sa, sb, sc of LOSS_AFFINE targets
constexpr int SS_SLOT = 0, SS_PMS = 16, SS_C = 48, SS_NV = 52, SS_BN = 56, SS_NV2 = 80, SS_SGD = 96, SS_AFF = 104, SS_FLOATS = 128;

// float offset of the staging row of feature k inside 4-row group g0 (+k/4).  Groups are 4 rows of RS floats
// plus 4 floats of skew: the group stride is 20 mod 32 banks, so the 8 groups a quarter-warp of dW tiles
// reads with LDS.128 land on 8 disjoint bank quads.
template <int RS>
__device__ __forceinline__ constexpr int goff(int g0, int k) { return (g0 + (k >> 2)) * (4 * RS + 4) + (k & 3) * RS; }

__device__ __forceinline__ float comp(const float2* v, int k) { return (k & 1) ? v[k >> 1].y : v[k >> 1].x; }

// staging tile access: lane l owns columns SPL*l .. SPL*l + SPL - 1 of a row (one 4- or 8-byte access)
template <int S>
__device__ __forceinline__ void stage_put(float* row, int lane, const float (&v)[S])
{
    if (S == 2) *reinterpret_cast<float2*>(row + 2 * lane) = f2(v[0], v[S - 1]);
    else row[lane] = v[0];
}
template <int S>
__device__ __forceinline__ void stage_get(const float* row, int lane, float (&v)[S])
{
    if (S == 2) {
        float2 t = *reinterpret_cast<const float2*>(row + 2 * lane);
        v[0] = t.x;
        v[S - 1] = t.y;
    } else {
        v[0] = row[lane];
    }
}

// activation of the unit pair (j2, j2 + 1) of hidden layer l: the chain's activation, or -- run-time compiled models whose
// chains differ in activation (MultiNNHybridModel with an activation per parameter, GenericHybridModel.jl:168-174) -- the
// one the generated functor names for each unit.  l and j2 are constants after unrolling, so nothing is selected at run time.
template <class C>
__device__ __forceinline__ float2 unit_act_fwd2(int l, int j2, float2 z, float2& aux)
{
    if constexpr (C::PM::UNIT_ACT) {
        const int cx = C::PM::unit_act(l, j2), cy = C::PM::unit_act(l, j2 + 1);
        float2 ax = f2s(0.f), ay = f2s(0.f);
        const float2 rx = act_fwd2_rt(cx, z, ax);
        const float2 ry = (cy == cx) ? rx : act_fwd2_rt(cy, z, ay);
        aux = f2(ax.x, cy == cx ? ax.y : ay.y);
        return f2(rx.x, ry.y);
    } else {
        return act_fwd2<C::ACT>(z, aux);
    }
}
template <class C>
__device__ __forceinline__ float2 unit_act_bwd2(int l, int j2, float2 a, float2 aux)
{
    if constexpr (C::PM::UNIT_ACT) {
        const int cx = C::PM::unit_act(l, j2), cy = C::PM::unit_act(l, j2 + 1);
        const float2 rx = act_bwd2_rt(cx, a, aux);
        const float2 ry = (cy == cx) ? rx : act_bwd2_rt(cy, a, aux);
        return f2(rx.x, ry.y);
    } else {
        return act_bwd2<C::ACT>(a, aux);
    }
}

// Dense chain forward for the sample of this lane (prepare_hidden_chain,
// src/models/NNModels.jl:225-230).  hp holds neuron PAIRS; returns a_NH in hp, outputs in zo.
template <class C, bool STAGE>
__device__ __forceinline__ void chain_forward(const float* sW, float* stage, int lane, const float* x, float2* hp,
                                              float* zo, const unsigned* pass = nullptr)
{
    constexpr ShapeDims D = C::D;
    constexpr int P = C::P, NH = C::NH, H = C::H, NOUT = C::NOUT, HP = C::H / 2;
    constexpr int RS = C::RS;
    static_assert(!STAGE || C::SPL == 1, "staging forward is the single-sample form");
    {
        const float4* b4 = reinterpret_cast<const float4*>(sW + D.off_b1());
#pragma unroll
        for (int j = 0; j < HP; j += 2) {
            float4 b = b4[j >> 1];
            hp[j] = f2(b.x, b.y);
            hp[j + 1] = f2(b.z, b.w);
        }
#pragma unroll
        for (int k = 0; k < P; k++) {
            const float4* w4 = reinterpret_cast<const float4*>(sW + D.off_w1f() + k * H);
#pragma unroll
            for (int j = 0; j < HP; j += 2) {
                float4 w = w4[j >> 1];
                hp[j] = fma2s(f2(w.x, w.y), x[k], hp[j]);
                hp[j + 1] = fma2s(f2(w.z, w.w), x[k], hp[j + 1]);
            }
        }
    }
#pragma unroll
    for (int l = 1; l <= NH; l++) {
#pragma unroll
        for (int j = 0; j < HP; j++) {
            float2 aux = f2s(0.f);
            const float2 z = hp[j];
            hp[j] = unit_act_fwd2<C>(l, 2 * j, hp[j], aux);
            if (C::PM::DYNAMIC && pass) {
                const unsigned m = pass[l - 1] >> (2 * j);
                if (m & 1u) hp[j].x = z.x;
                if (m & 2u) hp[j].y = z.y;
            }
            if (STAGE && l + 1 <= D.nlt()) {
                stage[goff<RS>(D.gA(l + 1), 2 * j) + lane] = hp[j].x;
                stage[goff<RS>(D.gA(l + 1), 2 * j + 1) + lane] = hp[j].y;
            }
            if (STAGE) {
                if (C::ACT == ACT_SWISH) {
                    stage[C::AUXOFF + ((l - 1) * H + 2 * j) * RS + lane] = aux.x;
                    stage[C::AUXOFF + ((l - 1) * H + 2 * j + 1) * RS + lane] = aux.y;
                }
            }
        }
        if (l < NH) {
            float2 z[HP];
            const float4* b4 = reinterpret_cast<const float4*>(sW + D.off_b(l + 1));
#pragma unroll
            for (int j = 0; j < HP; j += 2) {
                float4 b = b4[j >> 1];
                z[j] = f2(b.x, b.y);
                z[j + 1] = f2(b.z, b.w);
            }
#pragma unroll
            for (int k = 0; k < H; k++) {
                const float4* w4 = reinterpret_cast<const float4*>(sW + D.off_wf(l + 1) + k * H);
                const float ak = comp(hp, k);
#pragma unroll
                for (int j = 0; j < HP; j += 2) {
                    float4 w = w4[j >> 1];
                    z[j] = fma2s(f2(w.x, w.y), ak, z[j]);
                    z[j + 1] = fma2s(f2(w.z, w.w), ak, z[j + 1]);
                }
            }
#pragma unroll
            for (int j = 0; j < HP; j++) hp[j] = z[j];
        }
    }
    // linear output layer, dot form: pairs over k
#pragma unroll
    for (int o = 0; o < NOUT; o++) {
        const float4* w4 = reinterpret_cast<const float4*>(sW + D.off_wo() + o * H);
        float2 s0 = f2(sW[D.off_bo() + o], 0.f), s1 = f2s(0.f);
#pragma unroll
        for (int k = 0; k < HP; k += 2) {
            float4 w = w4[k >> 1];
            s0 = fma2(f2(w.x, w.y), hp[k], s0);
            s1 = fma2(f2(w.z, w.w), hp[k + 1], s1);
        }
        float2 s = add2(s0, s1);
        zo[o] = s.x + s.y;
    }
}

// Process parameters from their roles (GenericHybridModel.jl:377-414): NEURAL = chain output
// row (sigmoid-squashed into [lo, hi] iff scale_nn_outputs), GLOBAL / FIXED = per-step uniform
// value from shared memory.  sg keeps sigma(z) for the backward.
template <class C>
__device__ __forceinline__ void resolve_params(const PSlot* slot, const float* sS, const float* zo, float* pv, float* sg,
                                               const PmCtx& cx)
{
    const bool scale = C::PM::DYNAMIC ? (cx.scale_rt != 0) : C::SCALE;
#pragma unroll
    for (int s = 0; s < C::NPS; s++) {
        const PSlot sl = slot[s];
        sg[s] = 0.f;
        if (sl.role == ROLE_NEURAL) {
            float z = zo[0];
#pragma unroll
            for (int o = 1; o < C::NOUT; o++)
                if (sl.idx == o) z = zo[o];
            if (scale) {
                sg[s] = sigmoid1(z);
                pv[s] = fmaf(sg[s], sl.span, sl.lo);
            } else {
                pv[s] = z;
            }
        } else {
            pv[s] = sS[SS_SLOT + s];
        }
    }
}

// weight image + per-step scalars -> shared memory.  pblock = [theta/phi | slot values | slot scalars].
// LDCG: the block may have been rewritten by another CTA / kernel since this SM last cached it.
template <class C>
__device__ __forceinline__ void load_weights_and_scalars(const float* pblock, int nflat, const int* wsrc,
                                                         const float* bscal, int use_bn, float* sW, float* sS)
{
    for (int i = threadIdx.x; i < C::NW; i += blockDim.x) {
        int s = wsrc[i];
        sW[i] = s >= 0 ? __ldcg(pblock + s) : (s == -2 ? 1.f : 0.f);   // -2: the 1 of a pass-through unit
    }
    if (threadIdx.x < MAXPS) sS[SS_SLOT + threadIdx.x] = __ldcg(pblock + nflat + threadIdx.x);
    if (threadIdx.x < MAXPS * PMS_PER_SLOT) sS[SS_PMS + threadIdx.x] = __ldcg(pblock + nflat + MAXPS + threadIdx.x);
    if (threadIdx.x < MAXT) sS[SS_C + threadIdx.x] = bscal ? bscal[BS_C + threadIdx.x] : 0.f;
    if (threadIdx.x < 2 * C::P)
        sS[SS_BN + threadIdx.x] = use_bn ? bscal[BS_BN + threadIdx.x] : ((threadIdx.x & 1) ? 1.f : 0.f);
    if (threadIdx.x < 3 * MAXT) sS[SS_AFF + threadIdx.x] = bscal ? bscal[BS_AFF + threadIdx.x] : 0.f;
}

// constant rows of a warp's staging tile: the "1" feature of every augmented input, zero padding
template <class C>
__device__ __forceinline__ void init_stage_rows(float* stage, int lane)
{
    constexpr ShapeDims D = C::D;
    constexpr int RS = C::RS, S = C::SPL;
    float one[S], zero[S];
#pragma unroll
    for (int i = 0; i < S; i++) { one[i] = 1.f; zero[i] = 0.f; }
#pragma unroll
    for (int l = 1; l <= D.nlt(); l++) {
        const int din = D.din(l), ka = D.ka(l), gA = D.gA(l);
#pragma unroll
        for (int k = din; k < ka; k++) stage_put<S>(stage + goff<RS>(gA, k), lane, (k == din) ? one : zero);
    }
    if (!C::LR) {
#pragma unroll
        for (int o = C::NOUT; o < rup4(C::NOUT); o++) stage_put<S>(stage + goff<RS>(D.gD(C::NH + 1), o), lane, zero);
    }
}

// dW tile coordinates of this lane: float offsets of the delta / activation groups of tile b = lane + 32 i
template <class C>
__device__ __forceinline__ void tile_rows(int lane, int* rowD, int* rowA)
{
    constexpr ShapeDims D = C::D;
#pragma unroll
    for (int i = 0; i < C::NBI; i++) {
        int b = lane + 32 * i;
        rowD[i] = 0;
        rowA[i] = 0;
#pragma unroll
        for (int l = 1; l <= D.nlt(); l++) {
            const int b0 = D.blk0(l), nk = D.nk(l), nb = D.nj(l) * nk;
            if (b >= b0 && b < b0 + nb) {
                int jb = (b - b0) / nk, kb = (b - b0) % nk;
                rowD[i] = (D.gD(l) + jb) * C::GS;
                rowA[i] = (D.gA(l) + kb) * C::GS;
            }
        }
    }
}

// where a batch's samples come from
struct FetchArgs {
    const float4* rec;
    const int* idx;       // NULL: records rec_base .. rec_base + B
    long long rec_base;
    int B, nchunks;
    // optional: the records of samples [tile_s0, ...) of this batch already staged in shared memory (persistent
    // kernel: one cp.async.bulk per CTA and step, eh_epoch_kernel.cuh); fetch then reads the tile instead of HBM
    const float4* tile;
    int tile_s0;
};

// record q-th float4 of batch sample `smp` (valid sample): shared-memory tile if staged, else HBM (gather through idx
// or contiguous)
template <int R44>
__device__ __forceinline__ float4 fetch_rec4(const FetchArgs& fa, int smp, int q)
{
    if (fa.tile) return fa.tile[(size_t)(smp - fa.tile_s0) * R44 + q];
    long long i = fa.rec_base + smp;
    if (fa.idx) i = fa.idx[smp];
    return __ldg(fa.rec + i * R44 + q);
}

struct ChunkStats {
    float loss[MAXT];    // sum r^2 (or |r|) per target
    float gphi[MAXPS];   // sum g * dy/dslot for GLOBAL slots
};

// lane-owned gradient of the linear output layer (LR shapes): [NOUT][H/2] weight pairs + [NOUT] bias
template <class C>
struct LastAcc {
    float2 w[C::LR ? C::NOUT : 1][C::H / 2];
    float b[C::LR ? C::NOUT : 1];
    __device__ __forceinline__ void zero()
    {
#pragma unroll
        for (int o = 0; o < (C::LR ? C::NOUT : 1); o++) {
            b[o] = 0.f;
#pragma unroll
            for (int k = 0; k < C::H / 2; k++) w[o][k] = f2s(0.f);
        }
    }
};

// ---- per-sample phase: forward, physics, masked residual, backward data pass, staging ----
// A lane owns S = SPL samples (columns S*lane .. of the staging rows).  Every weight pair fetched from
// shared memory (one LDS.128 = two pairs) is used for all S samples: FFMA2 over NEURON pairs with the
// activation of sample s as the scalar-broadcast operand.
// rec[s]: record of sample s of this lane (canonical order), valid[s]: sample exists.
template <class C>
__device__ __forceinline__ void chunk_sample_phase(const float (*rec)[C::R4], const bool* valid, const float* sW,
                                                   const float* sS, float* stage, int lane, const PSlot* slot,
                                                   const int* loss_kind, const PmCtx& cx, ChunkStats& st, LastAcc<C>& la)
{
    constexpr ShapeDims D = C::D;
    constexpr int P = C::P, NH = C::NH, H = C::H, NOUT = C::NOUT, T = C::T, F = C::F, NPS = C::NPS, HP = C::H / 2;
    constexpr int RS = C::RS, S = C::SPL;
    using PM = typename C::PM;

    float x[S][P];
#pragma unroll
    for (int k = 0; k < P; k++) {
        float v[S];
#pragma unroll
        for (int s = 0; s < S; s++) {
            // input BatchNorm(affine=false): (x - mu) * rstd with per-batch statistics (identity: mu 0, rstd 1)
            x[s][k] = (rec[s][k] - sS[SS_BN + 2 * k]) * sS[SS_BN + 2 * k + 1];
            v[s] = x[s][k];
        }
        stage_put<S>(stage + goff<RS>(D.gA(1), k), lane, v);
    }

    // ---- forward (prepare_hidden_chain, src/models/NNModels.jl:225-230) ----
    float2 hp[S][HP];
    {
        const float4* b4 = reinterpret_cast<const float4*>(sW + D.off_b1());
#pragma unroll
        for (int j = 0; j < HP; j += 2) {
            float4 b = b4[j >> 1];
#pragma unroll
            for (int s = 0; s < S; s++) { hp[s][j] = f2(b.x, b.y); hp[s][j + 1] = f2(b.z, b.w); }
        }
#pragma unroll
        for (int k = 0; k < P; k++) {
            const float4* w4 = reinterpret_cast<const float4*>(sW + D.off_w1f() + k * H);
#pragma unroll
            for (int j = 0; j < HP; j += 2) {
                float4 w = w4[j >> 1];
#pragma unroll
                for (int s = 0; s < S; s++) {
                    hp[s][j] = fma2s(f2(w.x, w.y), x[s][k], hp[s][j]);
                    hp[s][j + 1] = fma2s(f2(w.z, w.w), x[s][k], hp[s][j + 1]);
                }
            }
        }
    }
#pragma unroll
    for (int l = 1; l <= NH; l++) {
#pragma unroll
        for (int j = 0; j < HP; j++) {
            float lo[S], hi[S], axl[S], axh[S];
#pragma unroll
            for (int s = 0; s < S; s++) {
                float2 aux = f2s(0.f);
                const float2 z = hp[s][j];
                hp[s][j] = unit_act_fwd2<C>(l, 2 * j, hp[s][j], aux);
                if (C::PM::DYNAMIC) {   // pass-through units of a shallower chain keep z
                    const unsigned m = cx.pass[l - 1] >> (2 * j);
                    if (m & 1u) hp[s][j].x = z.x;
                    if (m & 2u) hp[s][j].y = z.y;
                }
                lo[s] = hp[s][j].x; hi[s] = hp[s][j].y; axl[s] = aux.x; axh[s] = aux.y;
            }
            if (l + 1 <= D.nlt()) {
                stage_put<S>(stage + goff<RS>(D.gA(l + 1), 2 * j), lane, lo);
                stage_put<S>(stage + goff<RS>(D.gA(l + 1), 2 * j + 1), lane, hi);
            }
            if (C::ACT == ACT_SWISH) {
                stage_put<S>(stage + C::AUXOFF + ((l - 1) * H + 2 * j) * RS, lane, axl);
                stage_put<S>(stage + C::AUXOFF + ((l - 1) * H + 2 * j + 1) * RS, lane, axh);
            }
        }
        if (l < NH) {
            float2 z[S][HP];
            const float4* b4 = reinterpret_cast<const float4*>(sW + D.off_b(l + 1));
#pragma unroll
            for (int j = 0; j < HP; j += 2) {
                float4 b = b4[j >> 1];
#pragma unroll
                for (int s = 0; s < S; s++) { z[s][j] = f2(b.x, b.y); z[s][j + 1] = f2(b.z, b.w); }
            }
#pragma unroll
            for (int k = 0; k < H; k++) {
                const float4* w4 = reinterpret_cast<const float4*>(sW + D.off_wf(l + 1) + k * H);
                float ak[S];
#pragma unroll
                for (int s = 0; s < S; s++) ak[s] = comp(hp[s], k);
#pragma unroll
                for (int j = 0; j < HP; j += 2) {
                    float4 w = w4[j >> 1];
#pragma unroll
                    for (int s = 0; s < S; s++) {
                        z[s][j] = fma2s(f2(w.x, w.y), ak[s], z[s][j]);
                        z[s][j + 1] = fma2s(f2(w.z, w.w), ak[s], z[s][j + 1]);
                    }
                }
            }
#pragma unroll
            for (int s = 0; s < S; s++)
#pragma unroll
                for (int j = 0; j < HP; j++) hp[s][j] = z[s][j];
        }
    }
    // linear output layer, dot form: pairs over k
    float zo[S][NOUT];
#pragma unroll
    for (int o = 0; o < NOUT; o++) {
        const float4* w4 = reinterpret_cast<const float4*>(sW + D.off_wo() + o * H);
        float2 s0[S], s1[S];
#pragma unroll
        for (int s = 0; s < S; s++) { s0[s] = f2(sW[D.off_bo() + o], 0.f); s1[s] = f2s(0.f); }
#pragma unroll
        for (int k = 0; k < HP; k += 2) {
            float4 w = w4[k >> 1];
#pragma unroll
            for (int s = 0; s < S; s++) {
                s0[s] = fma2(f2(w.x, w.y), hp[s][k], s0[s]);
                s1[s] = fma2(f2(w.z, w.w), hp[s][k + 1], s1[s]);
            }
        }
#pragma unroll
        for (int s = 0; s < S; s++) {
            float2 t = add2(s0[s], s1[s]);
            zo[s][o] = t.x + t.y;
        }
    }

    // ---- process parameters (GenericHybridModel.jl:377-414), physics (:425), masked residual, seeds ----
    cx.wait_phi();   // persistent kernel: the global parameters' derived scalars of this step are in place
    float dz[S][NOUT];
#pragma unroll
    for (int s = 0; s < S; s++) {
        float f[F > 0 ? F : 1], y[T], pv[NPS], sg[NPS], yh[T], sv[PM::NSV], gy[T], gp[NPS];
#pragma unroll
        for (int k = 0; k < F; k++) f[k] = rec[s][P + k];
#pragma unroll
        for (int k = 0; k < T; k++) y[k] = rec[s][P + F + k];
        resolve_params<C>(slot, sS, zo[s], pv, sg, cx);
        PM::fwd(pv, f, cx, yh, sv);
        // valid_mask = !isnan(y) (train.jl:221-232); seeds dL/dyhat
#pragma unroll
        for (int t = 0; t < T; t++) {
            const bool m = valid[s] && (y[t] == y[t]);
            const float r = m ? yh[t] - y[t] : 0.f;
            const float c = sS[SS_C + t];
            if (loss_kind[t] == LOSS_MAE) {
                st.loss[t] += fabsf(r);
                gy[t] = r > 0.f ? c : (r < 0.f ? -c : 0.f);
            } else if (loss_kind[t] == LOSS_AFFINE) {
                // seeds of the prediction-statistics losses (pre-pass): sa + sb yhat + sc y, zero where the target is missing
                st.loss[t] = fmaf(r, r, st.loss[t]);
                gy[t] = m ? fmaf(sS[SS_AFF + MAXT + t], yh[t], fmaf(sS[SS_AFF + 2 * MAXT + t], y[t], sS[SS_AFF + t])) : 0.f;
            } else {
                st.loss[t] = fmaf(r, r, st.loss[t]);
                gy[t] = 2.f * c * r;
            }
        }
        PM::bwd(pv, f, cx, yh, sv, gy, gp);
#pragma unroll
        for (int o = 0; o < NOUT; o++) dz[s][o] = 0.f;
#pragma unroll
        for (int q = 0; q < NPS; q++) {
            const PSlot sl = slot[q];
            if (sl.role == ROLE_NEURAL) {
                float g = gp[q];
                if (C::PM::DYNAMIC ? (cx.scale_rt != 0) : C::SCALE) g *= sl.span * sg[q] * (1.f - sg[q]);
#pragma unroll
                for (int o = 0; o < NOUT; o++)
                    if (sl.idx == o) dz[s][o] += g;
            } else if (sl.role == ROLE_GLOBAL) {
                st.gphi[q] += gp[q];
            }
        }
    }
    if (C::LR) {
        // output-layer weight gradient in registers: dWo[o][k] += dz_o a_NH[k], dbo[o] += dz_o
#pragma unroll
        for (int o = 0; o < NOUT; o++)
#pragma unroll
            for (int s = 0; s < S; s++) {
                la.b[o] += dz[s][o];
#pragma unroll
                for (int k = 0; k < HP; k++) la.w[o][k] = fma2s(hp[s][k], dz[s][o], la.w[o][k]);
            }
    } else {
#pragma unroll
        for (int o = 0; o < NOUT; o++) {
            float v[S];
#pragma unroll
            for (int s = 0; s < S; s++) v[s] = dz[s][o];
            stage_put<S>(stage + goff<RS>(D.gD(NH + 1), o), lane, v);
        }
    }

    // ---- backward data pass: delta_l for l = NH .. 1 (hp still holds a_NH), neuron pairs ----
    float2 d[S][HP];
#pragma unroll
    for (int s = 0; s < S; s++)
#pragma unroll
        for (int k = 0; k < HP; k++) d[s][k] = f2s(0.f);
#pragma unroll
    for (int o = 0; o < NOUT; o++) {
        const float4* w4 = reinterpret_cast<const float4*>(sW + D.off_wo() + o * H);
#pragma unroll
        for (int k = 0; k < HP; k += 2) {
            float4 w = w4[k >> 1];
#pragma unroll
            for (int s = 0; s < S; s++) {
                d[s][k] = fma2s(f2(w.x, w.y), dz[s][o], d[s][k]);
                d[s][k + 1] = fma2s(f2(w.z, w.w), dz[s][o], d[s][k + 1]);
            }
        }
    }
#pragma unroll
    for (int l = NH; l >= 1; l--) {
        // times act'(a_l); a_l (and sigma for swish) come back from the staging tile for l < NH
#pragma unroll
        for (int k = 0; k < HP; k++) {
            float alo[S], ahi[S], xlo[S], xhi[S], dlo[S], dhi[S];
            if (l < NH) {
                stage_get<S>(stage + goff<RS>(D.gA(l + 1), 2 * k), lane, alo);
                stage_get<S>(stage + goff<RS>(D.gA(l + 1), 2 * k + 1), lane, ahi);
            }
            if (C::ACT == ACT_SWISH) {
                stage_get<S>(stage + C::AUXOFF + ((l - 1) * H + 2 * k) * RS, lane, xlo);
                stage_get<S>(stage + C::AUXOFF + ((l - 1) * H + 2 * k + 1) * RS, lane, xhi);
            }
#pragma unroll
            for (int s = 0; s < S; s++) {
                float2 al = (l < NH) ? f2(alo[s], ahi[s]) : hp[s][k];
                float2 aux = (C::ACT == ACT_SWISH) ? f2(xlo[s], xhi[s]) : f2s(0.f);
                float2 ga = unit_act_bwd2<C>(l, 2 * k, al, aux);
                if (C::PM::DYNAMIC) {
                    const unsigned m = cx.pass[l - 1] >> (2 * k);
                    if (m & 1u) ga.x = 1.f;
                    if (m & 2u) ga.y = 1.f;
                }
                d[s][k] = mul2(d[s][k], ga);
                dlo[s] = d[s][k].x;
                dhi[s] = d[s][k].y;
            }
            stage_put<S>(stage + goff<RS>(D.gD(l), 2 * k), lane, dlo);
            stage_put<S>(stage + goff<RS>(D.gD(l), 2 * k + 1), lane, dhi);
        }
        if (l > 1) {
            // delta_{l-1}[k] = sum_j W_l[j][k] delta_l[j]  (j-major copy of W_l, pairs over k)
            float2 dn[S][HP];
#pragma unroll
            for (int s = 0; s < S; s++)
#pragma unroll
                for (int k = 0; k < HP; k++) dn[s][k] = f2s(0.f);
#pragma unroll
            for (int j = 0; j < H; j++) {
                const float4* w4 = reinterpret_cast<const float4*>(sW + D.off_wb(l) + j * H);
                float dj[S];
#pragma unroll
                for (int s = 0; s < S; s++) dj[s] = comp(d[s], j);
#pragma unroll
                for (int k = 0; k < HP; k += 2) {
                    float4 w = w4[k >> 1];
#pragma unroll
                    for (int s = 0; s < S; s++) {
                        dn[s][k] = fma2s(f2(w.x, w.y), dj[s], dn[s][k]);
                        dn[s][k + 1] = fma2s(f2(w.z, w.w), dj[s], dn[s][k + 1]);
                    }
                }
            }
#pragma unroll
            for (int s = 0; s < S; s++)
#pragma unroll
                for (int k = 0; k < HP; k++) d[s][k] = dn[s][k];
        }
    }
}

// ---- weight-gradient phase: lane = one 4x4 tile of some layer's dW (bias = the "1" row) ----
// The tile accumulators live in registers only inside this phase: between chunks (and for the CTA
// reduction) they rest in the warp's row `wacc` of the reduction scratch, element-major
// (cell e*NB + b: consecutive lanes on consecutive banks).  `first`: no earlier chunk this step.
template <class C>
__device__ __forceinline__ void chunk_dw_phase(const float* stage, int lane, const int* rowD, const int* rowA,
                                               float* wacc, bool first)
{
    constexpr int RS = C::RS;
#pragma unroll
    for (int i = 0; i < C::NBI; i++) {
        if (lane + 32 * i < C::NB) {
            const int b = lane + 32 * i;
            float2 acc[16];
#pragma unroll
            for (int e = 0; e < 16; e++) acc[e] = f2(first ? 0.f : wacc[e * C::NBP + b], 0.f);
            const float* pd = stage + rowD[i];
            const float* pa = stage + rowA[i];
#pragma unroll 2
            for (int c = 0; c < C::CHUNKS; c += 4) {
                float4 dv[4], av[4];
#pragma unroll
                for (int j = 0; j < 4; j++) dv[j] = *reinterpret_cast<const float4*>(pd + j * RS + c);
#pragma unroll
                for (int k = 0; k < 4; k++) av[k] = *reinterpret_cast<const float4*>(pa + k * RS + c);
#pragma unroll
                for (int j = 0; j < 4; j++)
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        acc[j * 4 + k] = fma2(f2(dv[j].x, dv[j].y), f2(av[k].x, av[k].y), acc[j * 4 + k]);
                        acc[j * 4 + k] = fma2(f2(dv[j].z, dv[j].w), f2(av[k].z, av[k].w), acc[j * 4 + k]);
                    }
            }
#pragma unroll
            for (int e = 0; e < 16; e++) wacc[e * C::NBP + b] = acc[e].x + acc[e].y;
        }
    }
}

// ---- weight-gradient phase on the tensor pipe (3xTF32, fp32-level accuracy) ----
// dW_l[j][k] += sum over the chunk's samples of delta_l[j][s] * a_{l-1}[k][s] is an (H x samples) x (samples x ka)
// product: mma.m16n8k8 with the delta rows as A (row-major M x K, K = samples) and the activation rows as B (K x N),
// both read straight from the feature-major staging tiles (rows of 32 samples, conflict-free in the k = sample
// direction).  Compared with the 4x4 register tiles this trades 64 LDS.128 + 256 FFMA2 per chunk for 64 LDS.32 +
// 48 HMMA, and halves the accumulator registers.  Columns past ka(l) of the last n-tile read whatever rows follow in
// the staging tile (finite values) and are never stored.  Same wacc layout as chunk_dw_phase.
template <class C>
__device__ __forceinline__ void chunk_dw_phase_mma(const float* stage, int lane, float* wacc, bool first)
{
    constexpr ShapeDims D = C::D;
    constexpr int RS = C::RS;
    static_assert(C::H % 16 == 0 && C::LR == 1 && C::SPL == 1, "tensor-pipe dW: hidden width multiple of 16, output layer in registers");
    constexpr int MT = C::H / 16;
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int l = 1; l <= D.nlt(); l++) {
        const int ka = D.ka(l), nk = D.nk(l), b0 = D.blk0(l);
        const int NT = (ka + 7) / 8;
        float acc[MT][6][4];   // NT <= 6 for the shapes compiled (ka <= 48)
#pragma unroll
        for (int mt = 0; mt < MT; mt++)
#pragma unroll
            for (int nt = 0; nt < 6; nt++)
#pragma unroll
                for (int q = 0; q < 4; q++) acc[mt][nt][q] = 0.f;
#pragma unroll
        for (int ks = 0; ks < C::CHUNKS / 8; ks++) {
            const int k0 = 8 * ks + t;
            unsigned bh[6][2], bl[6][2];
#pragma unroll
            for (int nt = 0; nt < 6; nt++) {
                if (nt < NT) {
                    const float* pb = stage + goff<RS>(D.gA(l), 8 * nt + g) + k0;
                    split_tf32(pb[0], bh[nt][0], bl[nt][0]);
                    split_tf32(pb[4], bh[nt][1], bl[nt][1]);
                }
            }
#pragma unroll
            for (int mt = 0; mt < MT; mt++) {
                const float* pa0 = stage + goff<RS>(D.gD(l), 16 * mt + g) + k0;
                const float* pa1 = stage + goff<RS>(D.gD(l), 16 * mt + g + 8) + k0;
                unsigned ah[4], al[4];
                split_tf32(pa0[0], ah[0], al[0]);
                split_tf32(pa1[0], ah[1], al[1]);
                split_tf32(pa0[4], ah[2], al[2]);
                split_tf32(pa1[4], ah[3], al[3]);
#pragma unroll
                for (int nt = 0; nt < 6; nt++)
                    if (nt < NT) mma_3xtf32(acc[mt][nt], ah, al, bh[nt], bl[nt]);
            }
        }
        // C fragment: rows g / g + 8 of the m-tile, columns 2t / 2t + 1 of the n-tile -> 4x4 tile cells of the partial vector
#pragma unroll
        for (int mt = 0; mt < MT; mt++)
#pragma unroll
            for (int nt = 0; nt < 6; nt++) {
                if (nt < NT) {
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const int j = 16 * mt + g + ((q >> 1) ? 8 : 0), k = 8 * nt + 2 * t + (q & 1);
                        if (k < ka) {
                            const int cell = ((j & 3) * 4 + (k & 3)) * C::NBP + b0 + (j >> 2) * nk + (k >> 2);
                            wacc[cell] = first ? acc[mt][nt][q] : wacc[cell] + acc[mt][nt][q];
                        }
                    }
                }
            }
    }
}

// Sum NV per-lane values over the 32 lanes of a warp with a recursive-halving exchange: at each of
// the 5 levels a lane keeps half of its values and ships the other half to its partner, so the whole
// reduction costs NV-ish shuffles instead of 5*NV (SHFL shares the 128 B/clk LSU writeback path, which
// is the scarce resource of this kernel).  On return lane l holds the total of value index
// slot_of_lane(l) in v[0]; fixed exchange order -> bitwise reproducible.
template <int NV>
__device__ __forceinline__ void warp_transpose_reduce(float (&v)[32], int lane)
{
    static_assert(NV <= 32, "at most 32 values");
#pragma unroll
    for (int i = NV; i < 32; i++) v[i] = 0.f;
#pragma unroll
    for (int lvl = 0; lvl < 5; lvl++) {
        const int half = 16 >> lvl;          // values kept after this level
        const int mask = 16 >> lvl;          // partner = lane ^ mask
        const bool up = (lane & mask) != 0;
#pragma unroll
        for (int i = 0; i < half; i++) {
            // lanes with the bit set keep the upper half [half, 2 half), others the lower half
            float keep = up ? v[i + half] : v[i];
            float send = up ? v[i] : v[i + half];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, mask);
        }
    }
}
// value index whose total ends up in lane l after warp_transpose_reduce
__device__ __forceinline__ int transpose_reduce_slot(int lane)
{
    // level k (mask 16>>k) chooses the upper half when the lane bit is set: index bits from MSB down
    return lane;  // bit (4-k) of the index == bit (4-k) of the lane
}

// ---- CTA-level fixed-order reduction of the warps' rows (dW tiles + statistics) into out[NPART] ----
// scratch: the warps' rows of NPART floats, `stride` floats apart; the dW cells of a row were left there by
// chunk_dw_phase (element-major; the tile-major order of the partial vector is restored by the summing
// pass below), `nacc` = chunks this warp accumulated this step (0: its dW cells are stale and count as zero).
// (1) per warp, before the CTA barrier: the warp's statistics / output-layer sums into its row
template <class C>
__device__ __forceinline__ void cta_reduce_prepare(int nacc, const ChunkStats& st, const LastAcc<C>& la, float* scratch, int stride)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (nacc == 0)
        for (int q = lane; q < C::NBP * 16; q += 32) scratch[warp * stride + q] = 0.f;
    {
        // all per-lane scalars (loss sums, phi sums, output-layer gradient) in one exchange
        constexpr int NLASTV = C::LR ? C::NLAST : 0;
        constexpr int NV = C::T + C::NPS + NLASTV;
        static_assert(NV <= 32, "per-lane scalar set too large for one transpose-reduce");
        float v[32];
#pragma unroll
        for (int t = 0; t < C::T; t++) v[t] = st.loss[t];
#pragma unroll
        for (int s = 0; s < C::NPS; s++) v[C::T + s] = st.gphi[s];
        if (C::LR) {
#pragma unroll
            for (int o = 0; o < C::NOUT; o++) {
#pragma unroll
                for (int k = 0; k < C::H / 2; k++) {
                    v[C::T + C::NPS + o * (C::H + 1) + 2 * k] = la.w[o][k].x;
                    v[C::T + C::NPS + o * (C::H + 1) + 2 * k + 1] = la.w[o][k].y;
                }
                v[C::T + C::NPS + o * (C::H + 1) + C::H] = la.b[o];
            }
        }
        warp_transpose_reduce<NV>(v, lane);
        // lane l now holds the warp total of value l
        const int i = lane;
        int dst = -1;
        if (i < C::T) dst = C::D.npart_dw() + i;
        else if (i < C::T + C::NPS) dst = C::D.npart_dw() + MAXT + (i - C::T);
        else if (i < NV) dst = C::D.off_last() + (i - C::T - C::NPS);
        if (dst >= 0) scratch[warp * stride + C::SCR_PAD + dst] = v[0];
        // cells of the statistics block that no lane writes
        for (int q = C::D.npart_dw() + lane; q < C::NPART; q += 32) {
            bool used = (q < C::D.npart_dw() + C::T) || (q >= C::D.npart_dw() + MAXT && q < C::D.npart_dw() + MAXT + C::NPS) ||
                        (q >= C::D.off_last() && q < C::D.off_last() + NLASTV);
            if (!used) scratch[warp * stride + C::SCR_PAD + q] = 0.f;
        }
    }
}
// (2) after the barrier: position p of the partial vector summed over the first nw warps in fixed order (the cell of
// the element-major row it lives in: see StepCfg::NBP)
template <class C>
__device__ __forceinline__ float cta_reduce_sum_at(const float* scratch, int stride, int nw, int p)
{
    const int q = p < C::NB * 16 ? (p & 15) * C::NBP + (p >> 4) : p + C::SCR_PAD;
    // all rows are fetched before the first addition (independent loads; rows past nw count as +0, which leaves the sum's
    // bits alone): a rolled loop over a run-time warp count pays one shared-memory latency per row
    float r[16];
#pragma unroll
    for (int w = 0; w < 16; w++) r[w] = w < nw ? scratch[w * stride + q] : 0.f;
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 16; w++) s += r[w];
    for (int w = 16; w < nw; w++) s += scratch[w * stride + q];
    return s;
}

}  // namespace eh
