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
constexpr int BS_C = 0, BS_N = MAXT, BS_SS = 2 * MAXT, BS_BN = 3 * MAXT, BS_STRIDE = 3 * MAXT + 2 * MAXP;

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

template <int P_, int NH_, int H_, int NOUT_, int ACT_, bool SCALE_, class PM_>
struct StepCfg {
    static constexpr int P = P_, NH = NH_, H = H_, NOUT = NOUT_, ACT = ACT_;
    static constexpr bool SCALE = SCALE_;
    using PM = PM_;
    static constexpr int F = PM::NF, T = PM::NT, NPS = PM::NPS;
    // output-layer dW kept in registers when all per-lane scalars fit one 32-value transpose-reduce
    static constexpr int LR = (NOUT_ * (H_ + 1) + PM_::NT + PM_::NPS <= 32) ? 1 : 0;
    static constexpr ShapeDims D{P_, NH_, H_, NOUT_, LR};
    static constexpr int NLAST = NOUT_ * (H_ + 1);
    static constexpr int R4 = rup4(P_ + PM::NF + PM::NT);  // floats per record
    static constexpr int NB = D.nblocks();
    static constexpr int NBI = (NB + 31) / 32;             // dW tiles per lane
    static constexpr int NROWS = D.nrows() + (ACT_ == ACT_SWISH ? NH_ * H_ : 0);
    static constexpr int AUXROW0 = D.nrows();              // swish sigma rows
    static constexpr int STAGE_FLOATS = NROWS * ROWSTRIDE;  // per warp
    static constexpr int NW = D.nweights();
    static constexpr int NPART = D.npart();
    static_assert(P_ <= MAXP && H_ % 4 == 0 && NOUT_ <= 4, "shape limits");
};

// shared memory carve-up (floats): [weights NW pad4][scalars 128][per-warp stage ...]
//   scalars: [0..8) uniform slot values, [16..48) per-slot derived scalars, [48..52) c_t, [56..80) BN (mu, rstd)
constexpr int SS_SLOT = 0, SS_PMS = 16, SS_C = 48, SS_NV = 52, SS_BN = 56, SS_FLOATS = 128;

// row of feature k inside 4-row group g0 (+k/4): groups start every 5 rows (bank skew)
__device__ __forceinline__ constexpr int grow(int g0, int k) { return 5 * (g0 + (k >> 2)) + (k & 3); }

__device__ __forceinline__ float comp(const float2* v, int k) { return (k & 1) ? v[k >> 1].y : v[k >> 1].x; }

// Dense chain forward for the sample of this lane (prepare_hidden_chain,
// src/models/NNModels.jl:225-230).  hp holds neuron PAIRS; returns a_NH in hp, outputs in zo.
template <class C, bool STAGE>
__device__ __forceinline__ void chain_forward(const float* sW, float* stage, int lane, const float* x, float2* hp,
                                              float* zo)
{
    constexpr ShapeDims D = C::D;
    constexpr int P = C::P, NH = C::NH, H = C::H, NOUT = C::NOUT, HP = C::H / 2;
    constexpr int RS = ROWSTRIDE;
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
            hp[j] = act_fwd2<C::ACT>(hp[j], aux);
            if (STAGE && l + 1 <= D.nlt()) {
                stage[grow(D.gA(l + 1), 2 * j) * RS + lane] = hp[j].x;
                stage[grow(D.gA(l + 1), 2 * j + 1) * RS + lane] = hp[j].y;
            }
            if (STAGE) {
                if (C::ACT == ACT_SWISH) {
                    stage[(C::AUXROW0 + (l - 1) * H + 2 * j) * RS + lane] = aux.x;
                    stage[(C::AUXROW0 + (l - 1) * H + 2 * j + 1) * RS + lane] = aux.y;
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
__device__ __forceinline__ void resolve_params(const PSlot* slot, const float* sS, const float* zo, float* pv, float* sg)
{
#pragma unroll
    for (int s = 0; s < C::NPS; s++) {
        const PSlot sl = slot[s];
        sg[s] = 0.f;
        if (sl.role == ROLE_NEURAL) {
            float z = zo[0];
#pragma unroll
            for (int o = 1; o < C::NOUT; o++)
                if (sl.idx == o) z = zo[o];
            if (C::SCALE) {
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
        sW[i] = s >= 0 ? __ldcg(pblock + s) : 0.f;
    }
    if (threadIdx.x < MAXPS) sS[SS_SLOT + threadIdx.x] = __ldcg(pblock + nflat + threadIdx.x);
    if (threadIdx.x < MAXPS * PMS_PER_SLOT) sS[SS_PMS + threadIdx.x] = __ldcg(pblock + nflat + MAXPS + threadIdx.x);
    if (threadIdx.x < MAXT) sS[SS_C + threadIdx.x] = bscal ? bscal[BS_C + threadIdx.x] : 0.f;
    if (threadIdx.x < 2 * C::P)
        sS[SS_BN + threadIdx.x] = use_bn ? bscal[BS_BN + threadIdx.x] : ((threadIdx.x & 1) ? 1.f : 0.f);
}

// constant rows of a warp's staging tile: the "1" feature of every augmented input, zero padding
template <class C>
__device__ __forceinline__ void init_stage_rows(float* stage, int lane)
{
    constexpr ShapeDims D = C::D;
    constexpr int RS = ROWSTRIDE;
#pragma unroll
    for (int l = 1; l <= D.nlt(); l++) {
        const int din = D.din(l), ka = D.ka(l), gA = D.gA(l);
#pragma unroll
        for (int k = din; k < ka; k++) stage[grow(gA, k) * RS + lane] = (k == din) ? 1.f : 0.f;
    }
    if (!C::LR) {
#pragma unroll
        for (int o = C::NOUT; o < rup4(C::NOUT); o++) stage[grow(D.gD(C::NH + 1), o) * RS + lane] = 0.f;
    }
}

// dW tile coordinates of this lane: rows of the delta / activation groups of tile b = lane + 32 i
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
                rowD[i] = 5 * (D.gD(l) + jb);
                rowA[i] = 5 * (D.gA(l) + kb);
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
};

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
// rec: this lane's record (canonical order), valid: sample exists.
template <class C>
__device__ __forceinline__ void chunk_sample_phase(const float* rec, bool valid, const float* sW, const float* sS,
                                                   float* stage, int lane, const PSlot* slot, const int* loss_kind,
                                                   const PmCtx& cx, ChunkStats& st, LastAcc<C>& la)
{
    constexpr ShapeDims D = C::D;
    constexpr int P = C::P, NH = C::NH, H = C::H, NOUT = C::NOUT, T = C::T, F = C::F, NPS = C::NPS, HP = C::H / 2;
    constexpr int RS = ROWSTRIDE;
    using PM = typename C::PM;

    float x[P], f[F > 0 ? F : 1], y[T];
#pragma unroll
    for (int k = 0; k < P; k++) {
        // input BatchNorm(affine=false): (x - mu) * rstd with per-batch statistics (identity: mu 0, rstd 1)
        x[k] = (rec[k] - sS[SS_BN + 2 * k]) * sS[SS_BN + 2 * k + 1];
        stage[grow(D.gA(1), k) * RS + lane] = x[k];
    }
#pragma unroll
    for (int k = 0; k < F; k++) f[k] = rec[P + k];
#pragma unroll
    for (int k = 0; k < T; k++) y[k] = rec[P + F + k];

    float2 hp[HP];
    float zo[NOUT];
    chain_forward<C, true>(sW, stage, lane, x, hp, zo);

    // process parameters (GenericHybridModel.jl:377-414) and physics (:425)
    float pv[NPS], sg[NPS], yh[T], sv[4], gy[T], gp[NPS];
    resolve_params<C>(slot, sS, zo, pv, sg);
    PM::fwd(pv, f, cx, yh, sv);
    // masked residual: valid_mask = !isnan(y) (train.jl:221-232); seeds dL/dyhat
#pragma unroll
    for (int t = 0; t < T; t++) {
        const bool m = valid && (y[t] == y[t]);
        const float r = m ? yh[t] - y[t] : 0.f;
        const float c = sS[SS_C + t];
        if (loss_kind[t] == LOSS_MAE) {
            st.loss[t] += fabsf(r);
            gy[t] = r > 0.f ? c : (r < 0.f ? -c : 0.f);
        } else {
            st.loss[t] = fmaf(r, r, st.loss[t]);
            gy[t] = 2.f * c * r;
        }
    }
    PM::bwd(pv, f, cx, yh, sv, gy, gp);

    // delta at the linear output layer; phi statistics
    float dz[NOUT];
#pragma unroll
    for (int o = 0; o < NOUT; o++) dz[o] = 0.f;
#pragma unroll
    for (int s = 0; s < NPS; s++) {
        const PSlot sl = slot[s];
        if (sl.role == ROLE_NEURAL) {
            float g = gp[s];
            if (C::SCALE) g *= sl.span * sg[s] * (1.f - sg[s]);
#pragma unroll
            for (int o = 0; o < NOUT; o++)
                if (sl.idx == o) dz[o] += g;
        } else if (sl.role == ROLE_GLOBAL) {
            st.gphi[s] += gp[s];
        }
    }
    if (C::LR) {
        // output-layer weight gradient in registers: dWo[o][k] += dz_o a_NH[k], dbo[o] += dz_o
#pragma unroll
        for (int o = 0; o < NOUT; o++) {
            la.b[o] += dz[o];
#pragma unroll
            for (int k = 0; k < HP; k++) la.w[o][k] = fma2s(hp[k], dz[o], la.w[o][k]);
        }
    } else {
#pragma unroll
        for (int o = 0; o < NOUT; o++) stage[grow(D.gD(NH + 1), o) * RS + lane] = dz[o];
    }

    // backward data pass: delta_l for l = NH .. 1 (hp still holds a_NH), neuron pairs
    float2 d[HP];
#pragma unroll
    for (int k = 0; k < HP; k++) d[k] = f2s(0.f);
#pragma unroll
    for (int o = 0; o < NOUT; o++) {
        const float4* w4 = reinterpret_cast<const float4*>(sW + D.off_wo() + o * H);
#pragma unroll
        for (int k = 0; k < HP; k += 2) {
            float4 w = w4[k >> 1];
            d[k] = fma2s(f2(w.x, w.y), dz[o], d[k]);
            d[k + 1] = fma2s(f2(w.z, w.w), dz[o], d[k + 1]);
        }
    }
#pragma unroll
    for (int l = NH; l >= 1; l--) {
        // times act'(a_l); a_l (and sigma for swish) come back from the staging tile for l < NH
#pragma unroll
        for (int k = 0; k < HP; k++) {
            float2 al = hp[k];
            if (l < NH) al = f2(stage[grow(D.gA(l + 1), 2 * k) * RS + lane], stage[grow(D.gA(l + 1), 2 * k + 1) * RS + lane]);
            float2 aux = f2s(0.f);
            if (C::ACT == ACT_SWISH)
                aux = f2(stage[(C::AUXROW0 + (l - 1) * H + 2 * k) * RS + lane],
                         stage[(C::AUXROW0 + (l - 1) * H + 2 * k + 1) * RS + lane]);
            d[k] = mul2(d[k], act_bwd2<C::ACT>(al, aux));
            stage[grow(D.gD(l), 2 * k) * RS + lane] = d[k].x;
            stage[grow(D.gD(l), 2 * k + 1) * RS + lane] = d[k].y;
        }
        if (l > 1) {
            // delta_{l-1}[k] = sum_j W_l[j][k] delta_l[j]  (j-major copy of W_l, pairs over k)
            float2 dn[HP];
#pragma unroll
            for (int k = 0; k < HP; k++) dn[k] = f2s(0.f);
#pragma unroll
            for (int j = 0; j < H; j++) {
                const float4* w4 = reinterpret_cast<const float4*>(sW + D.off_wb(l) + j * H);
                const float dj = comp(d, j);
#pragma unroll
                for (int k = 0; k < HP; k += 2) {
                    float4 w = w4[k >> 1];
                    dn[k] = fma2s(f2(w.x, w.y), dj, dn[k]);
                    dn[k + 1] = fma2s(f2(w.z, w.w), dj, dn[k + 1]);
                }
            }
#pragma unroll
            for (int k = 0; k < HP; k++) d[k] = dn[k];
        }
    }
}

// ---- weight-gradient phase: lane = one 4x4 tile of some layer's dW (bias = the "1" row) ----
template <class C>
__device__ __forceinline__ void chunk_dw_phase(const float* stage, int lane, const int* rowD, const int* rowA,
                                               float2 (*acc)[16])
{
    constexpr int RS = ROWSTRIDE;
#pragma unroll
    for (int i = 0; i < C::NBI; i++) {
        if (lane + 32 * i < C::NB) {
            const float* pd = stage + rowD[i] * RS;
            const float* pa = stage + rowA[i] * RS;
#pragma unroll 2
            for (int c = 0; c < CHUNK; c += 4) {
                float4 dv[4], av[4];
#pragma unroll
                for (int j = 0; j < 4; j++) dv[j] = *reinterpret_cast<const float4*>(pd + j * RS + c);
#pragma unroll
                for (int k = 0; k < 4; k++) av[k] = *reinterpret_cast<const float4*>(pa + k * RS + c);
#pragma unroll
                for (int j = 0; j < 4; j++)
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        acc[i][j * 4 + k] = fma2(f2(dv[j].x, dv[j].y), f2(av[k].x, av[k].y), acc[i][j * 4 + k]);
                        acc[i][j * 4 + k] = fma2(f2(dv[j].z, dv[j].w), f2(av[k].z, av[k].w), acc[i][j * 4 + k]);
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

// ---- CTA-level fixed-order reduction of lane tiles + statistics into out[NPART] ----
// scratch: [nwarps][NPART] floats (may alias the staging tiles; caller syncs before).
template <class C>
__device__ __forceinline__ void cta_reduce(const float2 (*acc)[16], const ChunkStats& st, const LastAcc<C>& la,
                                           float* scratch, float* out, int out_is_global)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
#pragma unroll
    for (int i = 0; i < C::NBI; i++) {
        int b = lane + 32 * i;
        if (b < C::NB) {
#pragma unroll
            for (int e = 0; e < 16; e++) scratch[warp * C::NPART + b * 16 + e] = acc[i][e].x + acc[i][e].y;
        }
    }
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
        if (dst >= 0) scratch[warp * C::NPART + dst] = v[0];
        // cells of the statistics block that no lane writes
        for (int q = C::D.npart_dw() + lane; q < C::NPART; q += 32) {
            bool used = (q < C::D.npart_dw() + C::T) || (q >= C::D.npart_dw() + MAXT && q < C::D.npart_dw() + MAXT + C::NPS) ||
                        (q >= C::D.off_last() && q < C::D.off_last() + NLASTV);
            if (!used) scratch[warp * C::NPART + q] = 0.f;
        }
    }
    __syncthreads();
    for (int p = threadIdx.x; p < C::NPART; p += blockDim.x) {
        float s = 0.f;
        for (int w = 0; w < nwarps; w++) s += scratch[w * C::NPART + p];
        if (out_is_global) __stcg(out + p, s);
        else out[p] = s;
    }
}

}  // namespace eh
