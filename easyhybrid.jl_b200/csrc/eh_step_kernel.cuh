// eh_step_kernel.cuh -- K1: the fused hybrid training-step kernel (sm_100a).
//
// One launch = one minibatch: Dense-chain forward -> parameter squashing ->
// process model -> masked residual -> hand-derived backward -> per-CTA partial
// sums of the (already loss-scaled) gradient and of the loss statistics.
// K2 (eh_update_kernel.cuh) reduces the partials in fixed order and applies the
// optimiser.  Replaces, per step, the call
//   Lux.Training.single_train_step!(backend, loss_fn, batch, train_state)
// at src/training/epoch.jl:20-26 (forward: src/models/GenericHybridModel.jl:370-431,
// loss: src/losses/compute_loss.jl:20-35, src/losses/loss_fn.jl:58-81).
//
// Mapping (why): the small-MLP step is FP32-issue bound, not HBM bound
// (DESIGN.md section 4), so the layout minimises issued instructions per sample:
//   * a lane owns TWO samples packed in f32x2 registers: every multiply-add is an
//     FFMA2 whose weight operand is a scalar broadcast straight from one LDS.128
//     (4 weights -> 4 FFMA2); weights live in shared memory in the reference's
//     column-major layout (+ a transposed copy for the backward data pass);
//   * a warp owns a 64-sample chunk end to end; activations / deltas are staged
//     feature-major in a warp-private shared-memory tile so the weight-gradient
//     outer products become 4x4 register-blocked tiles with both operands read as
//     LDS.128 over 4 consecutive samples (again FFMA2 over sample pairs);
//   * no atomics anywhere: lane-owned dW tiles -> per-warp -> per-CTA partial ->
//     fixed-order second pass (K2).
#pragma once
#include "eh_pm.cuh"

namespace eh {

// per-batch scalar row (floats): seed scale c_t, n_valid_t, SS_tot_t, then (mu, rstd) per chain input
constexpr int BS_C = 0, BS_N = MAXT, BS_SS = 2 * MAXT, BS_BN = 3 * MAXT, BS_STRIDE = 3 * MAXT + 2 * 12;
constexpr int MAXP = 12;  // chain inputs

struct PSlot {
    int role;      // ROLE_*
    int idx;       // NEURAL: chain output row; GLOBAL: index g into phi; FIXED: unused
    float lo;      // lower bound
    float span;    // upper - lower
    float fixedv;  // FIXED: default value
};

struct StepArgs {
    const float4* rec;        // packed records, canonical order [x(P) | f(F) | y(T) | pad], R4 floats each
    const int* idx;           // 0-based sample ids of this batch (NULL: records rec_base .. rec_base+B)
    long long rec_base;
    int B;                    // samples in this batch
    const float* theta;       // flat parameters (reference ComponentArray order)
    const int* wsrc;          // [nweights] flat index feeding each smem weight cell, -1 = zero pad
    const float* bscal;       // per-batch scalar row (BS_* layout) of this batch
    float* partial;           // [gridDim.x][npart] per-CTA partial sums
    int npart;
    int loss_kind[MAXT];
    PSlot slot[MAXPS];
    float pmc[4];             // process-model constants
    int use_bn;
};

template <int P_, int NH_, int H_, int NOUT_, int ACT_, bool SCALE_, class PM_>
struct StepCfg {
    static constexpr int P = P_, NH = NH_, H = H_, NOUT = NOUT_, ACT = ACT_;
    static constexpr bool SCALE = SCALE_;
    using PM = PM_;
    static constexpr int F = PM::NF, T = PM::NT, NPS = PM::NPS;
    static constexpr ShapeDims D{P_, NH_, H_, NOUT_};
    static constexpr int R4 = rup4(P_ + PM::NF + PM::NT);  // floats per record
    static constexpr int NB = D.nblocks();
    static constexpr int NBI = (NB + 31) / 32;             // dW tiles per lane
    static constexpr int NROWS = D.nrows() + (ACT_ == ACT_SWISH ? NH_ * H_ : 0);
    static constexpr int AUXROW0 = D.nrows();              // swish sigma rows
    static constexpr int STAGE_FLOATS = NROWS * ROWSTRIDE;  // per warp
    static constexpr int NW = D.nweights();
    static constexpr int NPART = D.npart_dw() + NSTAT;
};

// shared memory carve-up (floats): [weights NW pad4][scalars 64][per-warp stage ...]
template <class C>
__host__ __device__ constexpr int step_smem_floats(int nwarps)
{
    return rup4(C::NW) + 64 + nwarps * C::STAGE_FLOATS;
}

template <class C>
struct StepCtx {
    PmScal pms;
    bool slot_uniform[MAXPS];
};

// row of feature k inside 4-row group g0 (+k/4): groups start every 5 rows
__device__ __forceinline__ constexpr int grow(int g0, int k) { return 5 * (g0 + (k >> 2)) + (k & 3); }


// Dense chain forward for the two samples of a lane (prepare_hidden_chain,
// src/models/NNModels.jl:225-230): hidden layers in outer-product form (loop over
// inputs k, 4 output neurons per LDS.128 of the column-major weight image), linear
// output layer in dot form.  STAGE: also write a_l (and swish sigma) feature-major
// into the warp's staging tile.  h returns a_NH, zo the chain outputs.
template <class C, bool STAGE>
__device__ __forceinline__ void chain_forward(const float* sW, float* stage, int lane, const float2* x, float2* h,
                                              float2* zo)
{
    constexpr ShapeDims D = C::D;
    constexpr int P = C::P, NH = C::NH, H = C::H, NOUT = C::NOUT;
    constexpr int RS = ROWSTRIDE;
    {
        const float4* b4 = reinterpret_cast<const float4*>(sW + D.off_b1());
#pragma unroll
        for (int j = 0; j < H; j += 4) {
            float4 b = b4[j >> 2];
            h[j] = f2s(b.x); h[j + 1] = f2s(b.y); h[j + 2] = f2s(b.z); h[j + 3] = f2s(b.w);
        }
#pragma unroll
        for (int k = 0; k < P; k++) {
            const float4* w4 = reinterpret_cast<const float4*>(sW + D.off_w1f() + k * H);
#pragma unroll
            for (int j = 0; j < H; j += 4) {
                float4 w = w4[j >> 2];
                h[j] = fma2s(x[k], w.x, h[j]);
                h[j + 1] = fma2s(x[k], w.y, h[j + 1]);
                h[j + 2] = fma2s(x[k], w.z, h[j + 2]);
                h[j + 3] = fma2s(x[k], w.w, h[j + 3]);
            }
        }
    }
#pragma unroll
    for (int l = 1; l <= NH; l++) {
#pragma unroll
        for (int j = 0; j < H; j++) {
            float2 aux = f2s(0.f);
            h[j] = act_fwd2<C::ACT>(h[j], aux);
            if (STAGE) {
                *reinterpret_cast<float2*>(stage + grow(D.gA(l + 1), j) * RS + 2 * lane) = h[j];
                if (C::ACT == ACT_SWISH)
                    *reinterpret_cast<float2*>(stage + (C::AUXROW0 + (l - 1) * H + j) * RS + 2 * lane) = aux;
            }
        }
        if (l < NH) {
            float2 z[H];
            const float4* b4 = reinterpret_cast<const float4*>(sW + D.off_b(l + 1));
#pragma unroll
            for (int j = 0; j < H; j += 4) {
                float4 b = b4[j >> 2];
                z[j] = f2s(b.x); z[j + 1] = f2s(b.y); z[j + 2] = f2s(b.z); z[j + 3] = f2s(b.w);
            }
#pragma unroll
            for (int k = 0; k < H; k++) {
                const float4* w4 = reinterpret_cast<const float4*>(sW + D.off_wf(l + 1) + k * H);
#pragma unroll
                for (int j = 0; j < H; j += 4) {
                    float4 w = w4[j >> 2];
                    z[j] = fma2s(h[k], w.x, z[j]);
                    z[j + 1] = fma2s(h[k], w.y, z[j + 1]);
                    z[j + 2] = fma2s(h[k], w.z, z[j + 2]);
                    z[j + 3] = fma2s(h[k], w.w, z[j + 3]);
                }
            }
#pragma unroll
            for (int j = 0; j < H; j++) h[j] = z[j];
        }
    }
#pragma unroll
    for (int o = 0; o < NOUT; o++) {
        const float4* w4 = reinterpret_cast<const float4*>(sW + D.off_wo() + o * H);
        float2 s0 = f2s(sW[D.off_bo() + o]), s1 = f2s(0.f);
#pragma unroll
        for (int k = 0; k < H; k += 4) {
            float4 w = w4[k >> 2];
            s0 = fma2s(h[k], w.x, s0);
            s1 = fma2s(h[k + 1], w.y, s1);
            s0 = fma2s(h[k + 2], w.z, s0);
            s1 = fma2s(h[k + 3], w.w, s1);
        }
        zo[o] = add2(s0, s1);
    }
}

// Process parameters from their roles (GenericHybridModel.jl:377-414): NEURAL = chain
// output row (sigmoid-squashed into [lo, hi] iff scale_nn_outputs), GLOBAL / FIXED =
// per-step uniform value prepared in shared memory.  sg keeps sigma(z) for the backward.
template <class C>
__device__ __forceinline__ void resolve_params(const PSlot* slot, const float* sS, const float2* zo, float2* pv,
                                               float2* sg)
{
#pragma unroll
    for (int s = 0; s < C::NPS; s++) {
        const PSlot sl = slot[s];
        if (sl.role == ROLE_NEURAL) {
            float2 z = zo[0];
#pragma unroll
            for (int o = 1; o < C::NOUT; o++)
                if (sl.idx == o) z = zo[o];
            if (C::SCALE) {
                sg[s] = sigmoid2(z);
                pv[s] = fma2s(sg[s], sl.span, f2s(sl.lo));
            } else {
                sg[s] = f2s(0.f);
                pv[s] = z;
            }
        } else {
            sg[s] = f2s(0.f);
            pv[s] = f2s(sS[s]);
        }
    }
}

// Shared prologue of the step and eval kernels: weight image + per-step scalars -> smem.
template <class C>
__device__ __forceinline__ void load_weights_and_scalars(const float* theta, const int* wsrc, const PSlot* slot,
                                                         const float* bscal, int use_bn, float* sW, float* sS)
{
    using PM = typename C::PM;
    for (int i = threadIdx.x; i < C::NW; i += blockDim.x) {
        int s = wsrc[i];
        sW[i] = s >= 0 ? theta[s] : 0.f;
    }
    if (threadIdx.x < C::NPS) {
        const PSlot sl = slot[threadIdx.x];
        float v = 0.f;
        if (sl.role == ROLE_GLOBAL) {
            // scale_single_param, GenericHybridModel.jl:348-352 (accurate expf: once per CTA)
            float raw = theta[wsrc[C::NW + sl.idx]];
            v = sl.lo + sl.span * (1.f / (1.f + expf(-raw)));
        } else if (sl.role == ROLE_FIXED) {
            v = sl.fixedv;
        }
        sS[threadIdx.x] = v;
    }
    if (threadIdx.x < MAXT) sS[32 + threadIdx.x] = bscal ? bscal[threadIdx.x] : 0.f;
    if (threadIdx.x < 2 * C::P)
        sS[40 + threadIdx.x] = use_bn ? bscal[BS_BN + threadIdx.x] : ((threadIdx.x & 1) ? 1.f : 0.f);
    __syncthreads();
    if (threadIdx.x == 0) {
        PmScal s;
#pragma unroll
        for (int i = 0; i < 8; i++) s.s[i] = 0.f;
        PM::prep(sS, s);
#pragma unroll
        for (int i = 0; i < 8; i++) sS[16 + i] = s.s[i];
    }
    __syncthreads();
}

template <class C>
__global__ void __launch_bounds__(256, 1) k_step(const StepArgs a)
{
    constexpr ShapeDims D = C::D;
    constexpr int P = C::P, NH = C::NH, H = C::H, NOUT = C::NOUT, T = C::T, F = C::F, NPS = C::NPS;
    constexpr int RS = ROWSTRIDE;
    using PM = typename C::PM;

    extern __shared__ float4 smem4[];
    float* sW = reinterpret_cast<float*>(smem4);
    float* sS = sW + rup4(C::NW);  // scalars: [0..MAXPS) uniform slot values, [16..16+8) PmScal, [32..32+MAXT) c_t, [40..) BN
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    float* stage = sS + 64 + warp * C::STAGE_FLOATS;

    // ---- prologue part 1 (independent of the previous step's update): fetch my samples
    const int gw = blockIdx.x * nwarps + warp;   // global warp id
    const int GW = gridDim.x * nwarps;
    const int nchunks = (a.B + CHUNK - 1) / CHUNK;
    int chunk = gw;
    float4 r0[C::R4 / 4], r1[C::R4 / 4];
    bool v0 = false, v1 = false;
    auto fetch = [&](int ch) {
        int s0 = ch * CHUNK + 2 * lane, s1 = s0 + 1;
        v0 = ch < nchunks && s0 < a.B;
        v1 = ch < nchunks && s1 < a.B;
        long long i0 = 0, i1 = 0;
        if (a.idx) {
            if (v0) i0 = a.idx[s0];
            if (v1) i1 = a.idx[s1];
        } else {
            i0 = a.rec_base + s0;
            i1 = a.rec_base + s1;
        }
#pragma unroll
        for (int q = 0; q < C::R4 / 4; q++) {
            r0[q] = v0 ? __ldg(a.rec + i0 * (C::R4 / 4) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
            r1[q] = v1 ? __ldg(a.rec + i1 * (C::R4 / 4) + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    fetch(chunk);

    // constant rows of my staging tile: the "1" feature of every augmented input and the zero padding
    {
#pragma unroll
        for (int l = 1; l <= NH + 1; l++) {
            const int din = D.din(l), ka = D.ka(l), gA = D.gA(l);
#pragma unroll
            for (int k = din; k < ka; k++) {
                float v = (k == din) ? 1.f : 0.f;
                *reinterpret_cast<float2*>(stage + grow(gA, k) * RS + 2 * lane) = f2s(v);
            }
        }
        // padding rows of the output-layer delta group
#pragma unroll
        for (int o = NOUT; o < rup4(NOUT); o++)
            *reinterpret_cast<float2*>(stage + grow(D.gD(NH + 1), o) * RS + 2 * lane) = f2s(0.f);
    }

    // ---- wait for the previous update kernel (PDL), then pull weights and scalars
    pdl_wait();
    load_weights_and_scalars<C>(a.theta, a.wsrc, a.slot, a.bscal, a.use_bn, sW, sS);
    pdl_launch_dependents();

    StepCtx<C> cx;
#pragma unroll
    for (int i = 0; i < 8; i++) cx.pms.s[i] = sS[16 + i];
#pragma unroll
    for (int i = 0; i < MAXPS; i++) cx.slot_uniform[i] = i < NPS ? (a.slot[i].role != ROLE_NEURAL) : true;

    // lane-owned dW tiles
    float2 acc[C::NBI][16];
#pragma unroll
    for (int i = 0; i < C::NBI; i++)
#pragma unroll
        for (int e = 0; e < 16; e++) acc[i][e] = f2s(0.f);
    // per-lane statistics
    float2 st_loss[T];
    float2 st_gphi[NPS];
#pragma unroll
    for (int t = 0; t < T; t++) st_loss[t] = f2s(0.f);
#pragma unroll
    for (int s = 0; s < NPS; s++) st_gphi[s] = f2s(0.f);

    // my dW tile coordinates (rows of the delta / activation groups)
    int rowD[C::NBI], rowA[C::NBI];
#pragma unroll
    for (int i = 0; i < C::NBI; i++) {
        int b = lane + 32 * i;
        rowD[i] = 0;
        rowA[i] = 0;
#pragma unroll
        for (int l = 1; l <= NH + 1; l++) {
            const int b0 = D.blk0(l), nk = D.nk(l), nb = D.nj(l) * nk;
            if (b >= b0 && b < b0 + nb) {
                int jb = (b - b0) / nk, kb = (b - b0) % nk;
                rowD[i] = 5 * (D.gD(l) + jb);
                rowA[i] = 5 * (D.gA(l) + kb);
            }
        }
    }

    for (; chunk < nchunks; chunk += GW) {
        // ================= per-sample phase (2 samples per lane) =================
        float2 x[P], f[F > 0 ? F : 1], y[T];
        {
            const float* p0 = reinterpret_cast<const float*>(r0);
            const float* p1 = reinterpret_cast<const float*>(r1);
#pragma unroll
            for (int k = 0; k < P; k++) {
                float2 raw = f2(p0[k], p1[k]);
                // input BatchNorm(affine=false): (x - mu) * rstd, batch statistics precomputed per batch
                x[k] = mul2s(sub2(raw, f2s(sS[40 + 2 * k])), sS[40 + 2 * k + 1]);
            }
#pragma unroll
            for (int k = 0; k < F; k++) f[k] = f2(p0[P + k], p1[P + k]);
#pragma unroll
            for (int k = 0; k < T; k++) y[k] = f2(p0[P + F + k], p1[P + F + k]);
        }
        const bool w0 = v0, w1 = v1;
        // prefetch the next chunk's records while this one is being processed
        fetch(chunk + GW);

        // stage a_0 = x
#pragma unroll
        for (int k = 0; k < P; k++) *reinterpret_cast<float2*>(stage + grow(D.gA(1), k) * RS + 2 * lane) = x[k];

        // ---- forward chain (stages a_1..a_NH for the weight-gradient phase)
        float2 h[H], zo[NOUT];
        chain_forward<C, true>(sW, stage, lane, x, h, zo);

        // ---- process parameters (GenericHybridModel.jl:377-414) and physics (:425)
        float2 pv[NPS], sg[NPS];
        resolve_params<C>(a.slot, sS, zo, pv, sg);
        float2 yh[T], sv[4], gy[T], gp[NPS];
        PM::fwd(pv, f, a.pmc, cx, yh, sv);
        // masked residual: valid_mask = !isnan(y) (train.jl:221-232); seeds dL/dyhat
#pragma unroll
        for (int t = 0; t < T; t++) {
            bool m0 = w0 && (y[t].x == y[t].x), m1 = w1 && (y[t].y == y[t].y);
            float2 r = f2(m0 ? yh[t].x - y[t].x : 0.f, m1 ? yh[t].y - y[t].y : 0.f);
            const float c = sS[32 + t];
            if (a.loss_kind[t] == LOSS_MAE) {
                st_loss[t] = add2(st_loss[t], f2(fabsf(r.x), fabsf(r.y)));
                gy[t] = f2(r.x > 0.f ? c : (r.x < 0.f ? -c : 0.f), r.y > 0.f ? c : (r.y < 0.f ? -c : 0.f));
            } else {
                st_loss[t] = fma2(r, r, st_loss[t]);
                gy[t] = mul2s(r, 2.f * c);
            }
        }
        PM::bwd(pv, f, a.pmc, cx, yh, sv, gy, gp);

        // ---- delta at the linear output layer; phi statistics
        float2 dz[NOUT];
#pragma unroll
        for (int o = 0; o < NOUT; o++) dz[o] = f2s(0.f);
#pragma unroll
        for (int s = 0; s < NPS; s++) {
            const PSlot sl = a.slot[s];
            if (sl.role == ROLE_NEURAL) {
                float2 g = gp[s];
                if (C::SCALE) g = mul2(g, mul2s(mul2(sg[s], sub2(f2s(1.f), sg[s])), sl.span));
#pragma unroll
                for (int o = 0; o < NOUT; o++)
                    if (sl.idx == o) dz[o] = add2(dz[o], g);
            } else if (sl.role == ROLE_GLOBAL) {
                st_gphi[s] = add2(st_gphi[s], gp[s]);
            }
        }
#pragma unroll
        for (int o = 0; o < NOUT; o++)
            *reinterpret_cast<float2*>(stage + grow(D.gD(NH + 1), o) * RS + 2 * lane) = dz[o];

        // ---- backward data pass: delta_l for l = NH .. 1 (h[] still holds a_NH)
        float2 d[H];
        {
            // through the output layer: d_k = sum_o Wo[o][k] dz_o, times act'(a_NH)
#pragma unroll
            for (int k = 0; k < H; k++) d[k] = f2s(0.f);
#pragma unroll
            for (int o = 0; o < NOUT; o++) {
                const float4* w4 = reinterpret_cast<const float4*>(sW + D.off_wo() + o * H);
#pragma unroll
                for (int k = 0; k < H; k += 4) {
                    float4 w = w4[k >> 2];
                    d[k] = fma2s(dz[o], w.x, d[k]);
                    d[k + 1] = fma2s(dz[o], w.y, d[k + 1]);
                    d[k + 2] = fma2s(dz[o], w.z, d[k + 2]);
                    d[k + 3] = fma2s(dz[o], w.w, d[k + 3]);
                }
            }
        }
#pragma unroll
        for (int l = NH; l >= 1; l--) {
            // multiply by act'(a_l); a_l (and sigma for swish) come back from the staging tile
#pragma unroll
            for (int k = 0; k < H; k++) {
                float2 al = (l == NH) ? h[k] : *reinterpret_cast<const float2*>(stage + grow(D.gA(l + 1), k) * RS + 2 * lane);
                float2 aux = f2s(0.f);
                if (C::ACT == ACT_SWISH)
                    aux = *reinterpret_cast<const float2*>(stage + (C::AUXROW0 + (l - 1) * H + k) * RS + 2 * lane);
                d[k] = mul2(d[k], act_bwd2<C::ACT>(al, aux));
                *reinterpret_cast<float2*>(stage + grow(D.gD(l), k) * RS + 2 * lane) = d[k];
            }
            if (l > 1) {
                // delta_{l-1}[k] = sum_j W_l[j][k] delta_l[j]  (j-major copy of W_l, vector over k)
                float2 dn[H];
#pragma unroll
                for (int k = 0; k < H; k++) dn[k] = f2s(0.f);
#pragma unroll
                for (int j = 0; j < H; j++) {
                    const float4* w4 = reinterpret_cast<const float4*>(sW + D.off_wb(l) + j * H);
#pragma unroll
                    for (int k = 0; k < H; k += 4) {
                        float4 w = w4[k >> 2];
                        dn[k] = fma2s(d[j], w.x, dn[k]);
                        dn[k + 1] = fma2s(d[j], w.y, dn[k + 1]);
                        dn[k + 2] = fma2s(d[j], w.z, dn[k + 2]);
                        dn[k + 3] = fma2s(d[j], w.w, dn[k + 3]);
                    }
                }
#pragma unroll
                for (int k = 0; k < H; k++) d[k] = dn[k];
            }
        }
        __syncwarp();

        // ================= weight-gradient phase (lane = one 4x4 tile) =================
#pragma unroll
        for (int i = 0; i < C::NBI; i++) {
            if (lane + 32 * i < C::NB) {
                const float* pd = stage + rowD[i] * RS;
                const float* pa = stage + rowA[i] * RS;
#pragma unroll 4
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
        __syncwarp();
    }

    // ================= CTA-level fixed-order reduction, no atomics =================
    __syncthreads();  // every warp is done with its staging tile; reuse it as [nwarps][NPART]
    float* red = sS + 64;
#pragma unroll
    for (int i = 0; i < C::NBI; i++) {
        int b = lane + 32 * i;
        if (b < C::NB) {
#pragma unroll
            for (int e = 0; e < 16; e++) red[warp * C::NPART + b * 16 + e] = acc[i][e].x + acc[i][e].y;
        }
    }
#pragma unroll
    for (int t = 0; t < MAXT; t++) {
        float v = t < T ? warp_sum(st_loss[t < T ? t : 0].x + st_loss[t < T ? t : 0].y) : 0.f;
        if (lane == 0) red[warp * C::NPART + C::D.npart_dw() + t] = v;
    }
#pragma unroll
    for (int s = 0; s < MAXPS; s++) {
        float v = s < NPS ? warp_sum(st_gphi[s < NPS ? s : 0].x + st_gphi[s < NPS ? s : 0].y) : 0.f;
        if (lane == 0) red[warp * C::NPART + C::D.npart_dw() + MAXT + s] = v;
    }
    __syncthreads();
    for (int p = threadIdx.x; p < C::NPART; p += blockDim.x) {
        float s = 0.f;
        for (int w = 0; w < nwarps; w++) s += red[w * C::NPART + p];
        a.partial[(size_t)blockIdx.x * a.npart + p] = s;
    }
}

}  // namespace eh
