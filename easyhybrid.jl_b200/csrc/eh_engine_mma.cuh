// eh_engine_mma.cuh -- compute engine 1: hidden-layer contractions on the tensor pipe.
//
// Why: measured on B200 (tools/ubench.cu) the FFMA2 engine runs at the broadcast-LDS.128 rate
// (2.26 cycles per instruction per SM, two FFMA2 per load), i.e. 2.3x below the FMA pipe and with the
// shared-memory writeback path saturated.  HMMA.1688.F32.TF32 sustains 481 MAC/clk/SM with every
// operand in registers, so this engine keeps the 16x16 hidden weights RESIDENT IN REGISTERS as
// m16n8k8 B-fragments for the whole step and feeds activations straight from the previous layer's
// accumulator registers:
//   * C-fragment -> A-fragment without data movement: lane (g, t) holds columns {2t, 2t+1} of an
//     n-tile; declaring logical k' = t <-> neuron 2t and k' = t+4 <-> neuron 2t+1 (and loading the
//     weight fragments with the same permutation) makes the accumulator registers valid A operands.
//   * fp32-level accuracy from TF32 inputs by 3xTF32: x = hi + lo (hi = x with the low 13 mantissa
//     bits cleared, lo = x - hi exactly); D += A_lo B_hi + A_hi B_lo + A_hi B_hi, fp32 accumulate.
//     Relative error ~2^-21 per product, well inside the 1e-5 parity budget (tests/test_gpu_parity.py).
//   * the weight gradient dW2 = delta2^T a1 needs samples along K: the two 16x16 tiles go through a
//     3 KB warp-private shared-memory transpose (conflict-free row stride 24) and 12 more MMAs,
//     accumulating in 8 registers per lane for the whole step.
//   * first layer (fan-in P <= 4) and the linear output layer are a few FFMA2 on the same
//     accumulator layout; process model, masked loss and seeds are evaluated per sample (the 4 lanes
//     of a row group compute them redundantly, statistics are taken from lane t = 0).
// A lane handles rows g and g+8 of a 16-sample tile; a warp takes two tiles (32 samples) per pass.
#pragma once
#include "eh_chunk.cuh"

namespace eh {

template <class C>
struct EngMma {
    using Cfg = C;
    static_assert(C::H == 16 && C::NH == 2, "the register-resident MMA engine covers two hidden layers of width <= 16");
    static_assert(C::P <= 4 && C::NOUT <= 2, "narrow first / output layers");
    static constexpr int ENGINE = 1;
    static constexpr int CHUNK = 32;                  // two 16-sample tiles per warp pass
    static constexpr int P = C::P, H = C::H, NOUT = C::NOUT, T = C::T, F = C::F, NPS = C::NPS;
    static constexpr int NT = H / 8;                  // n-tiles == k-steps of the hidden layer
    static constexpr int RS_T = H + 8;                // transpose tile row stride (floats), conflict-free
    static constexpr int STAGE_FLOATS = 2 * 16 * RS_T;
    static constexpr int MAX_WARPS = 14;              // 146 registers per thread; 148 x 14 warps cover 65 536 samples in one pass
    // partial vector: padded flat layout
    static constexpr int O_W1 = 0, O_B1 = H * P, O_W2 = O_B1 + H, O_B2 = O_W2 + H * H, O_WO = O_B2 + H,
                         O_BO = O_WO + NOUT * H, OFF_STATS = O_BO + 4, O_GP = OFF_STATS + MAXT, NPART = O_GP + MAXPS;

    struct State {
        // weights of this step, fragment / lane layout
        float2 w1[NT][P > 0 ? P : 1], b1[NT], b2[NT], wo[NOUT][NT];
        float bo[NOUT];
        float wf[NT][NT][2];   // forward  B[k' (of a1)][n = neuron of layer 2] (split hi/lo at use: 2 ALU ops)
        float wb[NT][NT][2];   // backward B[k' (of delta2)][n = neuron of layer 1]
        // gradient accumulators of this step
        float dW2[NT][4];
        float2 dW1[NT][P > 0 ? P : 1], db1[NT], db2[NT], dWo[NOUT][NT];
        float dbo[NOUT], loss[T], gphi[NPS];
        // prefetched records: [tile][row half]
        float4 r[2][2][C::R4 / 4];
        bool valid[2][2];
    };

    __device__ __forceinline__ static void init_warp(State&, float*, int) {}
    __device__ __forceinline__ static void after_reduce(State&, float*, int) {}

    // weight image (eh_layout.h offsets) -> fragments; zero the accumulators
    __device__ __forceinline__ static void step_begin(State& s, const float* sW, int lane)
    {
        constexpr ShapeDims D = C::D;
        const int g = lane >> 2, t = lane & 3;
#pragma unroll
        for (int nb = 0; nb < NT; nb++) {
            const int j = 8 * nb + 2 * t;  // my neuron pair (j, j+1) of this n-tile
#pragma unroll
            for (int k = 0; k < P; k++) s.w1[nb][k] = *reinterpret_cast<const float2*>(sW + D.off_w1f() + k * H + j);
            s.b1[nb] = *reinterpret_cast<const float2*>(sW + D.off_b1() + j);
            s.b2[nb] = *reinterpret_cast<const float2*>(sW + D.off_b(2) + j);
#pragma unroll
            for (int o = 0; o < NOUT; o++) s.wo[o][nb] = *reinterpret_cast<const float2*>(sW + D.off_wo() + o * H + j);
        }
#pragma unroll
        for (int o = 0; o < NOUT; o++) s.bo[o] = sW[D.off_bo() + o];
        // W2 fragments from the k-major image Wf2[k * H + j] = W2[j][k]
        const float* wf = sW + D.off_wf(2);
#pragma unroll
        for (int kb = 0; kb < NT; kb++)
#pragma unroll
            for (int nb = 0; nb < NT; nb++) {
                // forward: B[k' = t (+4)][n = g]  with k' <-> neuron 8kb + 2t (+1), n <-> neuron 8nb + g
                s.wf[kb][nb][0] = wf[(8 * kb + 2 * t) * H + 8 * nb + g];
                s.wf[kb][nb][1] = wf[(8 * kb + 2 * t + 1) * H + 8 * nb + g];
                // backward: B[j' = t (+4)][n = g]  = W2[j = 8kb + 2t (+1)][k = 8nb + g]
                s.wb[kb][nb][0] = wf[(8 * nb + g) * H + 8 * kb + 2 * t];
                s.wb[kb][nb][1] = wf[(8 * nb + g) * H + 8 * kb + 2 * t + 1];
            }
#pragma unroll
        for (int nb = 0; nb < NT; nb++) {
#pragma unroll
            for (int c = 0; c < 4; c++) s.dW2[nb][c] = 0.f;
#pragma unroll
            for (int k = 0; k < P; k++) s.dW1[nb][k] = f2s(0.f);
            s.db1[nb] = f2s(0.f);
            s.db2[nb] = f2s(0.f);
#pragma unroll
            for (int o = 0; o < NOUT; o++) s.dWo[o][nb] = f2s(0.f);
        }
#pragma unroll
        for (int o = 0; o < NOUT; o++) s.dbo[o] = 0.f;
#pragma unroll
        for (int i = 0; i < T; i++) s.loss[i] = 0.f;
#pragma unroll
        for (int i = 0; i < NPS; i++) s.gphi[i] = 0.f;
    }

    __device__ __forceinline__ static void fetch(State& s, const FetchArgs& fa, int chunk, int lane)
    {
        const int g = lane >> 2;
#pragma unroll
        for (int tile = 0; tile < 2; tile++)
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int smp = chunk * CHUNK + tile * 16 + g + 8 * h;
                const bool v = chunk < fa.nchunks && smp < fa.B;
                s.valid[tile][h] = v;
#pragma unroll
                for (int q = 0; q < C::R4 / 4; q++)
                    s.r[tile][h][q] = v ? fetch_rec4<C::R4 / 4>(fa, smp, q) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
    }
    // 3xTF32 product with a full-precision weight fragment kept in registers
    __device__ __forceinline__ static void mma_w(float (&d)[4], const unsigned (&ah)[4], const unsigned (&al)[4], const float (&w)[2])
    {
        unsigned bh[2], bl[2];
        split_tf32(w[0], bh[0], bl[0]);
        split_tf32(w[1], bh[1], bl[1]);
        mma_3xtf32(d, ah, al, bh, bl);
    }

    // accumulator registers of n-tile kb (pairs per row half) -> A fragment (k' permutation, see header)
    __device__ __forceinline__ static void make_a(const float2 (&v)[2], unsigned (&ah)[4], unsigned (&al)[4])
    {
        split_tf32(v[0].x, ah[0], al[0]);  // (row g,   k' = t)   = neuron 2t
        split_tf32(v[1].x, ah[1], al[1]);  // (row g+8, k' = t)
        split_tf32(v[0].y, ah[2], al[2]);  // (row g,   k' = t+4) = neuron 2t+1
        split_tf32(v[1].y, ah[3], al[3]);  // (row g+8, k' = t+4)
    }

    // process the prefetched chunk straight out of the record registers, then prefetch chunk `next`
    // (no double buffering: at 14 warps per SM the other warps hide that latency, and it keeps the
    // kernel under 146 registers)
    __device__ __forceinline__ static void chunk(State& s, const FetchArgs& fa, int next, const float* sW, const float* sS,
                                                 float* stage, int lane, const PSlot* slot, const int* loss_kind,
                                                 const PmCtx& cx)
    {
        using PM = typename C::PM;
        const int g = lane >> 2, t = lane & 3;
        float* Ta = stage;               // a1 of this tile      [16][RS_T]
        float* Td = stage + 16 * RS_T;   // delta2 of this tile  [16][RS_T]
#pragma unroll
        for (int tile = 0; tile < 2; tile++) {
            float x[2][P > 0 ? P : 1], f[2][F > 0 ? F : 1], y[2][T];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const float* rec = reinterpret_cast<const float*>(s.r[tile][h]);
#pragma unroll
                for (int k = 0; k < P; k++) x[h][k] = (rec[k] - sS[SS_BN + 2 * k]) * sS[SS_BN + 2 * k + 1];
#pragma unroll
                for (int k = 0; k < F; k++) f[h][k] = rec[P + k];
#pragma unroll
                for (int k = 0; k < T; k++) y[h][k] = rec[P + F + k];
            }
            // ---- layer 1 (FFMA2 on neuron pairs, accumulator layout) ----
            float2 a1[NT][2], x1[NT][2];
#pragma unroll
            for (int nb = 0; nb < NT; nb++)
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    float2 z = s.b1[nb];
#pragma unroll
                    for (int k = 0; k < P; k++) z = fma2s(s.w1[nb][k], x[h][k], z);
                    x1[nb][h] = f2s(0.f);
                    a1[nb][h] = act_fwd2<C::ACT>(z, x1[nb][h]);
                }
            // ---- layer 2 forward on the tensor pipe ----
            float d2[NT][4];
#pragma unroll
            for (int nb = 0; nb < NT; nb++) { d2[nb][0] = d2[nb][2] = s.b2[nb].x; d2[nb][1] = d2[nb][3] = s.b2[nb].y; }
#pragma unroll
            for (int kb = 0; kb < NT; kb++) {
                unsigned ah[4], al[4];
                make_a(a1[kb], ah, al);
#pragma unroll
                for (int nb = 0; nb < NT; nb++) mma_w(d2[nb], ah, al, s.wf[kb][nb]);
            }
            float2 a2[NT][2], x2[NT][2];
#pragma unroll
            for (int nb = 0; nb < NT; nb++)
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    x2[nb][h] = f2s(0.f);
                    a2[nb][h] = act_fwd2<C::ACT>(f2(d2[nb][2 * h], d2[nb][2 * h + 1]), x2[nb][h]);
                }
            // ---- linear output layer: my 2*NT neurons, then the 4 lanes of the row group ----
            float zo[2][NOUT];
#pragma unroll
            for (int h = 0; h < 2; h++)
#pragma unroll
                for (int o = 0; o < NOUT; o++) {
                    float2 p = f2s(0.f);
#pragma unroll
                    for (int nb = 0; nb < NT; nb++) p = fma2(s.wo[o][nb], a2[nb][h], p);
                    float v = p.x + p.y;
                    v += __shfl_xor_sync(0xffffffffu, v, 1);
                    v += __shfl_xor_sync(0xffffffffu, v, 2);
                    zo[h][o] = v + s.bo[o];
                }
            // ---- process parameters, physics, masked residual, seeds (per sample) ----
            cx.wait_phi();
            float dz[2][NOUT];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                float pv[NPS], sg[NPS], yh[T], sv[PM::NSV], gy[T], gp[NPS];
                resolve_params<C>(slot, sS, zo[h], pv, sg, cx);
                PM::fwd(pv, f[h], cx, yh, sv);
#pragma unroll
                for (int i = 0; i < T; i++) {
                    const bool m = s.valid[tile][h] && (y[h][i] == y[h][i]);
                    const float r = m ? yh[i] - y[h][i] : 0.f;
                    const float c = sS[SS_C + i];
                    if (loss_kind[i] == LOSS_MAE) {
                        s.loss[i] += fabsf(r);
                        gy[i] = r > 0.f ? c : (r < 0.f ? -c : 0.f);
                    } else {
                        s.loss[i] = fmaf(r, r, s.loss[i]);
                        gy[i] = 2.f * c * r;
                    }
                }
                PM::bwd(pv, f[h], cx, yh, sv, gy, gp);
#pragma unroll
                for (int o = 0; o < NOUT; o++) dz[h][o] = 0.f;
#pragma unroll
                for (int q = 0; q < NPS; q++) {
                    const PSlot sl = slot[q];
                    if (sl.role == ROLE_NEURAL) {
                        float gg = gp[q];
                        if (C::SCALE) gg *= sl.span * sg[q] * (1.f - sg[q]);
#pragma unroll
                        for (int o = 0; o < NOUT; o++)
                            if (sl.idx == o) dz[h][o] += gg;
                    } else if (sl.role == ROLE_GLOBAL) {
                        s.gphi[q] += gp[q];
                    }
                }
#pragma unroll
                for (int o = 0; o < NOUT; o++) s.dbo[o] += dz[h][o];
            }
            // ---- backward: output layer, delta2 ----
            float2 de2[NT][2];
#pragma unroll
            for (int nb = 0; nb < NT; nb++)
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    float2 v = f2s(0.f);
#pragma unroll
                    for (int o = 0; o < NOUT; o++) {
                        v = fma2s(s.wo[o][nb], dz[h][o], v);
                        s.dWo[o][nb] = fma2s(a2[nb][h], dz[h][o], s.dWo[o][nb]);
                    }
                    de2[nb][h] = mul2(v, act_bwd2<C::ACT>(a2[nb][h], x2[nb][h]));
                    s.db2[nb] = add2(s.db2[nb], de2[nb][h]);
                }
            // ---- layer 2 backward data pass on the tensor pipe ----
            float dp[NT][4];
#pragma unroll
            for (int nb = 0; nb < NT; nb++)
#pragma unroll
                for (int c = 0; c < 4; c++) dp[nb][c] = 0.f;
#pragma unroll
            for (int jb = 0; jb < NT; jb++) {
                unsigned ah[4], al[4];
                make_a(de2[jb], ah, al);
#pragma unroll
                for (int nb = 0; nb < NT; nb++) mma_w(dp[nb], ah, al, s.wb[jb][nb]);
            }
            // ---- delta1, first-layer gradients ----
#pragma unroll
            for (int nb = 0; nb < NT; nb++)
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    float2 d1 = mul2(f2(dp[nb][2 * h], dp[nb][2 * h + 1]), act_bwd2<C::ACT>(a1[nb][h], x1[nb][h]));
                    s.db1[nb] = add2(s.db1[nb], d1);
#pragma unroll
                    for (int k = 0; k < P; k++) s.dW1[nb][k] = fma2s(d1, x[h][k], s.dW1[nb][k]);
                }
            // ---- dW2 += delta2^T a1: samples along K through a warp-private transpose ----
#pragma unroll
            for (int nb = 0; nb < NT; nb++)
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    *reinterpret_cast<float2*>(Ta + (g + 8 * h) * RS_T + 8 * nb + 2 * t) = a1[nb][h];
                    *reinterpret_cast<float2*>(Td + (g + 8 * h) * RS_T + 8 * nb + 2 * t) = de2[nb][h];
                }
            __syncwarp();
#pragma unroll
            for (int sb = 0; sb < 2; sb++) {
                unsigned ah[4], al[4];
                // A[m = neuron j][k = sample]: rows g, g+8; samples 8sb + t, 8sb + t + 4
                split_tf32(Td[(8 * sb + t) * RS_T + g], ah[0], al[0]);
                split_tf32(Td[(8 * sb + t) * RS_T + g + 8], ah[1], al[1]);
                split_tf32(Td[(8 * sb + t + 4) * RS_T + g], ah[2], al[2]);
                split_tf32(Td[(8 * sb + t + 4) * RS_T + g + 8], ah[3], al[3]);
#pragma unroll
                for (int nb = 0; nb < NT; nb++) {
                    unsigned bh[2], bl[2];
                    // B[k = sample][n = input neuron 8nb + g]
                    split_tf32(Ta[(8 * sb + t) * RS_T + 8 * nb + g], bh[0], bl[0]);
                    split_tf32(Ta[(8 * sb + t + 4) * RS_T + 8 * nb + g], bh[1], bl[1]);
                    mma_3xtf32(s.dW2[nb], ah, al, bh, bl);
                }
            }
            __syncwarp();
        }
        fetch(s, fa, next, lane);
    }

    // sum over the 8 row groups (lanes with equal t)
    __device__ __forceinline__ static float gsum(float v)
    {
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        return v;
    }

    // the [nwarps][NPART] reduction scratch overlays the staging tiles: a CTA barrier is needed before reduce_prepare
    static constexpr bool SCRATCH_ALIASES_STAGE = true;
    // lane accumulators -> scratch[warp][NPART] (per warp, before the CTA barrier)
    __device__ __forceinline__ static void reduce_prepare(State& s, float* scratch)
    {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const int g = lane >> 2, t = lane & 3;
        float* sc = scratch + warp * NPART;
#pragma unroll
        for (int nb = 0; nb < NT; nb++) {
            // dW2 accumulator: rows j = g, g+8 ; cols k = 8nb + 2t, +1 ; flat cell j + k*H
#pragma unroll
            for (int c = 0; c < 4; c++) sc[O_W2 + (g + 8 * (c >> 1)) + (8 * nb + 2 * t + (c & 1)) * H] = s.dW2[nb][c];
            const int j = 8 * nb + 2 * t;
#pragma unroll
            for (int k = 0; k < P; k++) {
                float vx = gsum(s.dW1[nb][k].x), vy = gsum(s.dW1[nb][k].y);
                if (g == 0) { sc[O_W1 + j + k * H] = vx; sc[O_W1 + j + 1 + k * H] = vy; }
            }
            {
                float vx = gsum(s.db1[nb].x), vy = gsum(s.db1[nb].y);
                if (g == 0) { sc[O_B1 + j] = vx; sc[O_B1 + j + 1] = vy; }
                vx = gsum(s.db2[nb].x); vy = gsum(s.db2[nb].y);
                if (g == 0) { sc[O_B2 + j] = vx; sc[O_B2 + j + 1] = vy; }
            }
#pragma unroll
            for (int o = 0; o < NOUT; o++) {
                float vx = gsum(s.dWo[o][nb].x), vy = gsum(s.dWo[o][nb].y);
                if (g == 0) { sc[O_WO + o * H + j] = vx; sc[O_WO + o * H + j + 1] = vy; }
            }
        }
        // per-sample scalars were computed identically by the 4 lanes of a row group: take t = 0
#pragma unroll
        for (int o = 0; o < 4; o++) {
            float v = o < NOUT ? gsum(s.dbo[o < NOUT ? o : 0]) : 0.f;
            if (lane == 0) sc[O_BO + o] = v;
        }
#pragma unroll
        for (int i = 0; i < MAXT; i++) {
            float v = i < T ? gsum(s.loss[i < T ? i : 0]) : 0.f;
            if (lane == 0) sc[OFF_STATS + i] = v;
        }
#pragma unroll
        for (int i = 0; i < MAXPS; i++) {
            float v = i < NPS ? gsum(s.gphi[i < NPS ? i : 0]) : 0.f;
            if (lane == 0) sc[O_GP + i] = v;
        }
    }
    // after the barrier: element p summed over the first nw warps in fixed order (the partial vector is in flat order)
    __device__ __forceinline__ static float reduce_sum(const float* scratch, int nw, int q, int& p)
    {
        float sum = 0.f;
        for (int w = 0; w < nw; w++) sum += scratch[w * NPART + q];
        p = q;
        return sum;
    }
    __device__ __forceinline__ static float reduce_sum_at(const float* scratch, int nw, int p)
    {
        int dummy;
        return reduce_sum(scratch, nw, p, dummy);
    }
};

}  // namespace eh
