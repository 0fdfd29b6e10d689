// eh_device.cuh -- device-side helpers for sm_100a: packed f32x2 arithmetic
// (FFMA2 / FMUL2 / FADD2, Blackwell-only), MUFU-based transcendentals with
// fp32-level accuracy, warp reductions.  NVRTC-clean (no host headers).
#pragma once
#include "eh_layout.h"

namespace eh {

// ---- packed fp32x2 (one instruction per two samples; a 3-register FFMA issues at
// half rate on sm_100, FFMA2 restores the 128 FMA/clk/SM peak) ------------------
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ float2 f2s(float a) { return make_float2(a, a); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
// scalar-broadcast form: SASS `FFMA2 Rd, Ra.F32x2, Rs.F32, Rc.F32x2` (no MOV needed)
__device__ __forceinline__ float2 fma2s(float2 a, float s, float2 c) { return __ffma2_rn(a, make_float2(s, s), c); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 mul2s(float2 a, float s) { return __fmul2_rn(a, make_float2(s, s)); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return __ffma2_rn(b, make_float2(-1.f, -1.f), a); }

// ---- MUFU primitives -------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x)
{
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float2 ex2_2(float2 x) { return f2(ex2_approx(x.x), ex2_approx(x.y)); }
__device__ __forceinline__ float2 rcp_2(float2 x) { return f2(rcp_approx(x.x), rcp_approx(x.y)); }

constexpr float LOG2E = 1.4426950408889634f;

// sigmoid(z) = 1 / (1 + 2^(-z log2 e)); abs error ~2e-7 (ex2.approx 2^-22 rel, rcp 1 ulp)
__device__ __forceinline__ float2 sigmoid2(float2 z)
{
    float2 e = ex2_2(mul2s(z, -LOG2E));
    return rcp_2(add2(e, f2s(1.f)));
}
__device__ __forceinline__ float sigmoid1(float z) { return rcp_approx(1.f + ex2_approx(-LOG2E * z)); }

// tanh(z) = 1 - 2 / (1 + 2^(2 z log2 e)): absolute error ~1.5e-7 over the whole range (ex2.approx 2^-22 relative, rcp 1 ulp,
// one rounding of 1 + e), i.e. the rounding noise of the Float32 pre-activation itself.  With EH_TANH_POLY_BRANCH an odd
// minimax polynomial takes over below |z| = 0.3 (relative instead of absolute accuracy near 0) at ~12 more instructions per
// neuron pair -- measured: no change in the parity figures (tests/test_gpu_parity.py), 14 % of the chunk's instructions.
__device__ __forceinline__ float tanh1(float z)
{
    float e = ex2_approx(z * (2.f * LOG2E));
    float big = fmaf(-2.f, rcp_approx(e + 1.f), 1.f);
#ifdef EH_TANH_POLY_BRANCH
    float z2 = z * z;
    // tanh z ~ z (1 + z2 (-1/3 + z2 (2/15 + z2 (-17/315 + z2 62/2835)))), |z| < 0.3: rel err < 2e-8
    float p = fmaf(z2, 0.021869488f, -0.053968254f);
    p = fmaf(z2, p, 0.13333334f);
    p = fmaf(z2, p, -0.33333334f);
    float small = fmaf(z * z2, p, z);
    return fabsf(z) < 0.3f ? small : big;
#else
    return big;
#endif
}
__device__ __forceinline__ float2 tanh2(float2 z)
{
    float2 e = ex2_2(mul2s(z, 2.f * LOG2E));
    float2 r = rcp_2(add2(e, f2s(1.f)));
    float2 big = fma2s(r, -2.f, f2s(1.f));
#ifdef EH_TANH_POLY_BRANCH
    float2 z2 = mul2(z, z);
    float2 p = fma2s(z2, 0.021869488f, f2s(-0.053968254f));
    p = fma2(z2, p, f2s(0.13333334f));
    p = fma2(z2, p, f2s(-0.33333334f));
    float2 small = fma2(mul2(z, z2), p, z);
    return f2(fabsf(z.x) < 0.3f ? small.x : big.x, fabsf(z.y) < 0.3f ? small.y : big.y);
#else
    return big;
#endif
}

// hidden activation; `aux` receives sigma(z) for swish (needed by its derivative)
template <int ACT>
__device__ __forceinline__ float2 act_fwd2(float2 z, float2& aux)
{
    if (ACT == ACT_TANH) return tanh2(z);
    if (ACT == ACT_SIGMOID) return sigmoid2(z);
    if (ACT == ACT_RELU) return f2(fmaxf(z.x, 0.f), fmaxf(z.y, 0.f));
    if (ACT == ACT_SWISH) {
        aux = sigmoid2(z);
        return mul2(z, aux);
    }
    return z;
}
// derivative in terms of the stored output a (and aux = sigma for swish); SURVEY 10.4
template <int ACT>
__device__ __forceinline__ float2 act_bwd2(float2 a, float2 aux)
{
    if (ACT == ACT_TANH) return fma2(a, f2(-a.x, -a.y), f2s(1.f));
    if (ACT == ACT_SIGMOID) return mul2(a, sub2(f2s(1.f), a));
    if (ACT == ACT_RELU) return f2(a.x > 0.f ? 1.f : 0.f, a.y > 0.f ? 1.f : 0.f);
    if (ACT == ACT_SWISH) return fma2(a, sub2(f2s(1.f), aux), aux);
    return f2s(1.f);
}

// the same with the activation as a value: for call sites where it is a compile-time constant only after unrolling (the
// per-unit activations of run-time compiled models whose chains differ in activation); the switch folds away
__device__ __forceinline__ float2 act_fwd2_rt(int act, float2 z, float2& aux)
{
    switch (act) {
    case ACT_TANH: return act_fwd2<ACT_TANH>(z, aux);
    case ACT_SIGMOID: return act_fwd2<ACT_SIGMOID>(z, aux);
    case ACT_RELU: return act_fwd2<ACT_RELU>(z, aux);
    case ACT_SWISH: return act_fwd2<ACT_SWISH>(z, aux);
    default: return z;
    }
}
__device__ __forceinline__ float2 act_bwd2_rt(int act, float2 a, float2 aux)
{
    switch (act) {
    case ACT_TANH: return act_bwd2<ACT_TANH>(a, aux);
    case ACT_SIGMOID: return act_bwd2<ACT_SIGMOID>(a, aux);
    case ACT_RELU: return act_bwd2<ACT_RELU>(a, aux);
    case ACT_SWISH: return act_bwd2<ACT_SWISH>(a, aux);
    default: return f2s(1.f);
    }
}

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum_d(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---- tensor pipe, TF32 inputs with fp32-level accuracy (3xTF32) ------------------------------
__device__ __forceinline__ void mma_tf32(float (&d)[4], const unsigned (&a)[4], const unsigned (&b)[2])
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void split_tf32(float x, unsigned& hi, unsigned& lo)
{
    hi = __float_as_uint(x) & 0xffffe000u;         // the tensor pipe reads exactly these bits
    lo = __float_as_uint(x - __uint_as_float(hi));  // exact remainder (truncated to tf32 by the hardware)
}
// D += A B with fp32-level accuracy: small cross terms first
__device__ __forceinline__ void mma_3xtf32(float (&d)[4], const unsigned (&ah)[4], const unsigned (&al)[4],
                                           const unsigned (&bh)[2], const unsigned (&bl)[2])
{
    mma_tf32(d, al, bh);
    mma_tf32(d, ah, bl);
    mma_tf32(d, ah, bh);
}

// Programmatic dependent launch: wait for the producer grid / let the consumer start.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

}  // namespace eh
