// eh_pm.cuh -- process-model plug-ins evaluated in registers (one sample per lane).
//
// Reference boundary: `y_pred = m.mechanistic_model(; all_kwargs...)`,
// src/models/GenericHybridModel.jl:425.  Each functor works on canonical argument
// slots: p[] are process parameters (already resolved from NEURAL / GLOBAL /
// FIXED roles), f[] are forcings, both in the canonical order of the form.
//   fwd : predictions y[T]
//   bwd : gp[slot] = sum_t gy[t] * dy_t/dp_slot          (hand-derived)
// Per-step scalars of uniform (GLOBAL / FIXED) slots are prepared once per update by
// pm_prep_slot (4 floats per slot) and travel in the parameter block's tail.
#pragma once
#include "eh_device.cuh"

namespace eh {

constexpr int PMS_PER_SLOT = 4;

// uniform-slot context handed to the functors
struct PmCtx {
    const float* pms;          // [MAXPS * PMS_PER_SLOT] per-slot derived scalars
    const float* c;            // process-model constants
    unsigned uniform_mask;     // bit s set: slot s is GLOBAL / FIXED (same value for all samples)
    __device__ __forceinline__ bool uniform(int s) const { return (uniform_mask >> s) & 1u; }
};

// derived scalars of a uniform slot; called by the thread that owns the slot's update
__device__ inline void pm_prep_slot(int pm, int slot, float v, float* out4)
{
    out4[0] = out4[1] = out4[2] = out4[3] = 0.f;
    if (pm == PM_RBQ10 && slot == 1) {
        // Julia evaluates Float32^Float32 through Float64, i.e. near correctly rounded.  Here
        // Q10^e = 2^(e*L), L = log2 Q10 in double, split hi + lo so that e*L keeps ~2^-40.
        double L = log2((double)v);
        float Lh = (float)L;
        out4[0] = Lh;
        out4[1] = (float)(L - (double)Lh);
        out4[2] = 1.0f / v;
    }
}

// 2^(a*(Lh+Ll)) with the product carried in two floats
__device__ __forceinline__ float exp2_mul_hilo(float a, float Lh, float Ll)
{
    float t = a * Lh;
    float err = fmaf(a, Lh, -t);   // exact residual of the product
    float lo = fmaf(a, Ll, err);
    float base = ex2_approx(t);
    return fmaf(base, lo * 0.6931471806f, base);  // 2^(t+lo) = 2^t (1 + lo ln2)
}

// reco = rb * Q10^(0.1 (ta - tref))      README.md:148-151, test/test_split_data_train.jl:36-39
// slots: p0 = rb, p1 = Q10 ; f0 = ta ; const c0 = tref
struct PmRbQ10 {
    static constexpr int ID = PM_RBQ10, NPS = 2, NF = 1, NT = 1;
    __device__ __forceinline__ static void fwd(const float* p, const float* f, const PmCtx& cx, float* y, float* sv)
    {
        float e = 0.1f * (f[0] - cx.c[0]);
        float pw;
        if (cx.uniform(1)) pw = exp2_mul_hilo(e, cx.pms[1 * PMS_PER_SLOT + 0], cx.pms[1 * PMS_PER_SLOT + 1]);
        else pw = exp2f(e * log2f(p[1]));   // per-sample Q10 (NEURAL role)
        y[0] = p[0] * pw;
        sv[0] = pw;
        sv[1] = e;
    }
    __device__ __forceinline__ static void bwd(const float* p, const float* f, const PmCtx& cx, const float* y,
                                               const float* sv, const float* gy, float* gp)
    {
        // dy/drb = Q10^e ; dy/dQ10 = rb e Q10^(e-1) = y e / Q10
        gp[0] = gy[0] * sv[0];
        float rq = cx.uniform(1) ? cx.pms[1 * PMS_PER_SLOT + 2] : rcp_approx(p[1]);
        gp[1] = gy[0] * y[0] * sv[1] * rq;
    }
};

// Resp = Resp0 * exp(k T)                 projects/ExpoHybrid/ExpoHybridEstim.jl:69-85
// slots: p0 = Resp0, p1 = k ; f0 = T
struct PmExpo {
    static constexpr int ID = PM_EXPO, NPS = 2, NF = 1, NT = 1;
    __device__ __forceinline__ static void fwd(const float* p, const float* f, const PmCtx& cx, float* y, float* sv)
    {
        // exp(a) = 2^(a log2e), log2e = hi + lo
        float ex = exp2_mul_hilo(p[1] * f[0], 1.4426950216f, 1.9259630e-8f);
        y[0] = p[0] * ex;
        sv[0] = ex;
    }
    __device__ __forceinline__ static void bwd(const float* p, const float* f, const PmCtx& cx, const float* y,
                                               const float* sv, const float* gy, float* gp)
    {
        gp[0] = gy[0] * sv[0];          // dy/dResp0 = exp(kT)
        gp[1] = gy[0] * y[0] * f[0];    // dy/dk = Resp0 T exp(kT) = y T
    }
};

// y = a x + b                             src/models/LinearHM.jl:61-68, test/test_generic_hybrid_model.jl:10-12
// slots: p0 = a, p1 = b ; f0 = x
struct PmLinear {
    static constexpr int ID = PM_LINEAR, NPS = 2, NF = 1, NT = 1;
    __device__ __forceinline__ static void fwd(const float* p, const float* f, const PmCtx& cx, float* y, float* sv)
    {
        y[0] = fmaf(p[0], f[0], p[1]);
    }
    __device__ __forceinline__ static void bwd(const float* p, const float* f, const PmCtx& cx, const float* y,
                                               const float* sv, const float* gy, float* gp)
    {
        gp[0] = gy[0] * f[0];
        gp[1] = gy[0];
    }
};

// (var1 = a x + b, var2 = 2 a x + b)      test/test_compute_loss.jl:209-211
struct PmLinear2 {
    static constexpr int ID = PM_LINEAR2, NPS = 2, NF = 1, NT = 2;
    __device__ __forceinline__ static void fwd(const float* p, const float* f, const PmCtx& cx, float* y, float* sv)
    {
        float ax = p[0] * f[0];
        y[0] = ax + p[1];
        y[1] = fmaf(2.f, ax, p[1]);
    }
    __device__ __forceinline__ static void bwd(const float* p, const float* f, const PmCtx& cx, const float* y,
                                               const float* sv, const float* gy, float* gp)
    {
        gp[0] = fmaf(2.f, gy[1], gy[0]) * f[0];
        gp[1] = gy[0] + gy[1];
    }
};

// (var1 = Resp0 exp(k T), var2 = 2 Resp0 exp(k T)): the two-target form of BASELINE config 5 (the Expo model of
// projects/ExpoHybrid/ExpoHybridEstim.jl:69-85 with a second target that is twice the first, the same
// construction as test/test_compute_loss.jl:209-211 uses for the linear model)
struct PmExpo2 {
    static constexpr int ID = PM_EXPO2, NPS = 2, NF = 1, NT = 2;
    __device__ __forceinline__ static void fwd(const float* p, const float* f, const PmCtx& cx, float* y, float* sv)
    {
        float ex = exp2_mul_hilo(p[1] * f[0], 1.4426950216f, 1.9259630e-8f);
        y[0] = p[0] * ex;
        y[1] = 2.f * y[0];
        sv[0] = ex;
    }
    __device__ __forceinline__ static void bwd(const float* p, const float* f, const PmCtx& cx, const float* y,
                                               const float* sv, const float* gy, float* gp)
    {
        const float g = fmaf(2.f, gy[1], gy[0]);
        gp[0] = g * sv[0];
        gp[1] = g * y[0] * f[0];
    }
};

}  // namespace eh
