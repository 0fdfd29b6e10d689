// eh_pm.cuh -- process-model plug-ins evaluated in registers (two samples per lane).
//
// Reference boundary: `y_pred = m.mechanistic_model(; all_kwargs...)`,
// src/models/GenericHybridModel.jl:425.  Each functor works on canonical argument
// slots: p[] are process parameters (already resolved from NEURAL / GLOBAL /
// FIXED roles), f[] are forcings, both in the canonical order of the form.
//   fwd : predictions y[T]
//   bwd : gp[slot] += sum_t gy[t] * dy_t/dp_slot         (hand-derived)
// `PmScal` carries per-step scalars prepared once per CTA (e.g. log2 Q10).
#pragma once
#include "eh_device.cuh"

namespace eh {

struct PmScal {
    float s[8];
};

// reco = rb * Q10^(0.1 (ta - tref))      README.md:148-151, test/test_split_data_train.jl:36-39
// slots: p0 = rb, p1 = Q10 ; f0 = ta ; const c0 = tref
struct PmRbQ10 {
    static constexpr int NPS = 2, NF = 1, NT = 1;
    // Julia evaluates Float32^Float32 through Float64, i.e. near correctly rounded.  Here
    // Q10^e = 2^(e*L), L = log2 Q10 split into hi+lo floats so the product e*L keeps ~2^-40.
    // If Q10 is per-sample (NEURAL role) the log2 is taken per sample instead.
    template <class Ctx>
    __device__ __forceinline__ static void fwd(const float2* p, const float2* f, const float* c, const Ctx& cx,
                                               float2* y, float2* sv)
    {
        float2 e = mul2s(sub2(f[0], f2s(c[0])), 0.1f);
        float2 pw;
        if (cx.slot_uniform[1]) {
            float Lh = cx.pms.s[0], Ll = cx.pms.s[1];
            float2 a = mul2s(e, Lh);
            float2 err = fma2s(e, Lh, f2(-a.x, -a.y));  // exact residual of the product
            float2 lo = fma2s(e, Ll, err);
            float2 base = ex2_2(a);
            // 2^(a+lo) = 2^a (1 + lo ln2)
            pw = fma2(base, mul2s(lo, 0.6931471806f), base);
        } else {
            pw = f2(exp2f(e.x * log2f(p[1].x)), exp2f(e.y * log2f(p[1].y)));
        }
        y[0] = mul2(p[0], pw);
        sv[0] = pw;
        sv[1] = e;
    }
    template <class Ctx>
    __device__ __forceinline__ static void bwd(const float2* p, const float2* f, const float* c, const Ctx& cx,
                                               const float2* y, const float2* sv, const float2* gy, float2* gp)
    {
        // dy/drb = Q10^e ; dy/dQ10 = rb e Q10^(e-1) = y e / Q10
        gp[0] = mul2(gy[0], sv[0]);
        float2 rq = cx.slot_uniform[1] ? f2s(cx.pms.s[2]) : rcp_2(p[1]);
        gp[1] = mul2(mul2(gy[0], y[0]), mul2(sv[1], rq));
    }
    // per-step scalars from the resolved uniform slot values (called by one thread)
    __device__ static void prep(const float* slotval, PmScal& s)
    {
        double L = log2((double)slotval[1]);
        float Lh = (float)L;
        s.s[0] = Lh;
        s.s[1] = (float)(L - (double)Lh);
        s.s[2] = 1.0f / slotval[1];
    }
};

// Resp = Resp0 * exp(k T)                 projects/ExpoHybrid/ExpoHybridEstim.jl:69-85
// slots: p0 = Resp0, p1 = k ; f0 = T
struct PmExpo {
    static constexpr int NPS = 2, NF = 1, NT = 1;
    template <class Ctx>
    __device__ __forceinline__ static void fwd(const float2* p, const float2* f, const float* c, const Ctx& cx,
                                               float2* y, float2* sv)
    {
        float2 a = mul2(p[1], f[0]);
        // exp(a) = 2^(a log2e): split log2e = hi + lo to keep the argument exact to ~2^-40
        const float Lh = 1.4426950216f, Ll = 1.9259630e-8f;
        float2 t = mul2s(a, Lh);
        float2 err = fma2s(a, Lh, f2(-t.x, -t.y));
        float2 lo = fma2s(a, Ll, err);
        float2 base = ex2_2(t);
        float2 ex = fma2(base, mul2s(lo, 0.6931471806f), base);
        y[0] = mul2(p[0], ex);
        sv[0] = ex;
    }
    template <class Ctx>
    __device__ __forceinline__ static void bwd(const float2* p, const float2* f, const float* c, const Ctx& cx,
                                               const float2* y, const float2* sv, const float2* gy, float2* gp)
    {
        gp[0] = mul2(gy[0], sv[0]);              // dy/dResp0 = exp(kT)
        gp[1] = mul2(mul2(gy[0], y[0]), f[0]);   // dy/dk = Resp0 T exp(kT) = y T
    }
    __device__ static void prep(const float*, PmScal&) {}
};

// y = a x + b                             src/models/LinearHM.jl:61-68, test/test_generic_hybrid_model.jl:10-12
// slots: p0 = a, p1 = b ; f0 = x
struct PmLinear {
    static constexpr int NPS = 2, NF = 1, NT = 1;
    template <class Ctx>
    __device__ __forceinline__ static void fwd(const float2* p, const float2* f, const float* c, const Ctx& cx,
                                               float2* y, float2* sv)
    {
        y[0] = fma2(p[0], f[0], p[1]);
    }
    template <class Ctx>
    __device__ __forceinline__ static void bwd(const float2* p, const float2* f, const float* c, const Ctx& cx,
                                               const float2* y, const float2* sv, const float2* gy, float2* gp)
    {
        gp[0] = mul2(gy[0], f[0]);
        gp[1] = gy[0];
    }
    __device__ static void prep(const float*, PmScal&) {}
};

// (var1 = a x + b, var2 = 2 a x + b)      test/test_compute_loss.jl:209-211
struct PmLinear2 {
    static constexpr int NPS = 2, NF = 1, NT = 2;
    template <class Ctx>
    __device__ __forceinline__ static void fwd(const float2* p, const float2* f, const float* c, const Ctx& cx,
                                               float2* y, float2* sv)
    {
        float2 ax = mul2(p[0], f[0]);
        y[0] = add2(ax, p[1]);
        y[1] = fma2s(ax, 2.f, p[1]);
    }
    template <class Ctx>
    __device__ __forceinline__ static void bwd(const float2* p, const float2* f, const float* c, const Ctx& cx,
                                               const float2* y, const float2* sv, const float2* gy, float2* gp)
    {
        float2 g = fma2s(gy[1], 2.f, gy[0]);
        gp[0] = mul2(g, f[0]);
        gp[1] = add2(gy[0], gy[1]);
    }
    __device__ static void prep(const float*, PmScal&) {}
};

}  // namespace eh
