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

// a traced process model: straight-line SSA program (eh_pm_instr of the C ABI), at most PM_MAXLEN instructions
constexpr int PM_MAXLEN = 48;
struct PmProgData {
    int len, nt, nf, np;       // instructions, targets, forcings, parameters
    int out[MAXT];             // value id of each target
    short op[PM_MAXLEN], a[PM_MAXLEN], b[PM_MAXLEN];
    float imm[PM_MAXLEN];
};
enum : int {   // == eh_pm_op of easyhybrid_cuda.h
    POP_CONST = 0, POP_FORCING = 1, POP_PARAM = 2, POP_ADD = 10, POP_SUB = 11, POP_MUL = 12, POP_DIV = 13, POP_POW = 14,
    POP_MIN = 15, POP_MAX = 16, POP_NEG = 20, POP_EXP = 21, POP_LOG = 22, POP_SQRT = 23, POP_TANH = 24, POP_SIGMOID = 25,
    POP_ABS = 26, POP_SIN = 27, POP_COS = 28
};

// uniform-slot context handed to the functors
struct PmCtx {
    const float* pms;          // [MAXPS * PMS_PER_SLOT] per-slot derived scalars
    const float* c;            // process-model constants
    const PmProgData* prog;    // traced program (PmProgram only)
    int scale_rt;              // PmProgram only: scale_nn_outputs as a run-time flag (one compiled variant serves both)
    unsigned pass[3];          // PmProgram only: bit j of pass[l-1] = unit j of hidden layer l is a pass-through unit (identity
                               // activation) -- the tail of a chain shallower than the embedded depth (MultiNN, eh_plan.cu)
    unsigned uniform_mask;     // bit s set: slot s is GLOBAL / FIXED (same value for all samples)
    // persistent kernel only: shared-memory word that carries the number of the last optimiser step whose global-parameter
    // scalars (slot values, pms) are in place; the warp that owns the global parameters publishes it, consumers wait for
    // `phi_want` right before they read those scalars (NULL: nothing to wait for)
    const unsigned* phi_flag;
    unsigned phi_want;
    __device__ __forceinline__ bool uniform(int s) const { return (uniform_mask >> s) & 1u; }
    __device__ __forceinline__ void wait_phi() const
    {
        if (phi_flag) {
            const unsigned a = (unsigned)__cvta_generic_to_shared(phi_flag);
            unsigned v;
            do {
                asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
            } while (v != phi_want);
            __threadfence_block();
        }
    }
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

// 2^(-j/64), j = 0 .. 64, correctly rounded
static __constant__ double c_exp2_neg64[65] = {
    1.0, 0.9892280131939755, 0.9785720620877001, 0.9680308967461472,
    0.9576032806985737, 0.9472879907934828, 0.93708381705515, 0.9269895625416927,
    0.9170040432046712, 0.9071260877501994, 0.8973545375015536, 0.8876882462632606,
    0.8781260801866497, 0.8686669176368531, 0.859309649061239, 0.8500531768592617,
    0.8408964152537145, 0.8318382901633682, 0.8228777390769825, 0.8140137109286739,
    0.8052451659746271, 0.7965710756711335, 0.7879904225539432, 0.7795022001189185,
    0.7711054127039704, 0.7627990753722692, 0.7545822137967114, 0.7464538641456324,
    0.7384130729697497, 0.7304588970903235, 0.7225904034885233, 0.714806669195985,
    0.7071067811865476, 0.6994898362691556, 0.691954940981916, 0.6845012114872953,
    0.6771277734684463, 0.6698337620266515, 0.6626183215798707, 0.6554806057623822,
    0.6484197773255048, 0.6414350080393891, 0.6345254785958666, 0.6276903785123455,
    0.620928906036742, 0.614240268053435, 0.6076236799902345, 0.6010783657263515,
    0.5946035575013605, 0.5881984958251406, 0.5818624293887887, 0.5755946149764913,
    0.5693943173783458, 0.5632608093041209, 0.5571933712979462, 0.5511912916539204,
    0.5452538663326288, 0.5393803988785599, 0.5335702003384118, 0.5278225891802786,
    0.5221368912137069, 0.5165124395106142, 0.5109485743270583, 0.5054446430258502,
    0.5};

// log2 of a positive finite float in double precision (error < 1e-15): exponent + a 1/64 table step + a six-term log1p
// series -- about a dozen FP64 operations instead of the ~35 of the library log2.  It sits on the critical path of every
// optimiser step of the persistent kernel (the thread that owns Q10 computes it before the step-top barrier).
__device__ __forceinline__ double log2_pos_fast(float v)
{
    if (!(v > 1e-30f) || !(v < 1e30f)) return log2((double)v);   // zero / negative / denormal / huge: library semantics
    int e;
    float m = 2.f * frexpf(v, &e);                 // v = m 2^(e-1), m in [1, 2)
    const int j = __float2int_rn(__log2f(m) * 64.f);   // 0 .. 64
    const double u = fma((double)m, c_exp2_neg64[j], -1.0);   // |u| < 2^(1/128) - 1 + rounding of j = 0.0055
    double p = fma(u, -1.0 / 6.0, 0.2);
    p = fma(u, p, -0.25);
    p = fma(u, p, 1.0 / 3.0);
    p = fma(u, p, -0.5);
    p = fma(u, p, 1.0);                            // log1p(u) / u, next term u^6 / 7 < 4e-15 relative
    return (double)(e - 1) + (double)j * (1.0 / 64.0) + u * p * 1.4426950408889634;
}

// pm_prep_slot with log2_pos_fast (persistent kernel)
__device__ inline void pm_prep_slot_fast(int pm, int slot, float v, float* out4)
{
    out4[0] = out4[1] = out4[2] = out4[3] = 0.f;
    if (pm == PM_RBQ10 && slot == 1) {
        const double L = log2_pos_fast(v);
        const float Lh = (float)L;
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
    static constexpr bool DYNAMIC = false;
    static constexpr bool UNIT_ACT = false;
    static constexpr int NSV = 4;   // floats of forward state kept for the backward
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
    static constexpr bool DYNAMIC = false;
    static constexpr bool UNIT_ACT = false;
    static constexpr int NSV = 4;   // floats of forward state kept for the backward
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
    static constexpr bool DYNAMIC = false;
    static constexpr bool UNIT_ACT = false;
    static constexpr int NSV = 4;   // floats of forward state kept for the backward
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
    static constexpr bool DYNAMIC = false;
    static constexpr bool UNIT_ACT = false;
    static constexpr int NSV = 4;   // floats of forward state kept for the backward
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
    static constexpr bool DYNAMIC = false;
    static constexpr bool UNIT_ACT = false;
    static constexpr int NSV = 4;   // floats of forward state kept for the backward
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

// Any other process model: the host traces `mechanistic_model(; forcings..., params...)`
// (src/models/GenericHybridModel.jl:425) into a straight-line program; here it is interpreted per sample, and its
// pullback is the reverse sweep over the same program (what Zygote does with the Julia closure).  Slot s = parameter s
// of the descriptor, forcing k = forcing column k.  Compile-time maxima; the real counts come with the program.
struct PmProgram {
    static constexpr int ID = PM_PROGRAM, NPS = MAXPS, NF = 4, NT = 2;
    static constexpr bool DYNAMIC = true;
    static constexpr bool UNIT_ACT = false;   // per-unit activations: run-time compiled functors only (eh_jit.cu)
    static constexpr int NSV = PM_MAXLEN;   // the forward values of all instructions: the reverse sweep reuses them
    __device__ __forceinline__ static void eval(const float* p, const float* f, const PmProgData& pg, float* v)
    {
        for (int i = 0; i < pg.len; i++) {
            const int op = pg.op[i];
            const float x = v[pg.a[i] < i ? pg.a[i] : 0], y = v[pg.b[i] < i ? pg.b[i] : 0];
            float r;
            switch (op) {
            case POP_CONST: r = pg.imm[i]; break;
            case POP_FORCING: r = f[pg.a[i]]; break;
            case POP_PARAM: r = p[pg.a[i]]; break;
            case POP_ADD: r = x + y; break;
            case POP_SUB: r = x - y; break;
            case POP_MUL: r = x * y; break;
            case POP_DIV: r = x / y; break;
            case POP_POW: r = powf(x, y); break;
            case POP_MIN: r = x < y ? x : y; break;
            case POP_MAX: r = x > y ? x : y; break;
            case POP_NEG: r = -x; break;
            case POP_EXP: r = expf(x); break;
            case POP_LOG: r = logf(x); break;
            case POP_SQRT: r = sqrtf(x); break;
            case POP_TANH: r = tanhf(x); break;
            case POP_SIGMOID: r = 1.f / (1.f + expf(-x)); break;
            case POP_ABS: r = fabsf(x); break;
            case POP_SIN: r = sinf(x); break;
            default: r = cosf(x); break;
            }
            v[i] = r;
        }
    }
    __device__ __forceinline__ static void fwd(const float* p, const float* f, const PmCtx& cx, float* y, float* sv)
    {
        const PmProgData& pg = *cx.prog;
        float* v = sv;
        eval(p, f, pg, v);
        for (int t = 0; t < NT; t++) y[t] = t < pg.nt ? v[pg.out[t]] : 0.f;
    }
    __device__ __forceinline__ static void bwd(const float* p, const float* f, const PmCtx& cx, const float* y,
                                               const float* sv, const float* gy, float* gp)
    {
        const PmProgData& pg = *cx.prog;
        const float* v = sv;   // forward values, left there by fwd
        float g[PM_MAXLEN];
        for (int i = 0; i < pg.len; i++) g[i] = 0.f;
        for (int t = 0; t < NT; t++)
            if (t < pg.nt) g[pg.out[t]] += gy[t];
        for (int s = 0; s < NPS; s++) gp[s] = 0.f;
        for (int i = pg.len - 1; i >= 0; i--) {
            const float gi = g[i];
            const int op = pg.op[i], ia = pg.a[i], ib = pg.b[i];
            if (op == POP_CONST || op == POP_FORCING) continue;
            if (op == POP_PARAM) { gp[ia] += gi; continue; }
            const float x = v[ia], yv = op < POP_NEG ? v[ib] : 0.f;
            float ga = 0.f, gb = 0.f;
            switch (op) {
            case POP_ADD: ga = gi; gb = gi; break;
            case POP_SUB: ga = gi; gb = -gi; break;
            case POP_MUL: ga = gi * yv; gb = gi * x; break;
            case POP_DIV: ga = gi / yv; gb = -ga * v[i]; break;
            case POP_POW: ga = gi * yv * powf(x, yv - 1.f); gb = gi * v[i] * logf(x); break;
            case POP_MIN: if (x < yv) ga = gi; else gb = gi; break;
            case POP_MAX: if (x > yv) ga = gi; else gb = gi; break;
            case POP_NEG: ga = -gi; break;
            case POP_EXP: ga = gi * v[i]; break;
            case POP_LOG: ga = gi / x; break;
            case POP_SQRT: ga = gi / (2.f * v[i]); break;
            case POP_TANH: ga = gi * (1.f - v[i] * v[i]); break;
            case POP_SIGMOID: ga = gi * v[i] * (1.f - v[i]); break;
            case POP_ABS: ga = x < 0.f ? -gi : (x > 0.f ? gi : 0.f); break;
            case POP_SIN: ga = gi * cosf(x); break;
            default: ga = -gi * sinf(x); break;
            }
            g[ia] += ga;
            if (op < POP_NEG) g[ib] += gb;
        }
    }
};

}  // namespace eh
