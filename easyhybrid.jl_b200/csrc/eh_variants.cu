// eh_variants.cu -- ahead-of-time instantiations of the fused step / eval kernels
// for the configurations named in BASELINE.json (RbQ10 quickstart, ExpoHybrid,
// Linear_Regression, the two-target linear test model), compiled for sm_100a only.
#include "eh_variants.h"
#include "eh_eval_kernel.cuh"

namespace eh {

template <class C>
static cudaError_t prepare_t(size_t step_smem, size_t eval_smem)
{
    cudaError_t e = cudaFuncSetAttribute(k_step<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)step_smem);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_eval<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)eval_smem);
}

template <class C>
static cudaError_t launch_step_t(const StepArgs& a, int grid, int nwarps, size_t smem, cudaStream_t st, bool pdl)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)(nwarps * 32));
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, k_step<C>, a);
}

template <class C>
static cudaError_t launch_eval_t(const EvalArgs& a, int grid, int nwarps, size_t smem, cudaStream_t st)
{
    k_eval<C><<<grid, nwarps * 32, smem, st>>>(a);
    return cudaGetLastError();
}

template <class C>
static Variant make_variant(int pm, const char* name)
{
    Variant v{};
    v.pm = pm; v.P = C::P; v.NH = C::NH; v.H = C::H; v.NOUT = C::NOUT; v.act = C::ACT; v.scale = C::SCALE ? 1 : 0;
    v.dims = C::D;
    v.F = C::F; v.T = C::T; v.NPS = C::NPS; v.R4 = C::R4; v.NB = C::NB; v.NW = C::NW; v.NPART = C::NPART;
    v.stage_floats = C::STAGE_FLOATS;
    v.name = name;
    v.prepare = prepare_t<C>;
    v.launch_step = launch_step_t<C>;
    v.launch_eval = launch_eval_t<C>;
    return v;
}

// (pm id, functor, P, NH, H, NOUT, ACT, SCALE)
#define EH_VARIANT_LIST(X)                                   \
    X(PM_RBQ10, PmRbQ10, 2, 2, 16, 1, ACT_TANH, true)        \
    X(PM_RBQ10, PmRbQ10, 2, 2, 16, 1, ACT_SWISH, true)       \
    X(PM_RBQ10, PmRbQ10, 2, 2, 16, 1, ACT_SIGMOID, true)     \
    X(PM_RBQ10, PmRbQ10, 2, 2, 16, 1, ACT_TANH, false)       \
    X(PM_RBQ10, PmRbQ10, 2, 2, 32, 1, ACT_TANH, true)        \
    X(PM_RBQ10, PmRbQ10, 2, 2, 32, 1, ACT_TANH, false)       \
    X(PM_EXPO, PmExpo, 1, 2, 16, 1, ACT_SIGMOID, false)      \
    X(PM_EXPO, PmExpo, 1, 2, 16, 1, ACT_SIGMOID, true)       \
    X(PM_EXPO, PmExpo, 1, 2, 16, 1, ACT_TANH, false)         \
    X(PM_LINEAR, PmLinear, 2, 2, 16, 1, ACT_RELU, false)     \
    X(PM_LINEAR, PmLinear, 2, 2, 16, 1, ACT_TANH, false)     \
    X(PM_LINEAR, PmLinear, 2, 2, 32, 1, ACT_TANH, false)     \
    X(PM_LINEAR2, PmLinear2, 2, 2, 16, 1, ACT_TANH, false)   \
    X(PM_LINEAR2, PmLinear2, 2, 2, 16, 1, ACT_RELU, false)

#define EH_MAKE(pmid, PMF, P, NH, H, NOUT, ACT, SCALE) \
    make_variant<StepCfg<P, NH, H, NOUT, ACT, SCALE, PMF>>(pmid, #PMF "/P" #P "/NH" #NH "/H" #H "/O" #NOUT "/" #ACT "/scale=" #SCALE),

static const Variant g_variants[] = {EH_VARIANT_LIST(EH_MAKE)};

int num_variants() { return (int)(sizeof(g_variants) / sizeof(g_variants[0])); }
const Variant* variant_at(int i) { return &g_variants[i]; }

const Variant* find_variant(int pm, int P, int NH, int H, int NOUT, int act, int scale)
{
    const Variant* best = nullptr;
    for (const Variant& v : g_variants) {
        if (v.pm != pm || v.P != P || v.NH != NH || v.NOUT != NOUT || v.act != act || v.scale != scale) continue;
        if (v.H < H) continue;
        if (!best || v.H < best->H) best = &v;
    }
    return best;
}

}  // namespace eh
