// eh_variant_impl.cuh -- template glue that turns a StepCfg into a registry entry.
#pragma once
#include <cstdlib>
#include "eh_variants.h"
#include "eh_epoch_kernel.cuh"
#include "eh_eval_kernel.cuh"

namespace eh {

template <class E>
static cudaError_t prepare_t(size_t step_smem, size_t eval_smem)
{
    cudaError_t e = cudaFuncSetAttribute(k_step<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)step_smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_epoch<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)step_smem);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_eval<typename E::Cfg>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)eval_smem);
}

template <class E>
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
    return cudaLaunchKernelEx(&cfg, k_step<E>, a);
}

template <class C>
static cudaError_t launch_eval_t(const EvalArgs& a, int grid, int nwarps, size_t smem, cudaStream_t st)
{
    k_eval<C><<<grid, nwarps * 32, smem, st>>>(a);
    return cudaGetLastError();
}

// k_epoch: cooperative launch (the grid-wide exchange needs every CTA resident), one CTA per SM, no clusters.
// EH_NO_COOP=1 drops the co-residency guarantee (plain launch; the grid never exceeds the co-resident maximum anyway).
template <class E>
static cudaError_t launch_epoch_t(const EpochArgs& a, int grid, int nwarps, size_t smem, cudaStream_t st)
{
    if (getenv("EH_NO_COOP")) {
        k_epoch<E><<<grid, nwarps * 32, smem, st>>>(a);
        return cudaGetLastError();
    }
    void* args[] = {(void*)&a};
    return cudaLaunchCooperativeKernel((void*)k_epoch<E>, dim3((unsigned)grid), dim3((unsigned)(nwarps * 32)), args, smem, st);
}

// how many CTAs of this shape can be co-resident
template <class E>
static cudaError_t epoch_max_grid_t(int nwarps, size_t smem, int* max_ctas)
{
    int per_sm = 0, dev = 0, nsm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_epoch<E>, nwarps * 32, smem);
    if (e != cudaSuccess) return e;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    *max_ctas = per_sm * nsm;
    return cudaSuccess;
}

template <class E>
static Variant make_variant(const char* name)
{
    using C = typename E::Cfg;
    static_assert(E::NPART <= UPD_MAX_NPART, "partial vector longer than k_update's shared-memory buffer");
    Variant v{};
    v.pm = C::PM::ID; v.P = C::P; v.NH = C::NH; v.H = C::H; v.NOUT = C::NOUT; v.act = C::ACT; v.scale = C::SCALE ? 1 : 0;
    v.engine = E::ENGINE;
    v.chunk = E::CHUNK;
    v.dims = C::D;
    v.F = C::F; v.T = C::T; v.NPS = C::NPS; v.R4 = C::R4; v.NW = C::NW;
    v.NPART = E::NPART;
    v.off_stats = E::OFF_STATS;
    v.stage_floats = E::STAGE_FLOATS;
    v.max_warps = E::MAX_WARPS;
    v.wpc = 1; v.eng_bytes = 0;
    v.name = name;
    v.prepare = prepare_t<E>;
    v.launch_step = launch_step_t<E>;
    v.launch_eval = launch_eval_t<C>;
    v.launch_epoch = launch_epoch_t<E>;
    v.epoch_max_grid = epoch_max_grid_t<E>;
    v.epoch_func = (const void*)k_epoch<E>;
    return v;
}

// layout descriptor of a shape WITHOUT compiled-in kernels: what the planner and the launch geometry need to know about a
// generic variant that only exists once eh_jit.cu has compiled it (run-time specialisation, EH_FLAG_JIT)
template <class E>
static Variant make_shape(const char* name)
{
    using C = typename E::Cfg;
    Variant v{};
    v.pm = C::PM::ID; v.P = C::P; v.NH = C::NH; v.H = C::H; v.NOUT = C::NOUT; v.act = C::ACT; v.scale = C::SCALE ? 1 : 0;
    v.engine = E::ENGINE;
    v.chunk = E::CHUNK;
    v.dims = C::D;
    v.F = C::F; v.T = C::T; v.NPS = C::NPS; v.R4 = C::R4; v.NW = C::NW;
    v.NPART = E::NPART;
    v.off_stats = E::OFF_STATS;
    v.stage_floats = E::STAGE_FLOATS;
    v.max_warps = E::MAX_WARPS;
    v.wpc = 1; v.eng_bytes = 0;
    v.name = name;
    return v;
}

// tensor engine: persistent kernel only (single steps, eval and small batches stay on the FFMA2 variant of the same shape)
template <class E>
static cudaError_t prepare_epoch_only_t(size_t step_smem, size_t)
{
    return cudaFuncSetAttribute(k_epoch<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)step_smem);
}
template <class E>
static Variant make_variant_epoch_only(const char* name)
{
    using C = typename E::Cfg;
    Variant v{};
    v.pm = C::PM::ID; v.P = C::P; v.NH = C::NH; v.H = C::H; v.NOUT = C::NOUT; v.act = C::ACT; v.scale = C::SCALE ? 1 : 0;
    v.engine = E::ENGINE;
    v.chunk = E::CHUNK;
    v.dims = C::D;
    v.F = C::F; v.T = C::T; v.NPS = C::NPS; v.R4 = C::R4; v.NW = C::NW;
    v.NPART = E::NPART;
    v.off_stats = E::OFF_STATS;
    v.stage_floats = E::STAGE_FLOATS;
    v.max_warps = E::MAX_WARPS;
    v.wpc = E::WPC; v.eng_bytes = E::ENG_FLOATS * 4;
    v.name = name;
    v.prepare = prepare_epoch_only_t<E>;
    v.launch_epoch = launch_epoch_t<E>;
    v.epoch_max_grid = epoch_max_grid_t<E>;
    v.epoch_func = (const void*)k_epoch<E>;
    return v;
}
#define EH_MAKE_TC(PMF, P, NH, H, NOUT, ACT, SCALE) \
    make_variant_epoch_only<EngTc<StepCfg<P, NH, H, NOUT, ACT, SCALE, PMF>>>("tcgen05/" #PMF "/P" #P "/NH" #NH "/H" #H "/O" #NOUT "/" #ACT "/scale=" #SCALE),

// exact-fp32 FFMA2 engine and (where the shape allows) the register-resident tensor-pipe engine
#define EH_MAKE(PMF, P, NH, H, NOUT, ACT, SCALE) \
    make_variant<EngFfma<StepCfg<P, NH, H, NOUT, ACT, SCALE, PMF>>>("ffma2/" #PMF "/P" #P "/NH" #NH "/H" #H "/O" #NOUT "/" #ACT "/scale=" #SCALE),
#define EH_MAKE_X2(PMF, P, NH, H, NOUT, ACT, SCALE) \
    make_variant<EngFfma<StepCfg<P, NH, H, NOUT, ACT, SCALE, PMF, 2>>>("ffma2x2/" #PMF "/P" #P "/NH" #NH "/H" #H "/O" #NOUT "/" #ACT "/scale=" #SCALE),
#define EH_MAKE_MMA(PMF, P, NH, H, NOUT, ACT, SCALE) \
    make_variant<EngMma<StepCfg<P, NH, H, NOUT, ACT, SCALE, PMF>>>("mma3xtf32/" #PMF "/P" #P "/NH" #NH "/H" #H "/O" #NOUT "/" #ACT "/scale=" #SCALE),

}  // namespace eh
