// eh_variant_impl.cuh -- template glue that turns a StepCfg into a registry entry.
#pragma once
#include "eh_variants.h"
#include "eh_epoch_kernel.cuh"
#include "eh_eval_kernel.cuh"

namespace eh {

template <class C>
static cudaError_t prepare_t(size_t step_smem, size_t eval_smem)
{
    cudaError_t e = cudaFuncSetAttribute(k_step<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)step_smem);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_epoch<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)step_smem);
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

// cooperative launch: the grid barrier inside k_epoch needs every CTA resident
template <class C>
static cudaError_t launch_epoch_t(const EpochArgs& a, int grid, int nwarps, size_t smem, cudaStream_t st)
{
    void* args[] = {(void*)&a};
    return cudaLaunchCooperativeKernel((void*)k_epoch<C>, dim3((unsigned)grid), dim3((unsigned)(nwarps * 32)), args, smem, st);
}

template <class C>
static cudaError_t epoch_max_grid_t(int nwarps, size_t smem, int* blocks_per_sm)
{
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, k_epoch<C>, nwarps * 32, smem);
}

template <class C>
static Variant make_variant(const char* name)
{
    Variant v{};
    v.pm = C::PM::ID; v.P = C::P; v.NH = C::NH; v.H = C::H; v.NOUT = C::NOUT; v.act = C::ACT; v.scale = C::SCALE ? 1 : 0;
    v.dims = C::D;
    v.F = C::F; v.T = C::T; v.NPS = C::NPS; v.R4 = C::R4; v.NB = C::NB; v.NW = C::NW; v.NPART = C::NPART;
    v.stage_floats = C::STAGE_FLOATS;
    v.max_warps = 16;
    v.name = name;
    v.prepare = prepare_t<C>;
    v.launch_step = launch_step_t<C>;
    v.launch_eval = launch_eval_t<C>;
    v.launch_epoch = launch_epoch_t<C>;
    v.epoch_max_grid = epoch_max_grid_t<C>;
    return v;
}

#define EH_MAKE(PMF, P, NH, H, NOUT, ACT, SCALE) \
    make_variant<StepCfg<P, NH, H, NOUT, ACT, SCALE, PMF>>(#PMF "/P" #P "/NH" #NH "/H" #H "/O" #NOUT "/" #ACT "/scale=" #SCALE),

}  // namespace eh
