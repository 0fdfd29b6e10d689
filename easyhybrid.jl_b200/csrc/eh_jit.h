// eh_jit.h -- run-time specialisation of traced process models (host side).
//
// A traced process model (EH_PM_PROGRAM: the straight-line program the host traced out of the user's
// `mechanistic_model(; forcings..., params...)`, src/models/GenericHybridModel.jl:425) is interpreted per sample by the
// ahead-of-time generic variants (PmProgram, eh_pm.cuh).  Here the same program is written out as a straight-line C++
// functor (value and reverse sweep) and the three kernels of the register-tile path -- k_step, k_epoch, k_eval -- are
// compiled for it with NVRTC, for the chain shape the planner selected.  Layout constants (record, image, partial
// vector) are those of the PmProgram variant of the same shape, so everything around the kernels stays as it is.
#pragma once
#include <cuda_runtime.h>
#include <string>
#include "eh_variants.h"

namespace eh {

struct PmProgData;

struct JitKernels {
    cudaLibrary_t lib = nullptr;
    const void* k_step = nullptr;    // cudaKernel_t handles, usable wherever the runtime takes a kernel symbol
    const void* k_epoch = nullptr;
    const void* k_eval = nullptr;
    std::string name;                // "nvrtc/PmTraced#<hash>/P./NH./H./O./ACT"
    bool from_cache = false;
    double compile_seconds = 0.0;
};

// generated translation unit for program `pd` on the shape of `shape` (a PmProgram variant); exposed for tests
// unit_act: NULL, or [3][32] activation codes (ACT_*) of the units of hidden layers 1..3 for models whose chains differ in
// activation (the shape's own activation is then only the container: it decides whether swish's sigma rows exist)
std::string jit_source(const PmProgData& pd, const Variant& shape, const unsigned char* unit_act = nullptr);

// source -> cubin (disk cache first, NVRTC otherwise).  No device needed.  Returns false with *err set when NVRTC is not
// available or the compilation fails.  names[3]: lowered names of k_step, k_epoch, k_eval.
bool jit_compile(const PmProgData& pd, const Variant& shape, const unsigned char* unit_act, std::string* cubin, std::string names[3],
                 std::string* tag, bool* from_cache, double* seconds, std::string* err);

// cubin -> kernels on the current device
bool jit_load(const std::string& cubin, const std::string names[3], JitKernels* out, std::string* err);
void jit_unload(JitKernels* k);

// launch plumbing for kernel handles (what the templated launchers of eh_variant_impl.cuh do for compiled-in kernels);
// `args` is the kernel's single by-value argument struct
cudaError_t jit_prepare(const JitKernels& k, size_t step_smem, size_t eval_smem);
cudaError_t jit_launch(const void* kernel, const void* args, int grid, int threads, size_t smem, cudaStream_t st, bool pdl);
cudaError_t jit_launch_cooperative(const void* kernel, const void* args, int grid, int threads, size_t smem, cudaStream_t st);
cudaError_t jit_max_grid(const void* kernel, int threads, size_t smem, int* max_ctas);

}  // namespace eh
