// eh_variants.h -- registry of ahead-of-time compiled kernel variants (host side).
#pragma once
#include <cuda_runtime.h>
#include "eh_layout.h"

namespace eh {

struct StepArgs;
struct EvalArgs;
struct EpochArgs;

struct Variant {
    int pm, P, NH, H, NOUT, act, scale;
    int engine;      // 0: exact-fp32 FFMA2, one sample per lane; 2: same, two samples per lane (tile partial layout);
                     // 1: tensor pipe 3xTF32 (padded-flat partial layout); 4: tcgen05 tiles of 128 samples per four warps
                     // (persistent kernel only, tile partial layout)
    int chunk;       // samples per warp pass
    ShapeDims dims;
    int F, T, NPS, R4, NW, NPART, off_stats, stage_floats, max_warps;
    int wpc;         // warps that share one chunk (1; tensor engine: 4)
    int eng_bytes;   // engine-private shared memory of the persistent CTA (EpochArgs::eng_off)
    const char* name;
    cudaError_t (*prepare)(size_t step_smem, size_t eval_smem);
    cudaError_t (*launch_step)(const StepArgs& a, int grid, int nwarps, size_t smem, cudaStream_t st, bool pdl);
    cudaError_t (*launch_eval)(const EvalArgs& a, int grid, int nwarps, size_t smem, cudaStream_t st);
    cudaError_t (*launch_epoch)(const EpochArgs& a, int grid, int nwarps, size_t smem, cudaStream_t st);
    cudaError_t (*epoch_max_grid)(int nwarps, size_t smem, int* max_ctas);
    const void* epoch_func;   // k_epoch<E>, for the graph node of the persistent launch
};

const Variant* find_variant(int pm, int P, int NH, int H, int NOUT, int act, int scale, int engine);
// generic (PmProgram-layout) shapes that exist as descriptors only -- their kernels are compiled at run time (eh_jit.cu):
// chain inputs padded to 2 / 4 / 8 / 12, one to three hidden layers of width 8 / 16 / 24 / 32, one to four chain outputs,
// any activation.  The tightest shape that holds (P, H); NULL if none does.
const Variant* find_shape(int P, int NH, int H, int NOUT, int act);
int num_variants();
const Variant* variant_at(int i);

// one translation unit per process-model family (parallel build)
const Variant* variants_rbq10(int* n);
const Variant* variants_expo(int* n);
const Variant* variants_linear(int* n);
const Variant* variants_tc(int* n);   // tensor engine (eh_engine_tc.cuh), persistent kernel only
// generic exact-fp32 variants: the process model is a traced program interpreted per sample (PmProgram);
// one translation unit per activation
const Variant* variants_prog_tanh(int* n);
const Variant* variants_prog_sigmoid(int* n);
const Variant* variants_prog_relu(int* n);
const Variant* variants_prog_swish(int* n);
const Variant* variants_prog13_tanh(int* n);
const Variant* variants_prog13_sigmoid(int* n);
const Variant* variants_prog13_relu(int* n);
const Variant* variants_prog13_swish(int* n);

}  // namespace eh
