// Generic exact-fp32 variants, activation ACT_TANH, two hidden layers (chain inputs padded to 2 / 4 / 8): process model = traced program interpreted per sample
// (PmProgram); hidden width 16 / 32, one or two chain outputs; scale_nn_outputs is a run-time flag (compiled with
// SCALE = true).  One translation unit per (activation, depth group) keeps the parallel build balanced.  sm_100a
#include "eh_variant_impl.cuh"
namespace eh {
#define LIST(X) \
    X(PmProgram, 2, 2, 16, 1, ACT_TANH, true) \
    X(PmProgram, 4, 2, 16, 1, ACT_TANH, true) \
    X(PmProgram, 8, 2, 16, 1, ACT_TANH, true) \
    X(PmProgram, 2, 2, 32, 1, ACT_TANH, true) \
    X(PmProgram, 4, 2, 32, 1, ACT_TANH, true) \
    X(PmProgram, 8, 2, 32, 1, ACT_TANH, true) \
    X(PmProgram, 2, 2, 16, 2, ACT_TANH, true) \
    X(PmProgram, 4, 2, 16, 2, ACT_TANH, true) \
    X(PmProgram, 8, 2, 16, 2, ACT_TANH, true) \
    X(PmProgram, 2, 2, 32, 2, ACT_TANH, true) \
    X(PmProgram, 4, 2, 32, 2, ACT_TANH, true) \
    X(PmProgram, 8, 2, 32, 2, ACT_TANH, true)
static const Variant g[] = {LIST(EH_MAKE)};
const Variant* variants_prog_tanh(int* n) { *n = (int)(sizeof(g) / sizeof(g[0])); return g; }
}  // namespace eh
