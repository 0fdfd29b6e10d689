// RbQ10 variants (BASELINE configs 1, 3, 4; README / docs activations), sm_100a
#include "eh_variant_impl.cuh"
namespace eh {
#define LIST(X)                              \
    X(PmRbQ10, 2, 2, 16, 1, ACT_TANH, true)    \
    X(PmRbQ10, 2, 2, 16, 1, ACT_SWISH, true)   \
    X(PmRbQ10, 2, 2, 16, 1, ACT_SIGMOID, true) \
    X(PmRbQ10, 2, 2, 16, 1, ACT_TANH, false)   \
    X(PmRbQ10, 2, 2, 32, 1, ACT_TANH, true)    \
    X(PmRbQ10, 2, 2, 32, 1, ACT_TANH, false)
#define LIST_MMA(X) \
    X(PmRbQ10, 2, 2, 16, 1, ACT_TANH, true) \
    X(PmRbQ10, 2, 2, 16, 1, ACT_SWISH, true) \
    X(PmRbQ10, 2, 2, 16, 1, ACT_SIGMOID, true) \
    X(PmRbQ10, 2, 2, 16, 1, ACT_TANH, false)
static const Variant g[] = {LIST(EH_MAKE) LIST_MMA(EH_MAKE_MMA) LIST_MMA(EH_MAKE_X2)};
const Variant* variants_rbq10(int* n) { *n = (int)(sizeof(g) / sizeof(g[0])); return g; }
}  // namespace eh
