// Linear_Regression and two-target linear variants (BASELINE config 2), sm_100a
#include "eh_variant_impl.cuh"
namespace eh {
#define LIST(X)                               \
    X(PmLinear, 2, 2, 16, 1, ACT_RELU, false)   \
    X(PmLinear, 2, 2, 16, 1, ACT_TANH, false)   \
    X(PmLinear, 2, 2, 32, 1, ACT_TANH, false)   \
    X(PmLinear2, 2, 2, 16, 1, ACT_TANH, false)  \
    X(PmLinear2, 2, 2, 16, 1, ACT_RELU, false)
#define LIST_MMA(X) \
    X(PmLinear, 2, 2, 16, 1, ACT_RELU, false) \
    X(PmLinear, 2, 2, 16, 1, ACT_TANH, false) \
    X(PmLinear2, 2, 2, 16, 1, ACT_TANH, false) \
    X(PmLinear2, 2, 2, 16, 1, ACT_RELU, false)
static const Variant g[] = {LIST(EH_MAKE) LIST_MMA(EH_MAKE_MMA) LIST_MMA(EH_MAKE_X2)};
const Variant* variants_linear(int* n) { *n = (int)(sizeof(g) / sizeof(g[0])); return g; }
}  // namespace eh
