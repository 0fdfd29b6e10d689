// ExpoHybrid variants (BASELINE config 2), sm_100a
#include "eh_variant_impl.cuh"
namespace eh {
#define LIST(X)                               \
    X(PmExpo, 1, 2, 16, 1, ACT_SIGMOID, false)  \
    X(PmExpo, 1, 2, 16, 1, ACT_SIGMOID, true)   \
    X(PmExpo, 1, 2, 16, 1, ACT_TANH, false)
#define LIST_MMA(X) \
    X(PmExpo, 1, 2, 16, 1, ACT_SIGMOID, false) \
    X(PmExpo, 1, 2, 16, 1, ACT_SIGMOID, true) \
    X(PmExpo, 1, 2, 16, 1, ACT_TANH, false)
static const Variant g[] = {LIST(EH_MAKE) LIST_MMA(EH_MAKE_MMA) LIST_MMA(EH_MAKE_X2)};
const Variant* variants_expo(int* n) { *n = (int)(sizeof(g) / sizeof(g[0])); return g; }
}  // namespace eh
