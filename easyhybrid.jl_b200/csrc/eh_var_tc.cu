// tensor-engine variants (tcgen05, eh_engine_tc.cuh) of the BASELINE shapes with a 16-wide hidden layer, sm_100a
#include "eh_variant_impl.cuh"
namespace eh {
#define LIST(X)                              \
    X(PmRbQ10, 2, 2, 16, 1, ACT_TANH, true)    \
    X(PmRbQ10, 2, 2, 16, 1, ACT_TANH, false)   \
    X(PmRbQ10, 2, 2, 16, 1, ACT_SIGMOID, true)
static const Variant g[] = {LIST(EH_MAKE_TC)};
const Variant* variants_tc(int* n) { *n = (int)(sizeof(g) / sizeof(g[0])); return g; }
}  // namespace eh
