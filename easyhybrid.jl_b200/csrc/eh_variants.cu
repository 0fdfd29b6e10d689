// eh_variants.cu -- registry over the ahead-of-time compiled variant groups.
#include "eh_variants.h"

namespace eh {

constexpr int NGROUPS = 12;
static const Variant* group(int k, int* n)
{
    switch (k) {
    case 0: return variants_rbq10(n);
    case 1: return variants_expo(n);
    case 2: return variants_linear(n);
    case 3: return variants_prog_tanh(n);
    case 4: return variants_prog_sigmoid(n);
    case 5: return variants_prog_relu(n);
    case 6: return variants_prog_swish(n);
    case 7: return variants_prog13_tanh(n);
    case 8: return variants_prog13_sigmoid(n);
    case 9: return variants_prog13_relu(n);
    case 10: return variants_prog13_swish(n);
    case 11: return variants_tc(n);
    default: *n = 0; return nullptr;
    }
}

int num_variants()
{
    int tot = 0, n = 0;
    for (int k = 0; k < NGROUPS; k++) { group(k, &n); tot += n; }
    return tot;
}

const Variant* variant_at(int i)
{
    int n = 0;
    for (int k = 0; k < NGROUPS; k++) {
        const Variant* g = group(k, &n);
        if (i < n) return g + i;
        i -= n;
    }
    return nullptr;
}

const Variant* find_variant(int pm, int P, int NH, int H, int NOUT, int act, int scale, int engine)
{
    const Variant* best = nullptr;
    for (int i = 0; i < num_variants(); i++) {
        const Variant* v = variant_at(i);
        if (v->engine != engine) continue;
        if (v->pm != pm || v->NH != NH || v->NOUT != NOUT || v->act != act || v->scale != scale) continue;
        // hidden widths are padded up to the compiled width; the generic (interpreted process model) variants
        // also pad the chain inputs with zero columns
        if (v->H < H) continue;
        if (pm == PM_PROGRAM ? v->P < P : v->P != P) continue;
        if (!best || v->H < best->H || (v->H == best->H && v->P < best->P)) best = v;
    }
    return best;
}

}  // namespace eh
