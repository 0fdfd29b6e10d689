// eh_variants.cu -- registry over the ahead-of-time compiled variant groups.
#include "eh_variants.h"

namespace eh {

static const Variant* group(int k, int* n)
{
    switch (k) {
    case 0: return variants_rbq10(n);
    case 1: return variants_expo(n);
    case 2: return variants_linear(n);
    default: *n = 0; return nullptr;
    }
}

int num_variants()
{
    int tot = 0, n = 0;
    for (int k = 0; k < 3; k++) { group(k, &n); tot += n; }
    return tot;
}

const Variant* variant_at(int i)
{
    int n = 0;
    for (int k = 0; k < 3; k++) {
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
        if (v->pm != pm || v->P != P || v->NH != NH || v->NOUT != NOUT || v->act != act || v->scale != scale) continue;
        if (v->H < H) continue;
        if (!best || v->H < best->H) best = v;
    }
    return best;
}

}  // namespace eh
