// Layout descriptors of the generic exact-fp32 shapes whose kernels are compiled at run time (eh_jit.cu): no kernel is
// instantiated here, only the compile-time layout constants of StepCfg / EngFfma are read out.  sm_100a host code
#include <utility>
#include "eh_variant_impl.cuh"
namespace eh {
namespace {
constexpr int SH_P[4] = {2, 4, 8, 12};
constexpr int SH_N = 4 * 3 * 4 * 4 * 4;   // P x NH x H x NOUT x activation
template <int I>
Variant shape_at()
{
    constexpr int P = SH_P[I % 4], NH = 1 + (I / 4) % 3, H = 8 * (1 + (I / 12) % 4), NOUT = 1 + (I / 48) % 4, ACT = 1 + (I / 192) % 4;
    static_assert(ACT_TANH == 1 && ACT_SIGMOID == 2 && ACT_RELU == 3 && ACT_SWISH == 4, "activation codes");
    return make_shape<EngFfma<StepCfg<P, NH, H, NOUT, ACT, true, PmProgram>>>("generic shape (kernels compiled at run time)");
}
template <size_t... I>
const Variant* table(std::index_sequence<I...>)
{
    static const Variant t[] = {shape_at<(int)I>()...};
    return t;
}
}  // namespace

const Variant* find_shape(int P, int NH, int H, int NOUT, int act)
{
    const Variant* t = table(std::make_index_sequence<SH_N>{});
    const Variant* best = nullptr;
    for (int i = 0; i < SH_N; i++) {
        const Variant* v = t + i;
        if (v->NH != NH || v->NOUT != NOUT || v->act != act || v->H < H || v->P < P) continue;
        if (!best || v->H < best->H || (v->H == best->H && v->P < best->P)) best = v;
    }
    return best;
}
}  // namespace eh
