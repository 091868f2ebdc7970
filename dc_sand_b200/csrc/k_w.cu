// Launchers of the fast-FIR kernel for float32 input (ddc_kernel_w.cuh: ddc_fused_w_kernel).
#include <algorithm>
#include <cstdio>
#include <cstring>

#include "ddc_host.h"
#include "ddc_kernel_w.cuh"

using namespace ddck;

namespace ddch {
namespace {
template <int D, int JT>
int launch_w_t(ddcb200* h, RunParams& p, cudaStream_t st, double step) {
    using C = WCfg<D, JT>;
    auto kern = ddc_fused_w_kernel<D, JT>;
    const size_t smem = C::HDR_BYTES + (size_t)C::NSLOT * C::SLOT_FLOATS * sizeof(float);
    static bool attr_set[64] = {};
    if (h->device < 64 && !attr_set[h->device]) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[h->device] = true;
    }
    TapsParam<C::NTW> tp;
    std::memcpy(tp.c2, cached_wtaps(h, step, JT, D), sizeof(float2) * (size_t)C::NTW);
    const long long grid = std::min<long long>(p.total_tiles, h->sm_count);
    kern<<<(unsigned)grid, C::NWARPS * 32 + 32 * C::NPROD, smem, st>>>(p, tp);
    CUDA_TRY(cudaGetLastError());
    h->launches++;
    char name[96];
    snprintf(name, sizeof(name), "fused_fast_fir<D%d,R%d,J%d,SLOTS%d>", D, C::R, JT, C::NSLOT);
    h->last_variant = name;
    return DDCB200_OK;
}

template <int D>
int launch_w_j(ddcb200* h, RunParams& p, cudaStream_t st, double step, int jt) {
    switch (jt) {
        case 4: return launch_w_t<D, 4>(h, p, st, step);
        case 8: return launch_w_t<D, 8>(h, p, st, step);
        case 16: return launch_w_t<D, 16>(h, p, st, step);
        default: break;
    }
    if constexpr (D == 16) {   // long filters: passes of 16 tap blocks (T <= 1024)
        if (jt == 32) return launch_w_t<D, 32>(h, p, st, step);
        if (jt == 64) return launch_w_t<D, 64>(h, p, st, step);
    }
    return fail(DDCB200_EINVAL, "fast-FIR kernel: unsupported tap-block count %d at D = %d", jt, D);
}

}  // namespace

int launch_w(ddcb200* h, RunParams& p, cudaStream_t st, double step, int D, int jt) {
    switch (D) {
        case 16: return launch_w_j<16>(h, p, st, step, jt);   // D = 32 / 64: the sliced kernel (k_ws*.cu) does the fast FIR there
    }
    return fail(DDCB200_EINVAL, "fast-FIR kernel: unsupported decimation %d", D);
}
}  // namespace ddch
