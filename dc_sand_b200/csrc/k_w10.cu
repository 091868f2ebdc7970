// Launcher of the CUDA-core packed-10-bit kernel (ddc_kernel_w10.cuh).
#include <algorithm>
#include <cstdio>
#include <cstring>

#include "ddc_host.h"
#include "ddc_kernel_w10.cuh"

using namespace ddck;

namespace ddch {
namespace {
template <int D, int JT>
int launch_w10s_t(ddcb200* h, RunParams& p, cudaStream_t st, double step) {
    using C = W10SCfg<D, JT>;
    auto kern = ddc_fused_w10s_kernel<D, JT>;
    static bool attr_set[64] = {};
    if (h->device < 64 && !attr_set[h->device]) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
        attr_set[h->device] = true;
    }
    TapsParam<C::NTW> tp;
    std::memcpy(tp.c2, cached_wtaps(h, step, JT, D), sizeof(float2) * (size_t)C::NTW);
    const long long grid = std::min<long long>(p.total_tiles, h->sm_count);
    kern<<<(unsigned)grid, (C::NFIR + C::NUNP + 1) * 32, C::SMEM, st>>>(p, tp);
    CUDA_TRY(cudaGetLastError());
    h->launches++;
    char name[112];
    snprintf(name, sizeof(name), "fused_fast_fir_packed10_split<D%d,R%d,J%d,RAW%d,FLOAT%d,UNPACK%d>", D, C::R, JT, C::NR, C::NF, C::NUNP);
    h->last_variant = name;
    return DDCB200_OK;
}

}  // namespace

int launch_w10s(ddcb200* h, RunParams& p, cudaStream_t st, double step) { return launch_w10s_t<16, 16>(h, p, st, step); }
}  // namespace ddch
