// Launchers of the packed-10-bit fast-FIR kernels (ddc_kernel_w10.cuh).
#include <algorithm>
#include <cstdio>
#include <cstring>

#include "ddc_host.h"
#include "ddc_kernel_w10.cuh"

using namespace ddck;

namespace ddch {
namespace {
template <int D, int JT>
int launch_w10_t(ddcb200* h, RunParams& p, cudaStream_t st, double step) {
    using C = W10Cfg<D, JT>;
    auto kern = ddc_fused_w10_kernel<D, JT>;
    const size_t smem = 512 + (size_t)C::FLOAT_BYTES + (size_t)C::NRAW * C::RAW_BYTES;
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
    snprintf(name, sizeof(name), "fused_fast_fir_packed10<D%d,R%d,J%d,RAWSLOTS%d>", D, C::R, JT, C::NRAW);
    h->last_variant = name;
    return DDCB200_OK;
}

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

template <int D>
int launch_w10_j(ddcb200* h, RunParams& p, cudaStream_t st, double step, int jt) {
    switch (jt) {
        case 4: return launch_w10_t<D, 4>(h, p, st, step);
        case 8: return launch_w10_t<D, 8>(h, p, st, step);
        default: return launch_w10_t<D, 16>(h, p, st, step);
    }
}
}  // namespace

int launch_w10(ddcb200* h, RunParams& p, cudaStream_t st, double step, int D, int jt) {
    switch (D) {
        case 16: return launch_w10_j<16>(h, p, st, step, jt);
        case 32: return launch_w10_j<32>(h, p, st, step, jt);
        case 64: return launch_w10_j<64>(h, p, st, step, jt);
    }
    return fail(DDCB200_EINVAL, "packed fast-FIR kernel: unsupported decimation %d", D);
}

int launch_w10s(ddcb200* h, RunParams& p, cudaStream_t st, double step) { return launch_w10s_t<16, 16>(h, p, st, step); }
}  // namespace ddch
