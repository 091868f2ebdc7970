// Launchers of the phase-major direct-form kernels (ddc_kernel_p.cuh): the deferred-epilogue float32 kernel.
#include <algorithm>
#include <cstdio>
#include <cstring>

#include "ddc_host.h"
#include "ddc_kernel_p.cuh"

using namespace ddck;

namespace ddch {
namespace {
template <int D, int JT>
int launch_pd(ddcb200* h, RunParams& p, const float2* ctaps_host, cudaStream_t st) {
    using C = PCfg<D, JT, 1>;
    constexpr int MAXT = JT * D;
    auto kern = ddc_fused_pd_kernel<D, JT, MAXT>;
    const size_t smem = C::HDR_BYTES + (size_t)C::NSLOT * C::SLOT_FLOATS * sizeof(float);
    static bool attr_set[64] = {};
    if (h->device < 64 && !attr_set[h->device]) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[h->device] = true;
    }
    TapsParam<MAXT> tp;
    std::memset(&tp, 0, sizeof(tp));
    std::memcpy(tp.c2, ctaps_host, sizeof(float2) * (size_t)std::min(p.n_taps, MAXT));
    const long long grid = std::min<long long>(p.total_tiles, h->sm_count);
    kern<<<(unsigned)grid, C::NWARPS * 32 + 32 * C::NPROD, smem, st>>>(p, tp);
    CUDA_TRY(cudaGetLastError());
    h->launches++;
    char name[96];
    snprintf(name, sizeof(name), "fused_phase_major_deferred<D%d,R%d,J%d,SLOTS%d>", D, C::R, JT, C::NSLOT);
    h->last_variant = name;
    return DDCB200_OK;
}

template <int D>
int launch_pd_j(ddcb200* h, RunParams& p, const float2* ct, cudaStream_t st, int jt) {
    switch (jt) {
        case 4: return launch_pd<D, 4>(h, p, ct, st);
        case 8: return launch_pd<D, 8>(h, p, ct, st);
        default: return launch_pd<D, 16>(h, p, ct, st);
    }
}
}  // namespace

int launch_pd(ddcb200* h, RunParams& p, const float2* ct, cudaStream_t st, int D, int jt) {
    switch (D) {
        case 32: return launch_pd_j<32>(h, p, ct, st, jt);
        case 64: return launch_pd_j<64>(h, p, ct, st, jt);
    }
    return fail(DDCB200_EINVAL, "phase-major kernel: unsupported decimation %d", D);
}

}  // namespace ddch
