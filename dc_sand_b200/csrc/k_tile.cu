// Launchers of the rotating-window tile kernel (ddc_kernels.cuh: ddc_fused_kernel).
#include <algorithm>
#include <cstdio>
#include <cstring>

#include "ddc_host.h"
#include "ddc_kernel_tile.cuh"

using namespace ddck;

namespace ddch {
namespace {
constexpr int kStages = 3;     // TMA pipeline depth
constexpr int kS = 32;         // thread-rows per TMA bulk copy (super-row)
constexpr int kMaxTapsFused = 2048;

template <int D, int R, int KS, int MAXT>
int launch_fused(ddcb200* h, RunParams& p, const float2* ctaps_host, cudaStream_t st, int grid_limit) {
    using C = FusedCfg<D, R, kS, KS>;
    auto kern = ddc_fused_kernel<D, R, kS, KS, kStages, MAXT, false>;
    const size_t smem = 128 + C::XBUF_BYTES + (size_t)kStages * C::stage_floats(p.halo_rows) * sizeof(float);
    if (smem > 227 * 1024) return fail(DDCB200_EINVAL, "fused kernel needs %zu bytes of shared memory", smem);
    static size_t smem_set[64] = {};  // per device
    if (h->device < 64 && smem_set[h->device] < smem) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set[h->device] = smem;
    }
    TapsParam<MAXT> tp;
    std::memset(&tp, 0, sizeof(tp));
    std::memcpy(tp.c2, ctaps_host, sizeof(float2) * (size_t)p.n_taps);
    const long long grid = std::min<long long>(p.total_tiles, grid_limit);
    kern<<<(unsigned)grid, C::NT + 32, smem, st>>>(p, tp);
    CUDA_TRY(cudaGetLastError());
    h->launches++;
    char name[96];
    snprintf(name, sizeof(name), "fused_tma<D%d,R%d,S%d,KS%d,STAGES%d,MAXT%d>", D, R, kS, KS, kStages, MAXT);
    h->last_variant = name;
    return DDCB200_OK;
}

template <int D, int R>
int launch_fused_t(ddcb200* h, RunParams& p, const float2* ct, cudaStream_t st, int grid_limit, int ks) {
    // one tap-parameter size (2048 complex taps = 16 KB of kernel parameters): the kernel is the fall-back for cells no other
    // kernel takes, a second instantiation per (D, KS) for short filters is not worth its 8 entries
    if constexpr (R <= 4) {  // larger R: exchange buffer / register budget do not fit 544 threads
        if (ks == 2) return launch_fused<D, R, 2, kMaxTapsFused>(h, p, ct, st, grid_limit);
    }
    return launch_fused<D, R, 1, kMaxTapsFused>(h, p, ct, st, grid_limit);
}
}  // namespace

int launch_tile(ddcb200* h, RunParams& p, const float2* ct, cudaStream_t st, int D, int ks) {
    const int grid_limit = h->sm_count;
    switch (D) {
        case 4: return launch_fused_t<4, 16>(h, p, ct, st, grid_limit, ks);
        case 8: return launch_fused_t<8, 8>(h, p, ct, st, grid_limit, ks);
        case 16: return launch_fused_t<16, 4>(h, p, ct, st, grid_limit, ks);
        case 32: return launch_fused_t<32, 2>(h, p, ct, st, grid_limit, ks);
        case 64: return launch_fused_t<64, 1>(h, p, ct, st, grid_limit, ks);
    }
    return fail(DDCB200_EINVAL, "tile kernel: unsupported decimation %d", D);
}

}  // namespace ddch
