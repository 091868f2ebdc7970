// C-ABI implementation of include/ddcb200.h: handle management, per-call host-side preparation of the folded
// complex taps (float64 -> float32), kernel dispatch, and the chunked double-buffered host path.
#include "../../include/ddcb200.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "ddc_kernels.cuh"
#include "ddc_kernel_p.cuh"
#include "ddc_kernel_w.cuh"
#include "ddc_kernel_ws.cuh"

using namespace ddck;

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CUDA_TRY(expr)                                                                                      \
    do {                                                                                                    \
        cudaError_t e_ = (expr);                                                                            \
        if (e_ != cudaSuccess) return fail(DDCB200_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

constexpr int kStages = 3;     // TMA pipeline depth
constexpr int kS = 32;         // thread-rows per TMA bulk copy (super-row)
constexpr int kMaxTapsFused = 2048;

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) ok = false;
        if (ok && prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// Strided copy that never hands the driver a pitch it may reject (cudaDeviceProp::memPitch is 2^31 - 1 bytes): one row, or
// rows that happen to be contiguous, go as ONE flat copy; rows with a pitch beyond the limit go one by one.
cudaError_t copy_rows_async(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t rows,
                            cudaMemcpyKind kind, cudaStream_t st) {
    if (rows == 0 || width == 0) return cudaSuccess;
    if (rows == 1 || (dpitch == width && spitch == width)) return cudaMemcpyAsync(dst, src, width * rows, kind, st);
    constexpr size_t kMaxPitch = 0x7fffffffull;
    if (dpitch <= kMaxPitch && spitch <= kMaxPitch) return cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, rows, kind, st);
    for (size_t r = 0; r < rows; ++r) {
        cudaError_t e = cudaMemcpyAsync(static_cast<char*>(dst) + r * dpitch, static_cast<const char*>(src) + r * spitch, width, kind, st);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace

struct ddcb200 {
    int device = 0;
    int sm_count = 0;
    int decim = 1;
    std::vector<double> taps;       // raw, file order
    double taps_sum = 0.0;
    cudaStream_t stream = nullptr;  // compute
    cudaStream_t copy_in = nullptr, copy_out = nullptr;
    // device scratch for the generic/short kernels' taps (ring so that back-to-back calls do not race)
    static constexpr int kRing = 8;
    float2* d_ctaps[kRing] = {};
    float2* h_ctaps[kRing] = {};    // pinned
    cudaEvent_t ring_ev[kRing] = {};
    int ring_cap = 0, ring_pos = 0;
    // host-path workspace
    static constexpr int kBufs = 3;
    void* d_chunk_in[kBufs] = {};
    ddcb200_c64* d_chunk_out[kBufs] = {};
    size_t chunk_in_cap = 0, chunk_out_cap = 0;
    cudaEvent_t ev_in[kBufs] = {}, ev_k[kBufs] = {}, ev_out[kBufs] = {};
    int64_t chunk_samples = 1 << 24;
    int64_t launches = 0;
    int force_variant = 0;
    int debug_mode = 0;
    int stagger_cycles = 0;
    int l2_ahead = 0;
    unsigned long long* d_dbg = nullptr;   // diagnostic counters (option "dbg_counters")
    std::string last_variant = "none";
    bool smem_attr_set = false;
    // folded fast-FIR taps of the last (step, jt, D): streaming callers repeat the same step call after call
    std::vector<float2> wt_cache;
    double wt_step = 0.0;
    int wt_jt = 0, wt_d = 0, wt_nest = 0;
    // pinned staging for pageable host input (see staged_h2d)
    static constexpr int kStage = 2;
    static constexpr size_t kStageBytes = 8u << 20;
    void* h_stage[kStage] = {};
    cudaEvent_t ev_stage[kStage] = {};
    int stage_pos = 0;
    int copy_threads = 4;
    float* d_unpack_ws = nullptr;          // float32 workspace of the two-launch packed path (unpack, then a float32 kernel)
    size_t unpack_ws_cap = 0;
    cudaEvent_t unpack_ev = nullptr;       // end of the last kernel that read the workspace (calls may come on different streams)
    ddcb200_c64* h_ostage[kBufs] = {};   // pinned landing buffers for the complex128 host path (one per chunk buffer)
    size_t ostage_cap = 0;
    std::vector<float2> wq_cache;   // same for the small-decimation kernel
    double wq_step = 0.0;
    int wq_jt = 0, wq_nq = 0;
};

namespace {

// Host-path calls queue work on three streams and read / write the caller's host buffers asynchronously: whatever way such a
// call ends (also a mid-pipeline error), nothing may still be in flight when it returns -- the caller is free to release the
// buffers.  Synchronising idle streams costs nothing on the success path.
struct DrainGuard {
    ddcb200* h;
    explicit DrainGuard(ddcb200* hh) : h(hh) {}
    ~DrainGuard() {
        cudaStreamSynchronize(h->copy_in);
        cudaStreamSynchronize(h->stream);
        cudaStreamSynchronize(h->copy_out);
    }
};

// c[k] = taps[T-1-k]/sum * exp(-j 2 pi k step), float64 -> float32; zero padded to n_pad
void make_ctaps(const ddcb200* h, double step, int n_pad, float2* out) {
    const int T = (int)h->taps.size();
    const double fstep = step - std::floor(step);
    for (int k = 0; k < n_pad; ++k) {
        if (k < T) {
            const double hk = h->taps[T - 1 - k] / h->taps_sum;
            double ph = fstep * (double)k;
            ph -= std::floor(ph);
            const double a = -2.0 * M_PI * ph;
            out[k] = make_float2((float)(hk * std::cos(a)), (float)(hk * std::sin(a)));
        } else {
            out[k] = make_float2(0.f, 0.f);
        }
    }
}

// Kernel W (fast FIR, ddc_kernel_w.cuh): for tap pair i and phase d the three tap sets
//   seq 0: c[2i D + d],  seq 1: c[2i D + d] + c[(2i+1) D + d],  seq 2: c[(2i+1) D + d]      at index (3i + seq) D + d,
// formed in float64 from the folded taps and rounded to float32 once.
void make_wtaps(const ddcb200* h, double step, int jt, int D, float2* out) {
    const int T = (int)h->taps.size();
    const double fstep = step - std::floor(step);
    auto c = [&](int k, double& re, double& im) {
        if (k >= T) { re = im = 0.0; return; }
        const double hk = h->taps[T - 1 - k] / h->taps_sum;
        double ph = fstep * (double)k;
        ph -= std::floor(ph);
        const double a = -2.0 * M_PI * ph;
        re = hk * std::cos(a);
        im = hk * std::sin(a);
    };
    for (int i = 0; i < jt / 2; ++i)
        for (int d = 0; d < D; ++d) {
            double er, ei, orr, oi;
            c(2 * i * D + d, er, ei);
            c((2 * i + 1) * D + d, orr, oi);
            out[(3 * i + 0) * D + d] = make_float2((float)er, (float)ei);
            out[(3 * i + 1) * D + d] = make_float2((float)(er + orr), (float)(ei + oi));
            out[(3 * i + 2) * D + d] = make_float2((float)orr, (float)oi);
        }
}

// Kernel WQ (small decimations, ddc_kernel_w.cuh): NQ = 16 / D shifted tap sets c_q[k'] = c[k' - D q] (same phase law in k'),
// each laid out like kernel W's: index q * 3 (jt/2) 16 + (3 i + seq) 16 + d.
void make_wqtaps(const ddcb200* h, double step, int jt, int nq, int D, float2* out) {
    const int T = (int)h->taps.size();
    const double fstep = step - std::floor(step);
    for (int q = 0; q < nq; ++q) {
        auto c = [&](int kp, double& re, double& im) {
            const int k = kp - D * q;
            if (k < 0 || k >= T) { re = im = 0.0; return; }
            const double hk = h->taps[T - 1 - k] / h->taps_sum;
            double ph = fstep * (double)kp;
            ph -= std::floor(ph);
            const double a = -2.0 * M_PI * ph;
            re = hk * std::cos(a);
            im = hk * std::sin(a);
        };
        float2* o = out + (size_t)q * 3 * (jt / 2) * 16;
        for (int i = 0; i < jt / 2; ++i)
            for (int d = 0; d < 16; ++d) {
                double er, ei, orr, oi;
                c(2 * i * 16 + d, er, ei);
                c((2 * i + 1) * 16 + d, orr, oi);
                o[(3 * i + 0) * 16 + d] = make_float2((float)er, (float)ei);
                o[(3 * i + 1) * 16 + d] = make_float2((float)(er + orr), (float)(ei + oi));
                o[(3 * i + 2) * 16 + d] = make_float2((float)orr, (float)oi);
            }
    }
}

// Nested fast FIR (w2_fir_pg): for tap quad iota and phase d the nine tap sets (a, b) at index ((9 iota + 3 a + b) D + d):
//   g_0[i] = c[2i], g_1[i] = c[2i] + c[2i+1], g_2[i] = c[2i+1];   h_a0 = g_a[2 iota], h_a1 = g_a[2 iota] + g_a[2 iota + 1], h_a2 = g_a[2 iota + 1]
void make_w2taps(const ddcb200* h, double step, int jt, int D, float2* out) {
    const int T = (int)h->taps.size();
    const double fstep = step - std::floor(step);
    auto c = [&](int k, double* v) {
        v[0] = v[1] = 0.0;
        if (k >= T) return;
        const double hk = h->taps[T - 1 - k] / h->taps_sum;
        double ph = fstep * (double)k;
        ph -= std::floor(ph);
        const double a = -2.0 * M_PI * ph;
        v[0] = hk * std::cos(a);
        v[1] = hk * std::sin(a);
    };
    for (int io = 0; io < jt / 4; ++io)
        for (int d = 0; d < D; ++d) {
            double cb[4][2], g[3][2][2];   // c of blocks 4 io .. 4 io + 3; g[a][i - 2 io]
            for (int b = 0; b < 4; ++b) c((4 * io + b) * D + d, cb[b]);
            for (int i = 0; i < 2; ++i)
                for (int z = 0; z < 2; ++z) {
                    g[0][i][z] = cb[2 * i][z];
                    g[1][i][z] = cb[2 * i][z] + cb[2 * i + 1][z];
                    g[2][i][z] = cb[2 * i + 1][z];
                }
            for (int a = 0; a < 3; ++a) {
                float2* o = out + (size_t)(9 * io + 3 * a) * D + d;
                o[0] = make_float2((float)g[a][0][0], (float)g[a][0][1]);
                o[D] = make_float2((float)(g[a][0][0] + g[a][1][0]), (float)(g[a][0][1] + g[a][1][1]));
                o[2 * D] = make_float2((float)g[a][1][0], (float)g[a][1][1]);
            }
        }
}

// cached front end of make_wtaps (invalidated by set_taps / set_decimation through wt_jt = 0)
const float2* cached_wtaps(ddcb200* h, double step, int jt, int D) {
    if (h->wt_jt != jt || h->wt_d != D || h->wt_nest != 0 || h->wt_step != step || h->wt_cache.size() != (size_t)(3 * (jt / 2) * D)) {
        h->wt_nest = 0;
        h->wt_cache.resize((size_t)(3 * (jt / 2) * D));
        make_wtaps(h, step, jt, D, h->wt_cache.data());
        h->wt_step = step;
        h->wt_jt = jt;
        h->wt_d = D;
    }
    return h->wt_cache.data();
}

unsigned long long to_fx64(double frac01) {
    // frac01 in [0,1) -> round(frac * 2^64) mod 2^64 using long double (64-bit mantissa on x86)
    long double v = (long double)frac01 * 18446744073709551616.0L;
    if (v >= 18446744073709551615.0L) return 0ull;
    return (unsigned long long)(v + 0.5L);
}

unsigned long long phase_of(double step, int64_t sample_offset) {
    // frac(sample_offset * step) * 2^64, exact modular arithmetic on the fixed-point step
    const double fstep = step - std::floor(step);
    const unsigned long long sfx = to_fx64(fstep);
    return (unsigned long long)sample_offset * sfx;
}

int ensure_ring(ddcb200* h, int n_taps) {
    if (n_taps <= h->ring_cap) return DDCB200_OK;
    const int cap = std::max(n_taps, 1024);
    for (int i = 0; i < ddcb200::kRing; ++i) {
        if (h->d_ctaps[i]) cudaFree(h->d_ctaps[i]);
        if (h->h_ctaps[i]) cudaFreeHost(h->h_ctaps[i]);
        h->d_ctaps[i] = nullptr;
        h->h_ctaps[i] = nullptr;
        CUDA_TRY(cudaMalloc(&h->d_ctaps[i], sizeof(float2) * cap));
        CUDA_TRY(cudaMallocHost(&h->h_ctaps[i], sizeof(float2) * cap));
        if (!h->ring_ev[i]) CUDA_TRY(cudaEventCreateWithFlags(&h->ring_ev[i], cudaEventDisableTiming));
    }
    h->ring_cap = cap;
    return DDCB200_OK;
}

static inline bool aligned_f32(const void* d_in, int64_t in_stride, bool packed) {
    return !packed && (reinterpret_cast<uintptr_t>(d_in) % 16 == 0) && (in_stride % 4 == 0);
}

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
    if constexpr (R <= 4) {  // larger R: exchange buffer / register budget do not fit 544 threads
        if (ks == 2) {
            if (p.n_taps <= 512) return launch_fused<D, R, 2, 512>(h, p, ct, st, grid_limit);
            return launch_fused<D, R, 2, kMaxTapsFused>(h, p, ct, st, grid_limit);
        }
    }
    if (p.n_taps <= 512) return launch_fused<D, R, 1, 512>(h, p, ct, st, grid_limit);
    return launch_fused<D, R, 1, kMaxTapsFused>(h, p, ct, st, grid_limit);
}

template <int D, int JT, int KS>
int launch_p(ddcb200* h, RunParams& p, const float2* ctaps_host, cudaStream_t st) {
    using C = PCfg<D, JT, KS>;
    constexpr int MAXT = JT * D;
    auto kern = ddc_fused_p_kernel<D, JT, KS, MAXT>;
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
    snprintf(name, sizeof(name), "fused_phase_major<D%d,R%d,J%d,KS%d,SLOTS%d>", D, C::R, JT, KS, C::NSLOT);
    h->last_variant = name;
    return DDCB200_OK;
}

template <int D, int JT>
int launch_p10(ddcb200* h, RunParams& p, const float2* ctaps_host, cudaStream_t st) {
    using C = P10Cfg<D, JT>;
    constexpr int MAXT = JT * D;
    auto kern = ddc_fused_p10_kernel<D, JT, MAXT>;
    const size_t smem = 512 + (size_t)C::FLOAT_BYTES + (size_t)C::NRAW * C::RAW_BYTES;
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
    snprintf(name, sizeof(name), "fused_phase_major_packed10<D%d,R%d,J%d,RAWSLOTS%d>", D, C::R, JT, C::NRAW);
    h->last_variant = name;
    return DDCB200_OK;
}

template <int D>
int launch_p10_j(ddcb200* h, RunParams& p, const float2* ct, cudaStream_t st, int jt) {
    switch (jt) {
        case 4: return launch_p10<D, 4>(h, p, ct, st);
        case 8: return launch_p10<D, 8>(h, p, ct, st);
        default: return launch_p10<D, 16>(h, p, ct, st);
    }
}

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

template <int D, int JT, int NEST>
int launch_w(ddcb200* h, RunParams& p, cudaStream_t st, double step) {
    using C = WCfg<D, JT>;
    constexpr int NT = NEST ? C::NTW2 : C::NTW;
    auto kern = ddc_fused_w_kernel<D, JT, NEST>;
    const size_t smem = C::HDR_BYTES + (size_t)C::NSLOT * C::SLOT_FLOATS * sizeof(float);
    static bool attr_set[64] = {};
    if (h->device < 64 && !attr_set[h->device]) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[h->device] = true;
    }
    TapsParam<NT> tp;
    // folded taps of the last (step, jt, D, nesting): streaming callers repeat the same step call after call
    if (h->wt_jt != JT || h->wt_d != D || h->wt_nest != NEST || h->wt_step != step || h->wt_cache.size() != (size_t)NT) {
        h->wt_cache.assign((size_t)NT, make_float2(0.f, 0.f));
        if (NEST) make_w2taps(h, step, JT, D, h->wt_cache.data());
        else make_wtaps(h, step, JT, D, h->wt_cache.data());
        h->wt_step = step;
        h->wt_jt = JT;
        h->wt_d = D;
        h->wt_nest = NEST;
    }
    std::memcpy(tp.c2, h->wt_cache.data(), sizeof(float2) * (size_t)NT);
    const long long grid = std::min<long long>(p.total_tiles, h->sm_count);
    kern<<<(unsigned)grid, C::NWARPS * 32 + 32 * C::NPROD, smem, st>>>(p, tp);
    CUDA_TRY(cudaGetLastError());
    h->launches++;
    char name[96];
    snprintf(name, sizeof(name), "fused_fast_fir%s<D%d,R%d,J%d,SLOTS%d>", NEST ? "_nested" : "", D, C::R, JT, C::NSLOT);
    h->last_variant = name;
    return DDCB200_OK;
}

template <int D, int JT>
int launch_w2x(ddcb200* h, RunParams& p, cudaStream_t st, double step) {
    using C = WCfg<D, JT>;
    auto kern = ddc_fused_w2x_kernel<D, JT>;
    const size_t smem = C::HDR_BYTES + (size_t)C::NSLOT * C::SLOT_FLOATS * sizeof(float);
    static bool attr_set[64] = {};
    if (h->device < 64 && !attr_set[h->device]) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[h->device] = true;
    }
    TapsParam<C::NTW> tp;
    std::memcpy(tp.c2, cached_wtaps(h, step, JT, D), sizeof(float2) * (size_t)C::NTW);
    const long long grid = std::min<long long>(p.total_tiles, h->sm_count);
    kern<<<(unsigned)grid, 2 * C::NWARPS * 32 + 32 * C::NPROD, smem, st>>>(p, tp);
    CUDA_TRY(cudaGetLastError());
    h->launches++;
    char name[96];
    snprintf(name, sizeof(name), "fused_fast_fir_16w<D%d,R%d,J%d,SLOTS%d>", D, C::R, JT, C::NSLOT);
    h->last_variant = name;
    return DDCB200_OK;
}

template <int D>
int launch_w_j(ddcb200* h, RunParams& p, cudaStream_t st, double step, int jt, bool nest) {
    if constexpr (D == 16) {   // two nested levels need R = 8 outputs per thread
        // Experimental (option "variant" 9): measured SLOWER than one level -- 0.252 against 0.229 ms compute-only at
        // T = 256 -- because each tap fetch then feeds only two FFMA2 (R / 4 outputs); kept for one configuration only.
        if (nest && jt == 16) return launch_w<D, 16, 1>(h, p, st, step);
    }
    switch (jt) {
        case 4: return launch_w<D, 4, 0>(h, p, st, step);
        case 8: return launch_w<D, 8, 0>(h, p, st, step);
        case 16: return launch_w<D, 16, 0>(h, p, st, step);
        default: break;
    }
    if constexpr (D == 16) {   // long filters: passes of 16 tap blocks (T <= 1024)
        if (jt == 32) return launch_w<D, 32, 0>(h, p, st, step);
        if (jt == 64) return launch_w<D, 64, 0>(h, p, st, step);
    }
    return fail(DDCB200_EINVAL, "fast-FIR kernel: unsupported tap-block count %d at D = %d", jt, D);
}

template <int D, int JT>
int launch_w10(ddcb200* h, RunParams& p, cudaStream_t st, double step) {
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
int launch_w10s(ddcb200* h, RunParams& p, cudaStream_t st, double step) {
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
        case 4: return launch_w10<D, 4>(h, p, st, step);
        case 8: return launch_w10<D, 8>(h, p, st, step);
        default: return launch_w10<D, 16>(h, p, st, step);
    }
}

template <int JT, int NQ>
int launch_wq(ddcb200* h, RunParams& p, cudaStream_t st, double step) {
    using C = WQCfg<JT, NQ>;
    auto kern = ddc_fused_wq_kernel<JT, NQ>;
    const size_t smem = C::HDR_BYTES + (size_t)C::NSLOT * C::SLOT_FLOATS * sizeof(float);
    static bool attr_set[64] = {};
    if (h->device < 64 && !attr_set[h->device]) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set[h->device] = true;
    }
    static_assert(sizeof(TapsParam<C::NTW>) + sizeof(RunParams) <= 32764, "kernel parameter space");
    TapsParam<C::NTW> tp;
    if (h->wq_jt != JT || h->wq_nq != NQ || h->wq_step != step || h->wq_cache.size() != (size_t)C::NTW) {
        h->wq_cache.resize((size_t)C::NTW);
        make_wqtaps(h, step, JT, NQ, 16 / NQ, h->wq_cache.data());
        h->wq_jt = JT;
        h->wq_nq = NQ;
        h->wq_step = step;
    }
    std::memcpy(tp.c2, h->wq_cache.data(), sizeof(float2) * (size_t)C::NTW);
    const long long grid = std::min<long long>(p.total_tiles, h->sm_count);
    kern<<<(unsigned)grid, C::NWARPS * 32 + 32 * C::NPROD, smem, st>>>(p, tp);
    CUDA_TRY(cudaGetLastError());
    h->launches++;
    char name[96];
    snprintf(name, sizeof(name), "fused_fast_fir_subfilters<D%d,NQ%d,J%d,SLOTS%d>", 16 / NQ, NQ, JT, C::NSLOT);
    h->last_variant = name;
    return DDCB200_OK;
}

// The small-decimation kernel is only instantiated where it was measured faster than the rotating-window tile kernel:
// D = 8 with 513 .. 1048 taps (66 tap blocks), 0.495 against 0.545 ms at T = 1024, N = 2^26.  With fewer tap blocks the
// tile kernel's 82-94 % FMA-pipe utilisation beats the 7 % net flop saving (profiles/r1_sweep_taps_decimation.md).
template <int NQ>
int launch_wq_j(ddcb200* h, RunParams& p, cudaStream_t st, double step, int jt) {
    if constexpr (NQ == 2) {
        if (jt == 66) return launch_wq<66, NQ>(h, p, st, step);
    }
    return fail(DDCB200_EINVAL, "small-decimation kernel: unsupported tap-block count %d", jt);
}

// ---- kernel WS (sliced staging, D = 32 / 64): tensor map over the input + launch -------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled get_encode_tiled() {
    // magic static: initialised exactly once, thread-safe (handles on different devices are driven from different threads)
    static const PFN_encodeTiled fn = [] {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            return reinterpret_cast<PFN_encodeTiled>(sym);
        cudaGetLastError();
        return static_cast<PFN_encodeTiled>(nullptr);
    }();
    return fn;
}

// Per-block tiles: dims (fastest first) DB floats of a slice (DB = min(D, 16)), D/16 slices (1 for D < 16), 8 blocks of a
// thread-row, thread-rows, streams; box = (DB, 1, 1, 36, 1): the slice of one block index for the 36 thread-rows of a slot,
// lines of DB * 4 bytes written with the swizzle of that span (64 B / 32 B / none).
// Whole-row tiles (D = 4 / 8, short filters -- WSCfg::WHOLE): dims (32 floats = 128 bytes, 1, 128-byte pieces of a thread-row,
// thread-rows, streams), box (32, 1, 1, 36, 1), 128-byte swizzle; the same coordinate order (0, slice, tile, row, stream).
int make_slice_tmap(const float* d_in, int D, bool whole, int slot_rows, long long n_rows, long long n_streams, long long in_stride, CUtensorMap* out) {
    PFN_encodeTiled enc = get_encode_tiled();
    if (!enc) return fail(DDCB200_ECUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const int DB = D < 16 ? D : 16;
    const cuuint64_t gdim[5] = {(cuuint64_t)(whole ? 32 : DB), (cuuint64_t)(D < 16 ? 1 : D / 16), (cuuint64_t)(whole ? D / 4 : 8),
                                (cuuint64_t)n_rows, (cuuint64_t)n_streams};
    const cuuint64_t gstr[4] = {(cuuint64_t)(whole ? 128 : 64), (cuuint64_t)(whole ? 128 : D * 4), (cuuint64_t)D * 32,
                                (cuuint64_t)in_stride * 4};
    const cuuint32_t box[5] = {(cuuint32_t)(whole ? 32 : DB), 1, 1, (cuuint32_t)slot_rows, 1};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    // L2 promotion: the sliced gather (64 of every 4 D bytes per request) runs 5 % faster when a request promotes only its own
    // 64 bytes (T=512,D=32 0.0718 -> 0.0682 ms, T=1024,D=64 0.0740 -> 0.0702 ms; 256 B is 20 % slower at D = 64); whole blocks
    // (D <= 16) read every byte of a 128-byte line within one chunk and are equal or slower with 64 B (T=128,D=8 0.0701 -> 0.0735)
    const CUtensorMapL2promotion promo = D >= 32 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
    const CUtensorMapSwizzle sw = whole      ? CU_TENSOR_MAP_SWIZZLE_128B
                                  : DB == 16 ? CU_TENSOR_MAP_SWIZZLE_64B
                                             : (DB == 8 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_NONE);
    const CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(d_in), gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, sw, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(DDCB200_ECUDA, "cuTensorMapEncodeTiled failed with code %d", (int)r);
    return DDCB200_OK;
}

template <int D, int JT>
int launch_ws(ddcb200* h, RunParams& p, const float* d_in, long long n_rows, cudaStream_t st, double step) {
    using C = WSCfg<D, JT>;
    auto kern = ddc_fused_ws_kernel<D, JT>;
    static bool attr_set[64] = {};
    if (h->device < 64 && !attr_set[h->device]) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
        // two CTAs per SM need the whole shared-memory carve-out (the driver's default follows ONE CTA's request)
        if (C::CTAS > 1) CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        attr_set[h->device] = true;
    }
    p.cps_magic = p.tiles_per_stream > 1 ? ~0ull / (unsigned long long)p.tiles_per_stream + 1ull : 0ull;   // chunk_of()
    CUtensorMap tmap;
    int rc = make_slice_tmap(d_in, D, C::WHOLE, C::SLOT_ROWS, n_rows, p.n_streams, p.in_stride, &tmap);
    if (rc) return rc;
    TapsParam<C::NTW> tp;
    std::memcpy(tp.c2, cached_wtaps(h, step, JT, D), sizeof(float2) * (size_t)C::NTW);
    const long long grid = std::min<long long>(p.total_tiles, (long long)h->sm_count * C::CTAS);
    kern<<<(unsigned)grid, C::NWARPS * 32 + 32 * C::NPROD, C::SMEM, st>>>(p, tmap, tp);
    CUDA_TRY(cudaGetLastError());
    h->launches++;
    char name[96];
    snprintf(name, sizeof(name), "fused_fast_fir_%s<D%d,R%d,J%d,SLICES%d,SLOTS%d>", D >= 32 ? "sliced" : (C::WHOLE ? "row_staged" : "tensor_staged"), D, C::R, JT, C::LPQ,
             C::NSLOT);
    h->last_variant = name;
    return DDCB200_OK;
}

// tap-block counts beyond 32 come in multiples of 16 (one more pass and two more halo rows each), up to WSCfg::JT_MAX
template <int D, int JT>
int launch_ws_from(ddcb200* h, RunParams& p, const float* d_in, long long n_rows, cudaStream_t st, double step, int jt) {
    if (jt == JT) return launch_ws<D, JT>(h, p, d_in, n_rows, st, step);
    if constexpr (JT + 16 <= WSCfg<D, 8>::JT_MAX) return launch_ws_from<D, JT + 16>(h, p, d_in, n_rows, st, step, jt);
    return fail(DDCB200_EINVAL, "tensor-staged kernel: unsupported tap-block count %d at D = %d", jt, D);
}

template <int D>
int launch_ws_j(ddcb200* h, RunParams& p, const float* d_in, long long n_rows, cudaStream_t st, double step, int jt) {
    switch (jt) {
        case 8: return launch_ws<D, 8>(h, p, d_in, n_rows, st, step);
        case 16: return launch_ws<D, 16>(h, p, d_in, n_rows, st, step);
        default: return launch_ws_from<D, 32>(h, p, d_in, n_rows, st, step, jt);
    }
}

template <int D>
int launch_p_j(ddcb200* h, RunParams& p, const float2* ct, cudaStream_t st, int jt, int ks) {
    if (ks == 2) {
        switch (jt) {
            case 4: return launch_p<D, 4, 2>(h, p, ct, st);
            case 8: return launch_p<D, 8, 2>(h, p, ct, st);
            default: return launch_p<D, 16, 2>(h, p, ct, st);
        }
    }
    switch (jt) {
        case 4: return launch_p<D, 4, 1>(h, p, ct, st);
        case 8: return launch_p<D, 8, 1>(h, p, ct, st);
        default: return launch_p<D, 16, 1>(h, p, ct, st);
    }
}

// Core dispatcher for device-resident data.
int run_device(ddcb200* h, const void* d_in, bool packed, int64_t n_samples, int64_t n_streams, int64_t in_stride,
               double step, int64_t sample_offset, ddcb200_c64* d_out, int64_t out_stride, cudaStream_t st,
               int64_t m_limit = -1) {
    const int T = (int)h->taps.size();
    const int D = h->decim;
    if (!d_in || !d_out) return fail(DDCB200_EINVAL, "null device pointer");
    if (n_streams <= 0 || n_samples <= 0) return fail(DDCB200_EINVAL, "n_samples and n_streams must be positive");
    if (n_samples < T) return fail(DDCB200_ETOOSHORT, "n_samples (%lld) < n_taps (%d)", (long long)n_samples, T);
    if (packed && (n_samples % 4)) return fail(DDCB200_EINVAL, "packed input needs n_samples %% 4 == 0");
    if (n_streams > 65535) return fail(DDCB200_EINVAL, "at most 65535 streams per call");
    int64_t M = (n_samples - T) / D + 1;
    if (m_limit >= 0 && M > m_limit) M = m_limit;

    RunParams p{};
    p.in = d_in;
    p.out = reinterpret_cast<float2*>(d_out);
    p.n_samples = n_samples;
    p.in_stride = in_stride;
    p.out_stride = out_stride;
    p.n_out = M;
    p.n_streams = (int)n_streams;
    const double fstep = step - std::floor(step);
    p.step_fx = to_fx64(fstep);
    p.phase0_fx = phase_of(step, sample_offset);
    p.vec_store = ((reinterpret_cast<uintptr_t>(d_out) % 16) == 0 && (out_stride % 2) == 0) ? 1 : 0;
    p.debug_mode = h->debug_mode;
    p.stagger_cycles = h->stagger_cycles;
    p.l2_ahead = h->l2_ahead;
    p.dbg = h->d_dbg;

    // ---- fused path eligibility ---------------------------------------------------------------------------
    int R = 0;
    switch (D) {
        case 4: R = 16; break;
        case 8: R = 8; break;
        case 16: R = 4; break;
        case 32: R = 2; break;
        case 64: R = 1; break;
        default: R = 0;
    }
    long long tiles = 0;
    int n_taps_pad = T, J = 0, halo_rows = 0, ks = 1;
    const bool aligned = !packed && (reinterpret_cast<uintptr_t>(d_in) % 16 == 0) && (in_stride % 4 == 0);
    if (R > 0 && aligned && h->force_variant != 1 && h->force_variant != 5 && h->force_variant != 6) {
        J = (T + D - 1) / D;
        // tap split 2 (16 compute warps) whenever it costs no extra zero taps; option "variant" 2 / 3 force KS 1 / 2
        ks = (J % (2 * R) == 0 && R <= 4) ? 2 : 1;
        if (h->force_variant == 2) ks = 1;
        if (h->force_variant == 3 && R <= 4) ks = 2;
        J = ((J + ks * R - 1) / (ks * R)) * (ks * R);  // each thread's tap-block loop is unrolled R times
        n_taps_pad = J * D;
        // a thread-row reads blocks 0 .. J+R-2 of its own row space -> rows g .. g + (J+R-2)/R
        halo_rows = (J + R - 2) / R;
        const long long tile_out = 256LL * R;
        if (n_taps_pad <= kMaxTapsFused) tiles = (M + tile_out - 1) / tile_out;  // the last one may be ragged
    }

    // ---- packed 10-bit input: phase-major kernel with the unpack fused behind the TMA ring -------------------------
    if (packed && (reinterpret_cast<uintptr_t>(d_in) % 16 == 0) && (in_stride % 16 == 0) && (D == 16 || D == 32 || D == 64) &&
        (T + D - 1) / D <= 16 && h->force_variant != 1) {
        const int Jp = (T + D - 1) / D;
        const int jt = Jp <= 4 ? 4 : (Jp <= 8 ? 8 : 16);
        const long long chunk_out = 32LL * (128 / D);
        p.tiles_per_stream = (M + chunk_out - 1) / chunk_out;
        p.total_tiles = p.tiles_per_stream * n_streams;
        p.n_taps = jt * D;
        p.n_tap_blocks = jt;
        p.m_begin = 0;
        // fast-FIR variant where a thread has R = 8 outputs (D = 16); option "variant" 7 forces it, 5 forces the direct form
        // warp-specialised variant (unpack warps + FIR warps): 1.10 against 1.13 ms on 64 x 2^24 samples; default where it is
        // instantiated (D = 16, 129 .. 256 taps); option "variant" 7 forces the fused-unpack kernel, 10 this one
        if (D == 16 && jt == 16 && (h->force_variant == 10 || h->force_variant == 0)) return launch_w10s<16, 16>(h, p, st, step);
        if ((D == 16 && h->force_variant != 5) || h->force_variant == 7) {
            switch (D) {
                case 16: return launch_w10_j<16>(h, p, st, step, jt);
                case 32: return launch_w10_j<32>(h, p, st, step, jt);
                default: return launch_w10_j<64>(h, p, st, step, jt);
            }
        }
        std::vector<float2> ctp((size_t)jt * D);
        make_ctaps(h, step, jt * D, ctp.data());
        switch (D) {
            case 16: return launch_p10_j<16>(h, p, ctp.data(), st, jt);
            case 32: return launch_p10_j<32>(h, p, ctp.data(), st, jt);
            default: return launch_p10_j<64>(h, p, ctp.data(), st, jt);
        }
    }

    // ---- packed input without a fused-unpack kernel for this (T, D): unpack into a float32 workspace, then the float32 path --
    if (packed && h->force_variant != 1) {
        const long long pitch = (n_samples + 3) / 4 * 4;
        const size_t need = (size_t)pitch * (size_t)n_streams;
        if (need > h->unpack_ws_cap) {
            CUDA_TRY(cudaStreamSynchronize(st));   // a previous launch may still read the old workspace
            if (h->unpack_ev) CUDA_TRY(cudaEventSynchronize(h->unpack_ev));
            if (h->d_unpack_ws) cudaFree(h->d_unpack_ws);
            h->d_unpack_ws = nullptr;
            h->unpack_ws_cap = 0;
            if (cudaMalloc(&h->d_unpack_ws, need * sizeof(float)) != cudaSuccess) {
                cudaGetLastError();
                return fail(DDCB200_ENOMEM, "packed input: cannot allocate a %zu-byte unpack workspace", need * sizeof(float));
            }
            h->unpack_ws_cap = need;
        }
        // the workspace is shared by every call on this handle, whatever stream it comes on: order this unpack behind the
        // last kernel that read it
        if (!h->unpack_ev) CUDA_TRY(cudaEventCreateWithFlags(&h->unpack_ev, cudaEventDisableTiming));
        else CUDA_TRY(cudaStreamWaitEvent(st, h->unpack_ev, 0));
        const long long groups = n_samples / 4;
        dim3 grid((unsigned)((groups + 255) / 256), (unsigned)n_streams);
        unpack10_rows_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const uint8_t*>(d_in), in_stride, groups, h->d_unpack_ws, pitch);
        CUDA_TRY(cudaGetLastError());
        h->launches++;
        int rc2 = run_device(h, h->d_unpack_ws, false, n_samples, n_streams, pitch, step, sample_offset, d_out, out_stride, st, m_limit);
        CUDA_TRY(cudaEventRecord(h->unpack_ev, st));
        if (rc2) return rc2;
        h->last_variant = "unpack10+" + h->last_variant;
        return DDCB200_OK;
    }

    // ---- large decimations (D = 32, 64): sliced staging (ddc_kernel_ws.cuh), fast FIR with R = 8 outputs per thread ----------
    // Auto: where the direct form is FP32-bound (4 T / D flop per sample against 4 + 8 / D bytes at the measured ridge of
    // 11.4 flop/B); HBM-bound cells stay on the phase-major kernel, which over-fetches nothing.  Option "variant" 11 forces it.
    long long m_done = 0;
    bool sliced_done = false;
    {
        const int Jp = (T + D - 1) / D;
        const bool fp32_bound = 4.0 * T / D > 11.4 * (4.0 + 8.0 / D);
        // D = 4 / 8 / 16: the same tensor-staged kernel with whole blocks (R = 8 outputs per thread), ddc_kernel_ws.cuh.
        //   D = 4 / 8, up to 1024 taps: two CTAs per SM (sixteen compute warps; WSCfg::CTAS) -- measured at N = 2^26 against the tile
        //   kernel: D = 4: T = 64 0.085 vs 0.128 ms, 128 0.126 vs 0.173, 256 0.234 vs 0.274, 512 0.456 vs 0.503, 1024 0.895 vs 0.981;
        //   D = 8: T = 64 0.054 vs 0.062, 128 0.070 vs 0.083, 256 0.122 vs 0.141, 512 0.229 vs 0.265, 1024 0.449 vs 0.495 (sub-filter
        //   kernel).  Padded taps <= 64 use whole-row tiles (WSCfg::WHOLE).
        //   D = 16: whole-row tiles for the HBM-bound filters only (T <= 128: 0.0517 against 0.0524 / 0.0537 ms of the 1-D bulk-copy
        //   kernel); longer filters are much slower here than in ddc_kernel_w.cuh (T = 256: 0.312 vs 0.244 ms at 2^28): option 11.
        // The kernel pads the filter to 8, 16 or a multiple of 16 tap blocks, the tile kernel to a multiple of its R (16 / 8 at
        // D = 4 / 8): auto only where the padded work is within 12 % of the tile kernel's (its deficit there: 84-94 % against 99 %).
        const int jt_ws = Jp <= 8 ? 8 : (Jp + 15) / 16 * 16;
        const int j_tile = R > 0 ? (Jp + R - 1) / R * R : Jp;
        const bool pad_ok = jt_ws * 100 <= j_tile * 112;
        const bool small_auto = (D == 4 && T <= 1024 && pad_ok) || (D == 8 && T <= 1024 && pad_ok) || (D == 16 && T <= 128);
        const bool small_d = (D == 4 || D == 8 || D == 16) && (h->force_variant == 11 || (h->force_variant == 0 && small_auto));
        if (aligned_f32(d_in, in_stride, packed) && Jp <= (D == 4 ? 256 : (D == 8 ? 128 : 32)) && T >= D &&
            (small_d || ((D == 32 || D == 64) && (h->force_variant == 11 || (h->force_variant == 0 && fp32_bound))))) {
            const int jt = jt_ws;
            const long long n_blocks = (n_samples / (8LL * D)) * 8;         // whole thread-rows of 8 blocks: the tensor map covers exactly these
            long long m_f = n_blocks * D >= T ? (n_blocks * D - T) / D + 1 : 0;   // outputs whose window lies inside them
            if (m_f > M) m_f = M;
            // chunk_of() divides by multiplication: exact while total chunks x chunks per stream < 2^64
            if ((double)((m_f + 255) / 256) * (double)((m_f + 255) / 256) * (double)n_streams >= 1.8e19) m_f = 0;
            if (m_f > 0) {
                const long long m_all = p.n_out;
                p.n_out = m_f;
                p.tiles_per_stream = (m_f + 255) / 256;
                p.total_tiles = p.tiles_per_stream * n_streams;
                p.n_taps = jt * D;
                p.n_tap_blocks = jt;
                p.m_begin = 0;
                const float* fin = reinterpret_cast<const float*>(d_in);
                int rc2 = D == 4    ? launch_ws_j<4>(h, p, fin, n_blocks / 8, st, step, jt)
                          : D == 8  ? launch_ws_j<8>(h, p, fin, n_blocks / 8, st, step, jt)
                          : D == 16 ? launch_ws_j<16>(h, p, fin, n_blocks / 8, st, step, jt)
                          : D == 32 ? launch_ws_j<32>(h, p, fin, n_blocks / 8, st, step, jt)
                                    : launch_ws_j<64>(h, p, fin, n_blocks / 8, st, step, jt);
                if (rc2) return rc2;
                p.n_out = m_all;
                m_done = m_f;
                sliced_done = true;
            }
        }
    }

    // ---- small decimations (D = 4, 8): NQ = 16 / D interleaved decimate-by-16 fast FIRs with shifted tap sets --------------
    if (!sliced_done && aligned_f32(d_in, in_stride, packed) && (D == 4 || D == 8) && (h->force_variant == 0 || h->force_variant == 7)) {
        const int nq = 16 / D;
        const int Tq = T + D * (nq - 1);
        const int jneed = (Tq + 15) / 16;
        const int jt = (nq == 2 && jneed > 34 && jneed <= 66) ? 66 : 0;   // see launch_wq_j
        if (jt) {
            const long long mq = (M + nq - 1) / nq;           // decimate-by-16 outputs per tap set
            p.tiles_per_stream = (mq + 255) / 256;
            p.total_tiles = p.tiles_per_stream * n_streams;
            p.n_taps = jt * 16;
            p.n_tap_blocks = jt;
            p.m_begin = 0;
            return nq == 2 ? launch_wq_j<2>(h, p, st, step, jt) : launch_wq_j<4>(h, p, st, step, jt);
        }
    }

    // ---- kernel P (phase-major, R = 128/D outputs per thread): short polyphase branches, J = ceil(T/D) <= 16 ------
    const bool long_w = D == 16 && (T + D - 1) / D > 16 && (T + D - 1) / D <= 64 &&
                        (h->force_variant == 0 || h->force_variant == 7 || h->force_variant == 9);
    if (!sliced_done && aligned_f32(d_in, in_stride, packed) && (D == 16 || D == 32 || D == 64) && ((T + D - 1) / D <= 16 || long_w) &&
        (h->force_variant == 0 || (h->force_variant >= 5 && h->force_variant <= 9) || h->force_variant == 12)) {
        const int ksp = (h->force_variant == 6) ? 2 : 1;   // option "variant": 5 (= auto) one warp per chunk, 6 = two (slower)
        const int Jp = (T + D - 1) / D;
        const int jt = Jp <= 4 ? 4 : (Jp <= 8 ? 8 : (Jp <= 16 ? 16 : (Jp <= 32 ? 32 : 64)));
        const long long chunk_out = 32LL * (128 / D);   // PCfg::CHUNK_OUT
        p.tiles_per_stream = (M + chunk_out - 1) / chunk_out;
        p.total_tiles = p.tiles_per_stream * n_streams;
        p.n_taps = jt * D;
        p.n_tap_blocks = jt;
        p.m_begin = 0;
        // option "variant": 0 auto; 5 / 6 kernel P with one / two warps per chunk; 7 fast FIR (kernel W); 8 deferred-epilogue P.
        // Auto picks the fast-FIR kernel where a thread has R = 8 outputs (D = 16) and the output rows allow 16-byte stores.
        if (h->force_variant == 12 && D == 16 && jt == 16) return launch_w2x<16, 16>(h, p, st, step);   // 16 compute warps (experiment)
        const bool want_w = h->force_variant == 7 || h->force_variant == 9 || (h->force_variant == 0 && D == 16);
        if (want_w) {   // any complex64-aligned output: the epilogue picks its 16-byte pairing per thread
            const bool nest = h->force_variant == 9;   // option "variant" 9: two nested fast-FIR levels (D = 16)
            switch (D) {
                case 16: return launch_w_j<16>(h, p, st, step, jt, nest);
                case 32: return launch_w_j<32>(h, p, st, step, jt, nest);
                default: return launch_w_j<64>(h, p, st, step, jt, nest);
            }
        }
        std::vector<float2> ctp((size_t)jt * D);
        make_ctaps(h, step, jt * D, ctp.data());
        if (h->force_variant == 0 || h->force_variant == 8) {   // deferred-epilogue variant
            switch (D) {
                case 16: return launch_pd_j<16>(h, p, ctp.data(), st, jt);
                case 32: return launch_pd_j<32>(h, p, ctp.data(), st, jt);
                default: return launch_pd_j<64>(h, p, ctp.data(), st, jt);
            }
        }
        switch (D) {
            case 16: return launch_p_j<16>(h, p, ctp.data(), st, jt, ksp);
            case 32: return launch_p_j<32>(h, p, ctp.data(), st, jt, ksp);
            default: return launch_p_j<64>(h, p, ctp.data(), st, jt, ksp);
        }
    }

    std::vector<float2> ct((size_t)std::max(n_taps_pad, T));
    make_ctaps(h, step, (int)ct.size(), ct.data());

    int rc = DDCB200_OK;
    if (tiles > 0 && !sliced_done) {
        p.tiles_per_stream = tiles;
        p.total_tiles = tiles * n_streams;
        p.n_taps = n_taps_pad;
        p.n_tap_blocks = J;
        p.halo_rows = halo_rows;
        p.m_begin = 0;
        const int grid_limit = h->sm_count;
        switch (D) {
            case 4: rc = launch_fused_t<4, 16>(h, p, ct.data(), st, grid_limit, ks); break;
            case 8: rc = launch_fused_t<8, 8>(h, p, ct.data(), st, grid_limit, ks); break;
            case 16: rc = launch_fused_t<16, 4>(h, p, ct.data(), st, grid_limit, ks); break;
            case 32: rc = launch_fused_t<32, 2>(h, p, ct.data(), st, grid_limit, ks); break;
            case 64: rc = launch_fused_t<64, 1>(h, p, ct.data(), st, grid_limit, ks); break;
        }
        if (rc) return rc;
        m_done = M;
    }
    if (m_done < M) {
        // stream tails / everything the fused path does not cover
        rc = ensure_ring(h, T);
        if (rc) return rc;
        const int slot = h->ring_pos;
        h->ring_pos = (h->ring_pos + 1) % ddcb200::kRing;
        CUDA_TRY(cudaEventSynchronize(h->ring_ev[slot]));  // previous user of this slot has consumed it
        std::memcpy(h->h_ctaps[slot], ct.data(), sizeof(float2) * T);
        CUDA_TRY(cudaMemcpyAsync(h->d_ctaps[slot], h->h_ctaps[slot], sizeof(float2) * T, cudaMemcpyHostToDevice, st));
        p.n_taps = T;
        p.m_begin = m_done;
        const long long count = M - m_done;
        dim3 grid((unsigned)((count + 127) / 128), (unsigned)n_streams);
        if (packed)
            ddc_generic_kernel<true><<<grid, 128, 0, st>>>(p, h->d_ctaps[slot], D);
        else
            ddc_generic_kernel<false><<<grid, 128, 0, st>>>(p, h->d_ctaps[slot], D);
        CUDA_TRY(cudaGetLastError());
        CUDA_TRY(cudaEventRecord(h->ring_ev[slot], st));
        h->launches++;
        if (tiles == 0 && !sliced_done) h->last_variant = packed ? "generic<packed10>" : "generic<f32>";
    }
    return DDCB200_OK;
}

int ensure_chunks(ddcb200* h, size_t in_bytes, size_t out_elems) {
    if (in_bytes > h->chunk_in_cap) {
        h->chunk_in_cap = 0;   // only valid again once every buffer has been allocated
        for (int i = 0; i < ddcb200::kBufs; ++i) {
            if (h->d_chunk_in[i]) cudaFree(h->d_chunk_in[i]);
            h->d_chunk_in[i] = nullptr;
            CUDA_TRY(cudaMalloc(&h->d_chunk_in[i], in_bytes));
        }
        h->chunk_in_cap = in_bytes;
    }
    if (out_elems > h->chunk_out_cap) {
        h->chunk_out_cap = 0;
        for (int i = 0; i < ddcb200::kBufs; ++i) {
            if (h->d_chunk_out[i]) cudaFree(h->d_chunk_out[i]);
            h->d_chunk_out[i] = nullptr;
            CUDA_TRY(cudaMalloc(&h->d_chunk_out[i], out_elems * sizeof(ddcb200_c64)));
        }
        h->chunk_out_cap = out_elems;
    }
    for (int i = 0; i < ddcb200::kBufs; ++i) {
        if (!h->ev_in[i]) CUDA_TRY(cudaEventCreateWithFlags(&h->ev_in[i], cudaEventDisableTiming));
        if (!h->ev_k[i]) CUDA_TRY(cudaEventCreateWithFlags(&h->ev_k[i], cudaEventDisableTiming));
        if (!h->ev_out[i]) CUDA_TRY(cudaEventCreateWithFlags(&h->ev_out[i], cudaEventDisableTiming));
    }
    return DDCB200_OK;
}

// H2D copy from PAGEABLE host memory: cudaMemcpyAsync would fall back to the driver's single-threaded staging (about
// 12 GB/s).  Instead the bytes go through two pinned 8 MB staging buffers filled by a few host threads (memcpy is
// memory-bandwidth bound, one core does not saturate it) while the previous buffer is in flight on the copy engine.
bool is_pageable(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return a.type == cudaMemoryTypeUnregistered;
}

int staged_h2d(ddcb200* h, void* d_dst, const void* h_src, size_t bytes, cudaStream_t st) {
    for (int i = 0; i < ddcb200::kStage; ++i) {
        if (!h->h_stage[i]) CUDA_TRY(cudaMallocHost(&h->h_stage[i], ddcb200::kStageBytes));
        if (!h->ev_stage[i]) CUDA_TRY(cudaEventCreateWithFlags(&h->ev_stage[i], cudaEventDisableTiming));
    }
    const int nt = std::max(1, std::min(h->copy_threads, 16));
    size_t done = 0;
    while (done < bytes) {
        const size_t n = std::min(bytes - done, ddcb200::kStageBytes);
        const int b = h->stage_pos;
        h->stage_pos = (h->stage_pos + 1) % ddcb200::kStage;
        CUDA_TRY(cudaEventSynchronize(h->ev_stage[b]));   // the copy engine has drained this staging buffer
        char* dst = static_cast<char*>(h->h_stage[b]);
        const char* src = static_cast<const char*>(h_src) + done;
        if (nt == 1 || n < (1u << 20)) {
            std::memcpy(dst, src, n);
        } else {
            const size_t per = ((n + nt - 1) / nt + 63) & ~(size_t)63;
            std::vector<std::thread> th;
            for (int t = 1; t < nt; ++t) {
                const size_t o = per * t;
                if (o < n) th.emplace_back([=] { std::memcpy(dst + o, src + o, std::min(per, n - o)); });
            }
            std::memcpy(dst, src, std::min(per, n));
            for (auto& t : th) t.join();
        }
        CUDA_TRY(cudaMemcpyAsync(static_cast<char*>(d_dst) + done, dst, n, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaEventRecord(h->ev_stage[b], st));
        done += n;
    }
    return DDCB200_OK;
}

// complex64 -> complex128 on a few host threads, straight into the caller's (usually fresh, untouched) array: the page
// faults of the first touch are spread over the threads as well
void widen_c64_to_c128(const ddcb200_c64* src, double* dst, size_t n, int nt) {
    auto work = [=](size_t a, size_t b) {
        for (size_t i = a; i < b; ++i) {
            dst[2 * i] = (double)src[i].re;
            dst[2 * i + 1] = (double)src[i].im;
        }
    };
    nt = std::max(1, std::min(nt, 16));
    if (nt == 1 || n < (1u << 16)) {
        work(0, n);
        return;
    }
    const size_t per = (n + nt - 1) / nt;
    std::vector<std::thread> th;
    for (int t = 1; t < nt; ++t)
        if (per * t < n) th.emplace_back(work, per * t, std::min(n, per * (t + 1)));
    work(0, std::min(per, n));
    for (auto& t : th) t.join();
}

// Host path: every stream is cut into time chunks of `chunk_samples` (+ T-D halo); chunk c of all streams goes
// H2D on copy_in, through the fused kernel on `stream`, and D2H on copy_out, three buffers deep.
int run_host(ddcb200* h, const void* h_in, bool packed, int64_t n_samples, int64_t n_streams, int64_t in_stride,
             double step, int64_t sample_offset, ddcb200_c64* h_out, int64_t out_stride, double* h_out128 = nullptr) {
    // h_out128 (one stream only): the result is delivered as complex128 (the reference's dtype) -- every chunk lands in a
    // pinned buffer and is widened on the host threads while the next chunk is in flight
    if (h_out128) h_out = reinterpret_cast<ddcb200_c64*>(h_out128);   // only for the null check below
    const int T = (int)h->taps.size();
    const int D = h->decim;
    if (!h_in || !h_out) return fail(DDCB200_EINVAL, "null host pointer");
    if (n_streams <= 0 || n_samples <= 0) return fail(DDCB200_EINVAL, "n_samples and n_streams must be positive");
    if (n_samples < T) return fail(DDCB200_ETOOSHORT, "n_samples (%lld) < n_taps (%d)", (long long)n_samples, T);
    if (packed && (n_samples % 4)) return fail(DDCB200_EINVAL, "packed input needs n_samples %% 4 == 0");
    const int64_t M = (n_samples - T) / D + 1;
    // chunk length in outputs; inputs of a chunk start at a multiple of 4*D samples so that packed chunks start on a
    // byte boundary and float chunks stay 16-byte aligned
    int64_t per_stream = std::max<int64_t>(h->chunk_samples / n_streams, (int64_t)4 * T);
    int64_t m_chunk = std::max<int64_t>(per_stream / D, 1);
    m_chunk = ((m_chunk + 63) / 64) * 64;   // chunk starts stay 16-byte aligned for float32 AND packed (64 samples = 80 B)
    const int64_t n_chunks = (M + m_chunk - 1) / m_chunk;
    const int64_t in_chunk_samples = ((m_chunk - 1) * D + T + 63) / 64 * 64;
    const size_t in_elem_bytes_num = packed ? 5 : 16, in_elem_den = 4;  // bytes per 4 samples
    const size_t in_row_bytes = (size_t)in_chunk_samples / in_elem_den * in_elem_bytes_num;
    int rc = ensure_chunks(h, in_row_bytes * (size_t)n_streams + 64, (size_t)m_chunk * (size_t)n_streams);
    if (rc) return rc;
    DrainGuard drain(h);
    const bool pageable_in = h->copy_threads > 0 && is_pageable(h_in);
    if (h_out128) {
        if (n_streams != 1) return fail(DDCB200_EINVAL, "complex128 host output is for one stream per call");
        if ((size_t)m_chunk > h->ostage_cap) {
            for (int i = 0; i < ddcb200::kBufs; ++i) {
                if (h->h_ostage[i]) cudaFreeHost(h->h_ostage[i]);
                h->h_ostage[i] = nullptr;
                CUDA_TRY(cudaMallocHost(&h->h_ostage[i], (size_t)m_chunk * sizeof(ddcb200_c64)));
            }
            h->ostage_cap = (size_t)m_chunk;
        }
    }
    int64_t pend_m0 = -1, pend_mc = 0;   // chunk whose outputs still have to be widened
    int pend_b = 0;
    auto flush_pending = [&]() -> int {
        if (pend_m0 < 0) return DDCB200_OK;
        CUDA_TRY(cudaEventSynchronize(h->ev_out[pend_b]));
        widen_c64_to_c128(h->h_ostage[pend_b], h_out128 + 2 * pend_m0, (size_t)pend_mc, std::max(1, h->copy_threads));
        pend_m0 = -1;
        return DDCB200_OK;
    };

    for (int64_t c = 0; c < n_chunks; ++c) {
        const int b = (int)(c % ddcb200::kBufs);
        const int64_t m0 = c * m_chunk;
        const int64_t mc = std::min<int64_t>(m_chunk, M - m0);
        const int64_t n0 = m0 * D;
        const int64_t nc = (mc - 1) * D + T;                  // samples this chunk needs
        const int64_t nc4 = std::min<int64_t>((nc + 3) / 4 * 4, n_samples - n0);  // copy whole groups when available
        const size_t row_bytes = packed ? (size_t)(nc4 / 4 * 5) : (size_t)nc4 * 4;
        const size_t src_off = packed ? (size_t)(n0 / 4 * 5) : (size_t)n0 * 4;
        const size_t src_pitch = packed ? (size_t)in_stride : (size_t)in_stride * 4;
        // buffer b is free once the D2H of chunk c - kBufs finished (ev_out) -- wait on the copy-in stream
        if (c >= ddcb200::kBufs) CUDA_TRY(cudaStreamWaitEvent(h->copy_in, h->ev_k[b], 0));
        if (n_streams == 1 && pageable_in && row_bytes >= (16u << 20)) {   // below that the thread start-up costs more than it saves
            rc = staged_h2d(h, h->d_chunk_in[b], reinterpret_cast<const char*>(h_in) + src_off, row_bytes, h->copy_in);
            if (rc) return rc;
        } else {
            CUDA_TRY(copy_rows_async(h->d_chunk_in[b], in_row_bytes, reinterpret_cast<const char*>(h_in) + src_off, src_pitch,
                                     row_bytes, (size_t)n_streams, cudaMemcpyHostToDevice, h->copy_in));
        }
        CUDA_TRY(cudaEventRecord(h->ev_in[b], h->copy_in));
        CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_in[b], 0));
        if (c >= ddcb200::kBufs) CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_out[b], 0));
        const int64_t n_dev = packed ? (nc4 / 4 * 4) : nc;
        rc = run_device(h, h->d_chunk_in[b], packed, packed ? nc4 : n_dev, n_streams,
                        packed ? (int64_t)in_row_bytes : (int64_t)(in_row_bytes / 4), step, sample_offset + n0,
                        h->d_chunk_out[b], m_chunk, h->stream, mc);
        if (rc) return rc;
        CUDA_TRY(cudaEventRecord(h->ev_k[b], h->stream));
        CUDA_TRY(cudaStreamWaitEvent(h->copy_out, h->ev_k[b], 0));
        if (h_out128) {
            // the landing buffer of this chunk buffer was widened before its kernel was queued (kBufs = 3 chunks ago at the
            // latest: flush_pending runs one chunk behind)
            CUDA_TRY(cudaMemcpyAsync(h->h_ostage[b], h->d_chunk_out[b], (size_t)mc * sizeof(ddcb200_c64), cudaMemcpyDeviceToHost,
                                     h->copy_out));
            CUDA_TRY(cudaEventRecord(h->ev_out[b], h->copy_out));
            rc = flush_pending();   // the previous chunk, while this one is in flight
            if (rc) return rc;
            pend_m0 = m0;
            pend_mc = mc;
            pend_b = b;
        } else {
            CUDA_TRY(copy_rows_async(h_out + m0, (size_t)out_stride * sizeof(ddcb200_c64), h->d_chunk_out[b],
                                     (size_t)m_chunk * sizeof(ddcb200_c64), (size_t)mc * sizeof(ddcb200_c64), (size_t)n_streams,
                                     cudaMemcpyDeviceToHost, h->copy_out));
            CUDA_TRY(cudaEventRecord(h->ev_out[b], h->copy_out));
        }
    }
    rc = flush_pending();
    if (rc) return rc;
    CUDA_TRY(cudaStreamSynchronize(h->copy_out));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->copy_in));
    return DDCB200_OK;
}

}  // namespace

// ================================================================================================================
extern "C" {

int ddcb200_version(void) { return DDCB200_VERSION; }
const char* ddcb200_last_error(void) { return g_err; }

int64_t ddcb200_out_len(int64_t n_samples, int n_taps, int decimation) {
    if (n_samples <= 0 || n_taps <= 0 || decimation <= 0) return 0;
    const int64_t full = (n_samples >= n_taps ? n_samples - n_taps : n_taps - n_samples) + 1;
    return (full + decimation - 1) / decimation;
}

int ddcb200_set_taps(ddcb200_t* h, const double* taps, int n_taps) {
    if (!h || !taps || n_taps <= 0) return fail(DDCB200_EINVAL, "set_taps: bad arguments");
    double s = 0.0;
    for (int i = 0; i < n_taps; ++i) s += taps[i];  // same left-to-right float64 sum as Python's sum() (ddc.py:98)
    if (!(s != 0.0) || !std::isfinite(s)) return fail(DDCB200_EINVAL, "set_taps: sum of taps is %g", s);
    h->taps.assign(taps, taps + n_taps);
    h->taps_sum = s;
    h->wt_jt = 0;   // invalidate the folded-tap caches
    h->wq_jt = 0;
    return DDCB200_OK;
}

int ddcb200_set_decimation(ddcb200_t* h, int decimation) {
    if (!h || decimation <= 0) return fail(DDCB200_EINVAL, "set_decimation: bad arguments");
    h->decim = decimation;
    return DDCB200_OK;
}

int ddcb200_create(ddcb200_t** handle, int device, const double* taps, int n_taps, int decimation) {
    if (!handle) return fail(DDCB200_EINVAL, "create: null handle pointer");
    *handle = nullptr;
    if (!taps || n_taps <= 0 || decimation <= 0) return fail(DDCB200_EINVAL, "create: bad taps/decimation");
    int ndev = 0;
    CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(DDCB200_EINVAL, "create: device %d of %d", device, ndev);
    DeviceGuard g(device);
    if (!g.ok) return fail(DDCB200_ECUDA, "create: cannot select device %d", device);
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(DDCB200_ECUDA, "create: device %d is sm_%d%d; this library is sm_100a only", device, prop.major, prop.minor);
    ddcb200* h = new (std::nothrow) ddcb200();
    if (!h) return fail(DDCB200_ENOMEM, "create: out of memory");
    h->device = device;
    h->sm_count = prop.multiProcessorCount;
    int rc = ddcb200_set_taps(h, taps, n_taps);
    if (!rc) rc = ddcb200_set_decimation(h, decimation);
    if (rc) {
        delete h;
        return rc;
    }
    cudaError_t e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->copy_in, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->copy_out, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
        delete h;
        return fail(DDCB200_ECUDA, "create: stream creation failed: %s", cudaGetErrorString(e));
    }
    *handle = h;
    return DDCB200_OK;
}

void ddcb200_destroy(ddcb200_t* h) {
    if (!h) return;
    DeviceGuard g(h->device);
    cudaStreamSynchronize(h->stream);
    for (int i = 0; i < ddcb200::kRing; ++i) {
        if (h->d_ctaps[i]) cudaFree(h->d_ctaps[i]);
        if (h->h_ctaps[i]) cudaFreeHost(h->h_ctaps[i]);
        if (h->ring_ev[i]) cudaEventDestroy(h->ring_ev[i]);
    }
    for (int i = 0; i < ddcb200::kStage; ++i) {
        if (h->h_stage[i]) cudaFreeHost(h->h_stage[i]);
        if (h->ev_stage[i]) cudaEventDestroy(h->ev_stage[i]);
    }
    for (int i = 0; i < ddcb200::kBufs; ++i)
        if (h->h_ostage[i]) cudaFreeHost(h->h_ostage[i]);
    if (h->d_unpack_ws) cudaFree(h->d_unpack_ws);
    if (h->unpack_ev) cudaEventDestroy(h->unpack_ev);
    for (int i = 0; i < ddcb200::kBufs; ++i) {
        if (h->d_chunk_in[i]) cudaFree(h->d_chunk_in[i]);
        if (h->d_chunk_out[i]) cudaFree(h->d_chunk_out[i]);
        if (h->ev_in[i]) cudaEventDestroy(h->ev_in[i]);
        if (h->ev_k[i]) cudaEventDestroy(h->ev_k[i]);
        if (h->ev_out[i]) cudaEventDestroy(h->ev_out[i]);
    }
    if (h->d_dbg) cudaFree(h->d_dbg);
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->copy_in) cudaStreamDestroy(h->copy_in);
    if (h->copy_out) cudaStreamDestroy(h->copy_out);
    delete h;
}

int ddcb200_run_f32(ddcb200_t* h, const float* d_in, int64_t n_samples, int64_t n_streams, int64_t in_stride,
                    double step, int64_t sample_offset, ddcb200_c64* d_out, int64_t out_stride, void* cuda_stream) {
    if (!h) return fail(DDCB200_EINVAL, "null handle");
    DeviceGuard g(h->device);
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : h->stream;
    return run_device(h, d_in, false, n_samples, n_streams, in_stride, step, sample_offset, d_out, out_stride, st);
}

int ddcb200_run_packed10(ddcb200_t* h, const uint8_t* d_in, int64_t n_samples, int64_t n_streams, int64_t in_stride_bytes,
                         double step, int64_t sample_offset, ddcb200_c64* d_out, int64_t out_stride, void* cuda_stream) {
    if (!h) return fail(DDCB200_EINVAL, "null handle");
    DeviceGuard g(h->device);
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : h->stream;
    return run_device(h, d_in, true, n_samples, n_streams, in_stride_bytes, step, sample_offset, d_out, out_stride, st);
}

int ddcb200_unpack10(ddcb200_t* h, const uint8_t* d_in, int64_t n_samples, int16_t* o16, float* of32, void* cuda_stream) {
    if (!h || !d_in) return fail(DDCB200_EINVAL, "unpack10: bad arguments");
    if (n_samples < 0 || (n_samples % 4)) return fail(DDCB200_EINVAL, "unpack10: n_samples must be a multiple of 4");
    if (n_samples == 0) return DDCB200_OK;
    DeviceGuard g(h->device);
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : h->stream;
    const long long groups = n_samples / 4;
    unpack10_kernel<<<(unsigned)((groups + 255) / 256), 256, 0, st>>>(d_in, groups, o16, of32);
    CUDA_TRY(cudaGetLastError());
    h->launches++;
    return DDCB200_OK;
}

int ddcb200_pack10(ddcb200_t* h, const float* d_in, int64_t n_samples, int64_t n_streams, int64_t in_stride, uint8_t* d_out,
                   int64_t out_stride_bytes, void* cuda_stream) {
    if (!h || !d_in || !d_out) return fail(DDCB200_EINVAL, "pack10: bad arguments");
    if (n_samples <= 0 || (n_samples % 4) || n_streams <= 0 || n_streams > 65535)
        return fail(DDCB200_EINVAL, "pack10: n_samples must be a positive multiple of 4, 1 .. 65535 streams");
    if (in_stride < n_samples || out_stride_bytes < n_samples / 4 * 5) return fail(DDCB200_EINVAL, "pack10: strides shorter than a row");
    DeviceGuard g(h->device);
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : h->stream;
    const long long groups = n_samples / 4;
    dim3 grid((unsigned)((groups + 255) / 256), (unsigned)n_streams);
    pack10_rows_kernel<<<grid, 256, 0, st>>>(d_in, in_stride, groups, d_out, out_stride_bytes);
    CUDA_TRY(cudaGetLastError());
    h->launches++;
    return DDCB200_OK;
}

int ddcb200_run_short_f32(ddcb200_t* h, const float* d_in, int64_t n_samples, double step, int64_t sample_offset,
                          ddcb200_c64* d_out, void* cuda_stream) {
    if (!h || !d_in || !d_out) return fail(DDCB200_EINVAL, "run_short: bad arguments");
    const int T = (int)h->taps.size();
    if (n_samples <= 0 || n_samples >= T) return fail(DDCB200_EINVAL, "run_short: needs 0 < n_samples < n_taps");
    DeviceGuard g(h->device);
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : h->stream;
    int rc = ensure_ring(h, T);
    if (rc) return rc;
    const int slot = h->ring_pos;
    h->ring_pos = (h->ring_pos + 1) % ddcb200::kRing;
    CUDA_TRY(cudaEventSynchronize(h->ring_ev[slot]));
    float* hh = reinterpret_cast<float*>(h->h_ctaps[slot]);
    for (int i = 0; i < T; ++i) hh[i] = (float)(h->taps[i] / h->taps_sum);
    CUDA_TRY(cudaMemcpyAsync(h->d_ctaps[slot], hh, sizeof(float) * T, cudaMemcpyHostToDevice, st));
    RunParams p{};
    p.in = d_in;
    p.out = reinterpret_cast<float2*>(d_out);
    p.n_samples = n_samples;
    p.n_out = ddcb200_out_len(n_samples, T, h->decim);
    const double fstep = step - std::floor(step);
    p.step_fx = to_fx64(fstep);
    p.phase0_fx = phase_of(step, sample_offset);
    ddc_short_kernel<<<(unsigned)((p.n_out + 63) / 64), 64, 0, st>>>(p, reinterpret_cast<const float*>(h->d_ctaps[slot]), T, h->decim);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(h->ring_ev[slot], st));
    h->launches++;
    h->last_variant = "short<f32>";
    return DDCB200_OK;
}

int ddcb200_mix_f32(ddcb200_t* h, const float* d_x, const ddcb200_c64* d_cw, ddcb200_c64* d_out, int64_t n, void* cuda_stream) {
    if (!h || !d_x || !d_cw || !d_out || n <= 0) return fail(DDCB200_EINVAL, "mix: bad arguments");
    DeviceGuard g(h->device);
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : h->stream;
    mix_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_x, reinterpret_cast<const float2*>(d_cw),
                                                            reinterpret_cast<float2*>(d_out), n);
    CUDA_TRY(cudaGetLastError());
    h->launches++;
    return DDCB200_OK;
}

int ddcb200_fir_c64(ddcb200_t* h, const ddcb200_c64* d_in, int64_t n_in, ddcb200_c64* d_out, void* cuda_stream) {
    if (!h || !d_in || !d_out) return fail(DDCB200_EINVAL, "fir: bad arguments");
    const int T = (int)h->taps.size();
    if (n_in < T) return fail(DDCB200_ETOOSHORT, "fir: n_in (%lld) < n_taps (%d)", (long long)n_in, T);
    DeviceGuard g(h->device);
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : h->stream;
    int rc = ensure_ring(h, T);
    if (rc) return rc;
    const int slot = h->ring_pos;
    h->ring_pos = (h->ring_pos + 1) % ddcb200::kRing;
    CUDA_TRY(cudaEventSynchronize(h->ring_ev[slot]));
    float* hh = reinterpret_cast<float*>(h->h_ctaps[slot]);
    for (int k = 0; k < T; ++k) hh[k] = (float)(h->taps[T - 1 - k] / h->taps_sum);
    CUDA_TRY(cudaMemcpyAsync(h->d_ctaps[slot], hh, sizeof(float) * T, cudaMemcpyHostToDevice, st));
    const long long n_out = n_in - T + 1;
    fir_c64_kernel<<<(unsigned)((n_out + 127) / 128), 128, 0, st>>>(reinterpret_cast<const float2*>(d_in),
                                                                    reinterpret_cast<const float*>(h->d_ctaps[slot]), T,
                                                                    reinterpret_cast<float2*>(d_out), n_out);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaEventRecord(h->ring_ev[slot], st));
    h->launches++;
    return DDCB200_OK;
}

int ddcb200_decimate_c64(ddcb200_t* h, const ddcb200_c64* d_in, int64_t n_in, int64_t offset, ddcb200_c64* d_out,
                         void* cuda_stream) {
    if (!h || !d_in || !d_out || offset < 0) return fail(DDCB200_EINVAL, "decimate: bad arguments");
    if (n_in <= offset) return DDCB200_OK;
    DeviceGuard g(h->device);
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : h->stream;
    const long long n_out = (n_in - offset + h->decim - 1) / h->decim;
    decimate_c64_kernel<<<(unsigned)((n_out + 255) / 256), 256, 0, st>>>(reinterpret_cast<const float2*>(d_in), offset, h->decim,
                                                                         reinterpret_cast<float2*>(d_out), n_out);
    CUDA_TRY(cudaGetLastError());
    h->launches++;
    return DDCB200_OK;
}

int ddcb200_cwg(ddcb200_t* h, void* d_out, int64_t num_samples, int64_t n_streams, int64_t out_stride, int is_complex,
                double cw_scale, double phase_step_cycles, double phase0_cycles, int64_t sample_offset, int noise_mode,
                double noise_scale, uint64_t seed, void* cuda_stream) {
    if (!h || !d_out || num_samples <= 0 || n_streams <= 0 || n_streams > 65535 || out_stride < num_samples)
        return fail(DDCB200_EINVAL, "cwg: bad arguments");
    if (noise_mode < 0 || noise_mode > 2) return fail(DDCB200_EINVAL, "cwg: noise_mode must be 0, 1 or 2");
    DeviceGuard g(h->device);
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : h->stream;
    const unsigned long long step_fx = to_fx64(phase_step_cycles - std::floor(phase_step_cycles));
    const unsigned long long ph0 = to_fx64(phase0_cycles - std::floor(phase0_cycles)) + phase_of(phase_step_cycles, sample_offset);
    dim3 grid((unsigned)((num_samples + 255) / 256), (unsigned)n_streams);
    if (is_complex)
        cwg_kernel<true><<<grid, 256, 0, st>>>(d_out, num_samples, out_stride, (float)cw_scale, step_fx, ph0, noise_mode,
                                               (float)noise_scale, seed);
    else
        cwg_kernel<false><<<grid, 256, 0, st>>>(d_out, num_samples, out_stride, (float)cw_scale, step_fx, ph0, noise_mode,
                                                (float)noise_scale, seed);
    CUDA_TRY(cudaGetLastError());
    h->launches++;
    return DDCB200_OK;
}

int ddcb200_run_host_f32(ddcb200_t* h, const float* h_in, int64_t n_samples, int64_t n_streams, int64_t in_stride,
                         double step, int64_t sample_offset, ddcb200_c64* h_out, int64_t out_stride) {
    if (!h) return fail(DDCB200_EINVAL, "null handle");
    DeviceGuard g(h->device);
    const int T = (int)h->taps.size();
    if (n_samples > 0 && n_samples < T) {
        if (n_streams != 1) return fail(DDCB200_ETOOSHORT, "n_samples < n_taps is only supported for a single stream");
        if (!h_in || !h_out) return fail(DDCB200_EINVAL, "null host pointer");
        const int64_t m = ddcb200_out_len(n_samples, T, h->decim);
        int rc = ensure_chunks(h, (size_t)T * 4, (size_t)m);
        if (rc) return rc;
        CUDA_TRY(cudaMemcpyAsync(h->d_chunk_in[0], h_in, (size_t)n_samples * 4, cudaMemcpyHostToDevice, h->stream));
        rc = ddcb200_run_short_f32(h, reinterpret_cast<const float*>(h->d_chunk_in[0]), n_samples, step, sample_offset,
                                   h->d_chunk_out[0], h->stream);
        if (rc) return rc;
        CUDA_TRY(cudaMemcpyAsync(h_out, h->d_chunk_out[0], (size_t)m * sizeof(ddcb200_c64), cudaMemcpyDeviceToHost, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        return DDCB200_OK;
    }
    return run_host(h, h_in, false, n_samples, n_streams, in_stride, step, sample_offset, h_out, out_stride);
}

int ddcb200_run_host_f32_c128(ddcb200_t* h, const float* h_in, int64_t n_samples, double step, int64_t sample_offset,
                              double* h_out) {
    if (!h) return fail(DDCB200_EINVAL, "null handle");
    if (!h_in || !h_out) return fail(DDCB200_EINVAL, "null host pointer");
    DeviceGuard g(h->device);
    const int T = (int)h->taps.size();
    if (n_samples > 0 && n_samples < T) {   // the scipy operand swap: few outputs, go through the complex64 entry point
        const int64_t m = ddcb200_out_len(n_samples, T, h->decim);
        std::vector<ddcb200_c64> tmp((size_t)m);
        int rc = ddcb200_run_host_f32(h, h_in, n_samples, 1, n_samples, step, sample_offset, tmp.data(), m);
        if (rc) return rc;
        widen_c64_to_c128(tmp.data(), h_out, (size_t)m, 1);
        return DDCB200_OK;
    }
    return run_host(h, h_in, false, n_samples, 1, n_samples, step, sample_offset, nullptr, 0, h_out);
}

int ddcb200_run_host_packed10(ddcb200_t* h, const uint8_t* h_in, int64_t n_samples, int64_t n_streams,
                              int64_t in_stride_bytes, double step, int64_t sample_offset, ddcb200_c64* h_out,
                              int64_t out_stride) {
    if (!h) return fail(DDCB200_EINVAL, "null handle");
    DeviceGuard g(h->device);
    return run_host(h, h_in, true, n_samples, n_streams, in_stride_bytes, step, sample_offset, h_out, out_stride);
}

void* ddcb200_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) {
        fail(DDCB200_ENOMEM, "host_alloc(%zu) failed", bytes);
        return nullptr;
    }
    return p;
}
void ddcb200_host_free(void* p) {
    if (p) cudaFreeHost(p);
}

int ddcb200_sync(ddcb200_t* h) {
    if (!h) return fail(DDCB200_EINVAL, "null handle");
    DeviceGuard g(h->device);
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return DDCB200_OK;
}
void* ddcb200_stream(ddcb200_t* h) { return h ? (void*)h->stream : nullptr; }
int64_t ddcb200_launch_count(ddcb200_t* h) { return h ? h->launches : 0; }
const char* ddcb200_last_variant(ddcb200_t* h) { return h ? h->last_variant.c_str() : "none"; }

int ddcb200_set_option(ddcb200_t* h, const char* key, int64_t value) {
    if (!h || !key) return fail(DDCB200_EINVAL, "set_option: bad arguments");
    if (!strcmp(key, "variant")) {
        h->force_variant = (int)value;
        return DDCB200_OK;
    }
    if (!strcmp(key, "dbg_counters")) {   // 1: allocate + zero, 0: free, 2: print wait/total cycle ratio to stderr
        DeviceGuard g(h->device);
        if (value == 1) {
            if (!h->d_dbg) CUDA_TRY(cudaMalloc(&h->d_dbg, 64));
            CUDA_TRY(cudaMemset(h->d_dbg, 0, 64));
        } else if (value == 2 && h->d_dbg) {
            unsigned long long v[2] = {0, 0};
            CUDA_TRY(cudaDeviceSynchronize());
            CUDA_TRY(cudaMemcpy(v, h->d_dbg, 16, cudaMemcpyDeviceToHost));
            fprintf(stderr, "dbg_counters: compute warps waited %llu of %llu cycles = %.2f %%\n", v[0], v[1], v[1] ? 100.0 * v[0] / v[1] : 0.0);
        } else if (value == 0 && h->d_dbg) {
            cudaFree(h->d_dbg);
            h->d_dbg = nullptr;
        }
        return DDCB200_OK;
    }
    if (!strcmp(key, "l2_ahead")) {
        h->l2_ahead = (int)value;
        return DDCB200_OK;
    }
    if (!strcmp(key, "stagger_cycles")) {
        h->stagger_cycles = (int)value;
        return DDCB200_OK;
    }
    if (!strcmp(key, "debug_mode")) {
        h->debug_mode = (int)value;
        return DDCB200_OK;
    }
    if (!strcmp(key, "copy_threads")) {   // host threads of the pageable-input staging path (0 = let the driver stage)
        if (value < 0 || value > 16) return fail(DDCB200_EINVAL, "copy_threads must be 0 .. 16");
        h->copy_threads = (int)value;
        return DDCB200_OK;
    }
    if (!strcmp(key, "chunk_samples")) {
        if (value < 1024) return fail(DDCB200_EINVAL, "chunk_samples too small");
        h->chunk_samples = value;
        return DDCB200_OK;
    }
    return fail(DDCB200_EINVAL, "unknown option '%s'", key);
}

// ================================================================================================================
// Streaming sessions: a continuous digitiser stream arrives in arbitrary pieces; the session keeps the
// last T-D .. T-1 samples of every stream in device memory so that no output is lost or duplicated at push boundaries,
// and the absolute sample index so that the NCO phase is continuous.
// ================================================================================================================
struct ddcb200_session {
    ddcb200* h = nullptr;
    int64_t n_streams = 0, max_chunk = 0;
    double step = 0.0;
    bool packed = false;   // work buffers hold packed 10-bit bytes (5 per 4 samples) instead of float32
    int64_t carry = 0;     // samples of every stream waiting at the head of work[cur] (a multiple of 4 when packed)
    int64_t abs0 = 0;      // absolute index of the first carried sample (always a multiple of D)
    int64_t pitch = 0;     // BYTES per stream row of a work buffer (multiple of 16: rows stay 16-byte aligned)
    int64_t out_cap = 0;   // outputs per stream a push of max_chunk samples can produce
    unsigned char* work[2] = {};
    ddcb200_c64* dout[2] = {};
    cudaEvent_t ev_in[2] = {}, ev_k[2] = {}, ev_out[2] = {};
    cudaEvent_t ev_user = nullptr;   // end of the last asynchronous device push (on the caller's stream)
    bool user_pending = false;
    int cur = 0;
    size_t bytes(int64_t samples) const { return packed ? (size_t)(samples / 4 * 5) : (size_t)samples * 4; }
};

namespace {
// One piece (n <= max_chunk samples per stream, already in work[cur] behind the carry): run the DDC over carry + n samples
// into d_out and move the unconsumed tail to the head of the other work buffer.  Everything on `st`.
int stream_step(ddcb200_session* s, int64_t n, ddcb200_c64* d_out, int64_t out_stride, int64_t* n_out, cudaStream_t st) {
    ddcb200* h = s->h;
    const int T = (int)h->taps.size(), D = h->decim;
    const int64_t have = s->carry + n;
    int64_t m = 0;
    if (have >= T) {
        m = (have - T) / D + 1;
        int rc = run_device(h, s->work[s->cur], s->packed, have, s->n_streams, s->packed ? s->pitch : s->pitch / 4, s->step, s->abs0,
                            d_out, out_stride, st);
        if (rc) return rc;
    }
    const int64_t used = m * D, rest = have - used;
    if (rest > 0 && used > 0)
        CUDA_TRY(copy_rows_async(s->work[s->cur ^ 1], (size_t)s->pitch, s->work[s->cur] + s->bytes(used), (size_t)s->pitch,
                                 s->bytes(rest), (size_t)s->n_streams, cudaMemcpyDeviceToDevice, st));
    if (used > 0) s->cur ^= 1;
    s->carry = rest;
    s->abs0 += used;
    *n_out = m;
    return DDCB200_OK;
}

int session_open(ddcb200_t* h, int64_t n_streams, int64_t max_chunk_samples, double phase_step_cycles, bool packed,
                 ddcb200_session_t** out) {
    if (!h || !out || n_streams <= 0 || n_streams > 65535 || max_chunk_samples <= 0)
        return fail(DDCB200_EINVAL, "session_open: bad arguments");
    *out = nullptr;
    DeviceGuard g(h->device);
    const int T = (int)h->taps.size(), D = h->decim;
    if (T < D) return fail(DDCB200_EINVAL, "session_open: streaming needs n_taps (%d) >= decimation (%d)", T, D);
    if (packed && (D % 4)) return fail(DDCB200_EINVAL, "session_open: packed streaming needs a decimation that is a multiple of 4");
    auto* s = new (std::nothrow) ddcb200_session();
    if (!s) return fail(DDCB200_ENOMEM, "session_open: out of host memory");
    s->h = h;
    s->n_streams = n_streams;
    s->packed = packed;
    s->max_chunk = std::max<int64_t>(max_chunk_samples, (int64_t)T);
    if (packed) s->max_chunk = (s->max_chunk + 3) / 4 * 4;
    s->step = phase_step_cycles;
    const int64_t row_samples = ((int64_t)(T + D) + s->max_chunk + 63) / 64 * 64;   // 64 samples = 256 B float32 = 80 B packed
    s->pitch = (int64_t)s->bytes(row_samples);
    s->out_cap = (s->max_chunk + T + D) / D + 1;
    for (int i = 0; i < 2; ++i) {
        if (cudaMalloc(&s->work[i], (size_t)s->pitch * (size_t)n_streams) != cudaSuccess ||
            cudaMalloc(&s->dout[i], (size_t)s->out_cap * sizeof(ddcb200_c64) * (size_t)n_streams) != cudaSuccess ||
            cudaEventCreateWithFlags(&s->ev_in[i], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&s->ev_k[i], cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&s->ev_out[i], cudaEventDisableTiming) != cudaSuccess) {
            cudaGetLastError();
            ddcb200_session_close(s);
            return fail(DDCB200_ENOMEM, "session_open: device allocation failed (%lld streams x %lld bytes)",
                        (long long)n_streams, (long long)s->pitch);
        }
    }
    if (cudaEventCreateWithFlags(&s->ev_user, cudaEventDisableTiming) != cudaSuccess) {
        cudaGetLastError();
        ddcb200_session_close(s);
        return fail(DDCB200_ECUDA, "session_open: event creation failed");
    }
    *out = s;
    return DDCB200_OK;
}

// device-pointer push; in_stride in elements of the input type (floats, or bytes when packed)
int session_push_dev(ddcb200_session_t* s, const void* d_in, bool packed, int64_t n_samples, int64_t in_stride, ddcb200_c64* d_out,
                     int64_t out_stride, int64_t* n_out, void* cuda_stream) {
    if (!s || !d_in || !n_out || n_samples <= 0) return fail(DDCB200_EINVAL, "session_push: bad arguments");
    if (packed != s->packed) return fail(DDCB200_EINVAL, "session_push: this session was opened for %s input", s->packed ? "packed" : "float32");
    if (packed && (n_samples % 4)) return fail(DDCB200_EINVAL, "session_push: packed pushes need n_samples %% 4 == 0");
    if (n_samples > s->max_chunk) return fail(DDCB200_EINVAL, "session_push: %lld samples > max_chunk_samples %lld",
                                              (long long)n_samples, (long long)s->max_chunk);
    const int64_t m = ddcb200_session_out_len(s, n_samples);
    if (m > 0 && (!d_out || out_stride < m)) return fail(DDCB200_EINVAL, "session_push: output too small for %lld outputs", (long long)m);
    DeviceGuard g(s->h->device);
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : s->h->stream;
    CUDA_TRY(copy_rows_async(s->work[s->cur] + s->bytes(s->carry), (size_t)s->pitch, d_in, (size_t)in_stride * (packed ? 1 : 4),
                             s->bytes(n_samples), (size_t)s->n_streams, cudaMemcpyDeviceToDevice, st));
    int rc = stream_step(s, n_samples, d_out, out_stride, n_out, st);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(s->ev_user, st));
    s->user_pending = true;
    return DDCB200_OK;
}

int session_push_host(ddcb200_session_t* s, const void* h_in, bool packed, int64_t n_samples, int64_t in_stride, ddcb200_c64* h_out,
                      int64_t out_stride, int64_t* n_out) {
    if (!s || !h_in || !n_out || n_samples <= 0) return fail(DDCB200_EINVAL, "session_push_host: bad arguments");
    if (packed != s->packed) return fail(DDCB200_EINVAL, "session_push_host: this session was opened for %s input", s->packed ? "packed" : "float32");
    if (packed && (n_samples % 4)) return fail(DDCB200_EINVAL, "session_push_host: packed pushes need n_samples %% 4 == 0");
    const int64_t m_total = ddcb200_session_out_len(s, n_samples);
    if (m_total > 0 && (!h_out || out_stride < m_total)) return fail(DDCB200_EINVAL, "session_push_host: output too small");
    ddcb200* h = s->h;
    DeviceGuard g(h->device);
    DrainGuard drain(h);
    // Pieces of max_chunk samples, two work buffers deep: the H2D copy of piece i + 1 (copy_in stream) runs under the
    // kernel of piece i (compute stream), the D2H copy of its outputs on copy_out.
    if (s->user_pending) {   // an asynchronous device push may still be using the work buffers on the caller's stream
        CUDA_TRY(cudaStreamWaitEvent(h->copy_in, s->ev_user, 0));
        CUDA_TRY(cudaStreamWaitEvent(h->stream, s->ev_user, 0));
        s->user_pending = false;
    }
    const size_t src_pitch = (size_t)in_stride * (packed ? 1 : 4);
    int64_t done = 0, m_done = 0;
    bool rec_k[2] = {false, false}, rec_out[2] = {false, false};
    while (done < n_samples) {
        const int64_t n = std::min<int64_t>(s->max_chunk, n_samples - done);
        const int b = s->cur;
        if (rec_k[b]) CUDA_TRY(cudaStreamWaitEvent(h->copy_in, s->ev_k[b], 0));     // last kernel / tail copy reading work[b]
        CUDA_TRY(copy_rows_async(s->work[b] + s->bytes(s->carry), (size_t)s->pitch,
                                 reinterpret_cast<const unsigned char*>(h_in) + s->bytes(done), src_pitch, s->bytes(n),
                                 (size_t)s->n_streams, cudaMemcpyHostToDevice, h->copy_in));
        CUDA_TRY(cudaEventRecord(s->ev_in[b], h->copy_in));
        CUDA_TRY(cudaStreamWaitEvent(h->stream, s->ev_in[b], 0));
        if (rec_out[b]) CUDA_TRY(cudaStreamWaitEvent(h->stream, s->ev_out[b], 0));  // dout[b] has been copied out
        int64_t m = 0;
        int rc = stream_step(s, n, s->dout[b], s->out_cap, &m, h->stream);
        if (rc) return rc;
        CUDA_TRY(cudaEventRecord(s->ev_k[b], h->stream));
        rec_k[b] = true;
        if (m > 0) {
            CUDA_TRY(cudaStreamWaitEvent(h->copy_out, s->ev_k[b], 0));
            CUDA_TRY(copy_rows_async(h_out + m_done, (size_t)out_stride * sizeof(ddcb200_c64), s->dout[b],
                                     (size_t)s->out_cap * sizeof(ddcb200_c64), (size_t)m * sizeof(ddcb200_c64),
                                     (size_t)s->n_streams, cudaMemcpyDeviceToHost, h->copy_out));
            CUDA_TRY(cudaEventRecord(s->ev_out[b], h->copy_out));
            rec_out[b] = true;
        }
        done += n;
        m_done += m;
    }
    CUDA_TRY(cudaStreamSynchronize(h->copy_out));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->copy_in));
    *n_out = m_done;
    return DDCB200_OK;
}
}  // namespace

int ddcb200_session_open(ddcb200_t* h, int64_t n_streams, int64_t max_chunk_samples, double phase_step_cycles,
                         ddcb200_session_t** out) {
    return session_open(h, n_streams, max_chunk_samples, phase_step_cycles, false, out);
}
int ddcb200_session_open_packed10(ddcb200_t* h, int64_t n_streams, int64_t max_chunk_samples, double phase_step_cycles,
                                  ddcb200_session_t** out) {
    return session_open(h, n_streams, max_chunk_samples, phase_step_cycles, true, out);
}

void ddcb200_session_close(ddcb200_session_t* s) {
    if (!s) return;
    DeviceGuard g(s->h->device);
    cudaDeviceSynchronize();
    for (int i = 0; i < 2; ++i) {
        if (s->work[i]) cudaFree(s->work[i]);
        if (s->dout[i]) cudaFree(s->dout[i]);
        if (s->ev_in[i]) cudaEventDestroy(s->ev_in[i]);
        if (s->ev_k[i]) cudaEventDestroy(s->ev_k[i]);
        if (s->ev_out[i]) cudaEventDestroy(s->ev_out[i]);
    }
    if (s->ev_user) cudaEventDestroy(s->ev_user);
    delete s;
}

int ddcb200_session_reset(ddcb200_session_t* s, int64_t first_sample_index) {
    if (!s || first_sample_index < 0) return fail(DDCB200_EINVAL, "session_reset: bad arguments");
    if (first_sample_index % s->h->decim) return fail(DDCB200_EINVAL, "session_reset: index must be a multiple of the decimation");
    s->carry = 0;
    s->abs0 = first_sample_index;
    return DDCB200_OK;
}

int64_t ddcb200_session_pending(ddcb200_session_t* s) { return s ? s->carry : 0; }
int64_t ddcb200_session_position(ddcb200_session_t* s) { return s ? s->abs0 + s->carry : 0; }

int64_t ddcb200_session_out_len(ddcb200_session_t* s, int64_t n_samples) {
    if (!s || n_samples < 0) return 0;
    const int64_t have = s->carry + n_samples, T = (int64_t)s->h->taps.size();
    return have >= T ? (have - T) / s->h->decim + 1 : 0;
}

int ddcb200_session_push_f32(ddcb200_session_t* s, const float* d_in, int64_t n_samples, int64_t in_stride, ddcb200_c64* d_out,
                             int64_t out_stride, int64_t* n_out, void* cuda_stream) {
    return session_push_dev(s, d_in, false, n_samples, in_stride, d_out, out_stride, n_out, cuda_stream);
}
int ddcb200_session_push_packed10(ddcb200_session_t* s, const uint8_t* d_in, int64_t n_samples, int64_t in_stride_bytes,
                                  ddcb200_c64* d_out, int64_t out_stride, int64_t* n_out, void* cuda_stream) {
    return session_push_dev(s, d_in, true, n_samples, in_stride_bytes, d_out, out_stride, n_out, cuda_stream);
}
int ddcb200_session_push_host_f32(ddcb200_session_t* s, const float* h_in, int64_t n_samples, int64_t in_stride, ddcb200_c64* h_out,
                                  int64_t out_stride, int64_t* n_out) {
    return session_push_host(s, h_in, false, n_samples, in_stride, h_out, out_stride, n_out);
}
int ddcb200_session_push_host_packed10(ddcb200_session_t* s, const uint8_t* h_in, int64_t n_samples, int64_t in_stride_bytes,
                                       ddcb200_c64* h_out, int64_t out_stride, int64_t* n_out) {
    return session_push_host(s, h_in, true, n_samples, in_stride_bytes, h_out, out_stride, n_out);
}

}  // extern "C"
