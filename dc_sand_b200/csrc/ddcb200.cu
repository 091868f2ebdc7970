// C-ABI implementation of include/ddcb200.h: handle management, per-call host-side preparation of the folded
// complex taps (float64 -> float32), kernel dispatch, and the chunked double-buffered host path.  The fused kernels are
// launched through ddch::launch_* (one translation unit per kernel family, k_*.cu).
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

#include "ddc_host.h"
#include "ddc_kernels.cuh"

using namespace ddck;

namespace ddch {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

}  // namespace ddch

using ddch::fail;

namespace {

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

}  // namespace

namespace ddch {
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
namespace {
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
}  // namespace

// cached front end of make_wtaps (invalidated by set_taps / set_decimation through wt_jt = 0)
const float2* cached_wtaps(ddcb200* h, double step, int jt, int D) {
    if (h->wt_jt != jt || h->wt_d != D || h->wt_step != step || h->wt_cache.size() != (size_t)(3 * (jt / 2) * D)) {
        h->wt_cache.resize((size_t)(3 * (jt / 2) * D));
        make_wtaps(h, step, jt, D, h->wt_cache.data());
        h->wt_step = step;
        h->wt_jt = jt;
        h->wt_d = D;
    }
    return h->wt_cache.data();
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
int launch_ws(ddcb200* h, RunParams& p, const float* d_in, long long n_rows, cudaStream_t st, double step, int D, int jt) {
    switch (D) {
        case 4: return launch_ws4(h, p, d_in, n_rows, st, step, jt);
        case 8: return launch_ws8(h, p, d_in, n_rows, st, step, jt);
        case 16: return launch_ws16(h, p, d_in, n_rows, st, step, jt);
        case 32: return launch_ws32(h, p, d_in, n_rows, st, step, jt);
        case 64: return launch_ws64(h, p, d_in, n_rows, st, step, jt);
    }
    return fail(DDCB200_EINVAL, "tensor-staged kernel: unsupported decimation %d", D);
}
}  // namespace ddch

using ddch::make_ctaps;

namespace {

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
    h->ring_cap = 0;   // only valid again once every slot has been allocated (a failure below leaves null slots behind)
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

// ---------------------------------------------------------------------------------------------------------------------
// Kernel selection: a PURE function of the call's shape and the handle's options (exported as ddcb200_plan so that the
// selection table is testable without a GPU).  Rules in order; the first that applies wins:
//   1  packed, 16-byte aligned rows, D a power of two 4 .. 64, filter fits       -> tensor-core engine (ddc_kernel_tc.cuh)
//   2  packed, aligned, D = 16, 129 .. 256 taps, packed_engine = 0 / variant 10  -> warp-specialised CUDA-core kernel (ddc_kernel_w10.cuh)
//   3  packed otherwise (variant != 1)                                           -> unpack stage + the float32 plan of the cell
//   4  float32, aligned, tensor-staged / sliced fast FIR (ddc_kernel_ws.cuh):
//        D = 4 / 8 up to 1024 taps where its tap padding costs <= 12 % over the tile kernel's, D = 16 up to 128 taps,
//        D = 32 / 64 where the direct form is FP32-bound (4 T / D flop against 4 + 8 / D bytes at the ridge of 11.4 flop/B)
//   5  float32, aligned, ring kernels on 1-D bulk copies: D = 16 up to 64 tap blocks -> fast FIR (ddc_kernel_w.cuh);
//        D = 32 / 64 up to 16 tap blocks -> phase-major direct form (ddc_kernel_p.cuh)
//   6  float32, aligned, D a power of two, padded filter <= 2048 taps            -> rotating-window tile kernel
//   7  everything else (odd decimations, unaligned rows, > 2048 taps, variant 1) -> generic kernel
// `skip` excludes families a caller has found inapplicable at run time (tile counts beyond an index range, inputs shorter
// than one thread-row).  Measurements behind the thresholds: DESIGN.md section 4.
// ---------------------------------------------------------------------------------------------------------------------
enum KernelFamily { KF_TENSOR10 = 0, KF_W10S, KF_UNPACK_F32, KF_WS, KF_W, KF_PD, KF_TILE, KF_GENERIC, KF_COUNT };
const char* const kFamilyName[KF_COUNT] = {"tensor10", "w10s", "unpack+f32", "ws", "w", "pd", "tile", "generic"};

struct KernelPlan {
    KernelFamily family;
    int jt;   // tap blocks the launcher is instantiated for (WS / W / PD), padded blocks of the tile kernel
    int ks;   // tile kernel: warps that share a tap split
};

KernelPlan plan_kernel(int T, int D, bool packed, bool aligned, int fv, int packed_engine, bool tensor_ok, unsigned skip = 0) {
    const int Jp = (T + D - 1) / D;   // tap blocks of D taps
    const bool pow2_d = D == 4 || D == 8 || D == 16 || D == 32 || D == 64;
    auto allowed = [&](KernelFamily f) { return !(skip & (1u << f)); };
    if (packed) {
        if (aligned && pow2_d && tensor_ok && (fv == 13 || (fv == 0 && packed_engine == 1)) && allowed(KF_TENSOR10)) return {KF_TENSOR10, 0, 0};
        if (aligned && D == 16 && Jp > 8 && Jp <= 16 && (fv == 0 || fv == 10)) return {KF_W10S, 16, 0};
        if (fv != 1) return {KF_UNPACK_F32, 0, 0};
        return {KF_GENERIC, 0, 0};
    }
    if (aligned && pow2_d && fv != 1 && allowed(KF_WS)) {
        const int r_tile = 64 / D;   // outputs per thread of the tile kernel
        const bool fp32_bound = 4.0 * T / D > 11.4 * (4.0 + 8.0 / D);
        const int jt_ws = Jp <= 8 ? 8 : (Jp <= 64 ? (Jp + 15) / 16 * 16 : (Jp + 31) / 32 * 32);   // the instantiations of k_ws.inc
        const int j_tile = (Jp + r_tile - 1) / r_tile * r_tile;
        const bool pad_ok = jt_ws * 100 <= j_tile * 112;
        const bool auto_ws = ((D == 4 || D == 8) && T <= 1024 && pad_ok) || (D == 16 && T <= 128) || ((D == 32 || D == 64) && fp32_bound);
        if (Jp <= (D == 4 ? 256 : (D == 8 ? 128 : 32)) && T >= D && (fv == 11 || (fv == 0 && auto_ws))) return {KF_WS, jt_ws, 0};
    }
    if (aligned && ((D == 16 && Jp <= 64) || ((D == 32 || D == 64) && Jp <= 16)) && (fv == 0 || fv == 7 || fv == 8)) {
        const int jt = Jp <= 4 ? 4 : (Jp <= 8 ? 8 : (Jp <= 16 ? 16 : (Jp <= 32 ? 32 : 64)));
        return {D == 16 ? KF_W : KF_PD, jt, 0};
    }
    if (aligned && pow2_d && fv != 1) {
        const int R = 64 / D;
        // tap split 2 (16 compute warps) whenever it costs no extra zero taps; option "variant" 2 / 3 force KS 1 / 2
        int ks = (Jp % (2 * R) == 0 && R <= 4) ? 2 : 1;
        if (fv == 2) ks = 1;
        if (fv == 3 && R <= 4) ks = 2;
        const int J = ((Jp + ks * R - 1) / (ks * R)) * (ks * R);   // each thread's tap-block loop is unrolled R times
        if (J * D <= 2048) return {KF_TILE, J, ks};
    }
    return {KF_GENERIC, 0, 0};
}

// Core dispatcher for device-resident data: plan_kernel() picks the family, this function sizes the launch.  Option "variant"
// (ddcb200_set_option): 0 auto; 1 generic kernel; 2 / 3 tile kernel without / with the tap split; 7 fast FIR (ddc_kernel_w.cuh,
// D = 16); 8 phase-major direct form (D = 32 / 64); 10 warp-specialised CUDA-core packed kernel; 11 tensor-staged / sliced fast
// FIR; 13 tensor-core engine for packed input; a choice that is not built for the call's (taps, decimation) falls through.
int run_device(ddcb200* h, const void* d_in, bool packed, int64_t n_samples, int64_t n_streams, int64_t in_stride,
               double step, int64_t sample_offset, ddcb200_c64* d_out, int64_t out_stride, cudaStream_t st,
               int64_t m_limit = -1) {
    const int T = (int)h->taps.size();
    const int D = h->decim;
    const int fv = h->force_variant;
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
    p.dbg = h->d_dbg;

    // ring kernels (ddc_kernel_p / _w / _w10.cuh): chunks of 32 thread-rows of 128 samples
    auto ring_geometry = [&](int jt) {
        const long long chunk_out = 32LL * (128 / D);
        p.tiles_per_stream = (M + chunk_out - 1) / chunk_out;
        p.total_tiles = p.tiles_per_stream * n_streams;
        p.n_taps = jt * D;
        p.n_tap_blocks = jt;
        p.m_begin = 0;
    };

    const bool aligned = packed ? (reinterpret_cast<uintptr_t>(d_in) % 16 == 0) && (in_stride % 16 == 0) : aligned_f32(d_in, in_stride, packed);
    const bool tensor_ok = packed && ddch::tc10_supported(h, T, D) && (double)((M + 127) / 128) * (double)n_streams < 2.0e9;
    KernelPlan plan = plan_kernel(T, D, packed, aligned, fv, h->packed_engine, tensor_ok);
    long long m_done = 0;        // outputs [0, m_done) of every stream are written by the fused kernel of the plan
    bool fused_done = false;

    if (plan.family == KF_TENSOR10) return ddch::launch_tc10(h, p, st, step, D);

    if (plan.family == KF_W10S) {
        ring_geometry(plan.jt);
        return ddch::launch_w10s(h, p, st, step);
    }

    if (plan.family == KF_UNPACK_F32) {
        // packed input without a fused-unpack kernel for this (T, D): unpack into a float32 workspace, then the float32 plan
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

    if (plan.family == KF_WS) {
        // tensor-staged / sliced fast FIR: its tensor map covers whole thread-rows of 8 blocks; the few outputs that need the last
        // partial thread-row come from the generic kernel below.  Inputs shorter than one thread-row re-plan without it.
        const long long n_blocks = (n_samples / (8LL * D)) * 8;
        long long m_f = n_blocks * D >= T ? (n_blocks * D - T) / D + 1 : 0;   // outputs whose window lies inside them
        if (m_f > M) m_f = M;
        // chunk_of() divides by multiplication: exact while total chunks x chunks per stream < 2^64
        if ((double)((m_f + 255) / 256) * (double)((m_f + 255) / 256) * (double)n_streams >= 1.8e19) m_f = 0;
        if (m_f > 0) {
            p.n_out = m_f;
            p.tiles_per_stream = (m_f + 255) / 256;
            p.total_tiles = p.tiles_per_stream * n_streams;
            p.n_taps = plan.jt * D;
            p.n_tap_blocks = plan.jt;
            p.m_begin = 0;
            int rc2 = ddch::launch_ws(h, p, reinterpret_cast<const float*>(d_in), n_blocks / 8, st, step, D, plan.jt);
            if (rc2) return rc2;
            p.n_out = M;
            m_done = m_f;
            fused_done = true;
        } else {
            plan = plan_kernel(T, D, packed, aligned, fv, h->packed_engine, tensor_ok, 1u << KF_WS);
        }
    }

    if (plan.family == KF_W || plan.family == KF_PD) {
        ring_geometry(plan.jt);
        if (plan.family == KF_W) return ddch::launch_w(h, p, st, step, D, plan.jt);
        std::vector<float2> ctp((size_t)plan.jt * D);
        make_ctaps(h, step, plan.jt * D, ctp.data());
        return ddch::launch_pd(h, p, ctp.data(), st, D, plan.jt);
    }

    if (plan.family == KF_TILE) {
        const int R = 64 / D;
        const int n_taps_pad = plan.jt * D;
        std::vector<float2> ct((size_t)n_taps_pad);
        make_ctaps(h, step, n_taps_pad, ct.data());
        const long long tile_out = 256LL * R;
        p.tiles_per_stream = (M + tile_out - 1) / tile_out;   // the last one may be ragged
        p.total_tiles = p.tiles_per_stream * n_streams;
        p.n_taps = n_taps_pad;
        p.n_tap_blocks = plan.jt;
        p.halo_rows = (plan.jt + R - 2) / R;   // a thread-row reads blocks 0 .. J+R-2 of its own row space
        p.m_begin = 0;
        int rc2 = ddch::launch_tile(h, p, ct.data(), st, D, plan.ks);
        if (rc2) return rc2;
        m_done = M;
        fused_done = true;
    }

    if (m_done < M) {
        // stream tails / everything the fused kernels do not cover (odd decimations, unaligned rows, more than 2048 taps)
        int rc = ensure_ring(h, T);
        if (rc) return rc;
        const int slot = h->ring_pos;
        h->ring_pos = (h->ring_pos + 1) % ddcb200::kRing;
        CUDA_TRY(cudaEventSynchronize(h->ring_ev[slot]));  // previous user of this slot has consumed it
        make_ctaps(h, step, T, h->h_ctaps[slot]);
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
        if (!fused_done) h->last_variant = packed ? "generic<packed10>" : "generic<f32>";
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

// the handle's parked host threads (host_pool.h): copy_threads - 1 workers, never more than the machine has cores to spare
ddch::HostPool& host_pool(ddcb200* h) {
    const int hw = (int)std::thread::hardware_concurrency();
    const int want = std::max(0, std::min({h->copy_threads, 16, hw > 1 ? hw - 1 : 1}) - 1);
    if (!h->pool || h->pool->workers() != want) h->pool.reset(new ddch::HostPool(want));
    return *h->pool;
}

int staged_h2d(ddcb200* h, void* d_dst, const void* h_src, size_t bytes, cudaStream_t st) {
    for (int i = 0; i < ddcb200::kStage; ++i) {
        if (!h->h_stage[i]) CUDA_TRY(cudaMallocHost(&h->h_stage[i], ddcb200::kStageBytes));
        if (!h->ev_stage[i]) CUDA_TRY(cudaEventCreateWithFlags(&h->ev_stage[i], cudaEventDisableTiming));
    }
    ddch::HostPool& pool = host_pool(h);
    size_t done = 0;
    while (done < bytes) {
        const size_t n = std::min(bytes - done, ddcb200::kStageBytes);
        const int b = h->stage_pos;
        h->stage_pos = (h->stage_pos + 1) % ddcb200::kStage;
        CUDA_TRY(cudaEventSynchronize(h->ev_stage[b]));   // the copy engine has drained this staging buffer
        char* dst = static_cast<char*>(h->h_stage[b]);
        const char* src = static_cast<const char*>(h_src) + done;
        if (n < (1u << 20)) {
            std::memcpy(dst, src, n);
        } else {
            const size_t per = 512u << 10;   // pieces of 512 KB: enough of them that a late worker does not set the pace
            const int parts = (int)((n + per - 1) / per);
            pool.run(parts, [=](int i) { std::memcpy(dst + per * i, src + per * i, std::min(per, n - per * i)); });
        }
        CUDA_TRY(cudaMemcpyAsync(static_cast<char*>(d_dst) + done, dst, n, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaEventRecord(h->ev_stage[b], st));
        done += n;
    }
    return DDCB200_OK;
}

// complex64 -> complex128 on the host threads, straight into the caller's (usually fresh, untouched) array: the page
// faults of the first touch are spread over the threads as well
void widen_c64_to_c128(ddcb200* h, const ddcb200_c64* src, double* dst, size_t n) {
    auto work = [=](size_t a, size_t b) {
        for (size_t i = a; i < b; ++i) {
            dst[2 * i] = (double)src[i].re;
            dst[2 * i + 1] = (double)src[i].im;
        }
    };
    if (!h || n < (1u << 16)) {
        work(0, n);
        return;
    }
    const size_t per = 1u << 15;
    host_pool(h).run((int)((n + per - 1) / per), [=](int i) { work(per * i, std::min(n, per * (i + 1))); });
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
    // Chunking.  A batch whose whole streams fit the chunk budget is cut BY STREAM: a chunk is a group of whole rows, which for
    // contiguous rows is one flat copy each way (no 2-D copy, no halo copied twice).  Longer streams are cut in TIME: chunk
    // length in outputs, inputs of a chunk start at a multiple of 64 samples so that packed chunks start on a byte boundary and
    // float chunks stay 16-byte aligned.  Option "host_chunk_mode" = 1 forces time chunks.
    const bool by_stream = n_streams > 1 && !h_out128 && h->host_chunk_mode != 1 && n_samples <= h->chunk_samples;
    const int64_t spc = by_stream ? std::max<int64_t>(1, std::min<int64_t>(h->chunk_samples / n_samples, (n_streams + 2) / 3))
                                  : n_streams;                       // streams per chunk (at least three chunks: the pipeline)
    const int64_t n_sgroups = (n_streams + spc - 1) / spc;
    int64_t per_stream = std::max<int64_t>(h->chunk_samples / n_streams, (int64_t)4 * T);
    int64_t m_chunk = by_stream ? M : std::max<int64_t>(per_stream / D, 1);
    m_chunk = ((m_chunk + 63) / 64) * 64;   // chunk starts stay 16-byte aligned for float32 AND packed (64 samples = 80 B)
    const int64_t n_tchunks = (M + m_chunk - 1) / m_chunk;
    const int64_t n_chunks = n_tchunks * n_sgroups;
    const int64_t in_chunk_samples = std::max<int64_t>(((m_chunk - 1) * D + T + 63) / 64 * 64, by_stream ? (n_samples + 63) / 64 * 64 : 0);
    const size_t in_elem_bytes_num = packed ? 5 : 16, in_elem_den = 4;  // bytes per 4 samples
    const size_t in_row_bytes = (size_t)in_chunk_samples / in_elem_den * in_elem_bytes_num;
    int rc = ensure_chunks(h, in_row_bytes * (size_t)spc + 64, (size_t)m_chunk * (size_t)spc);
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
        widen_c64_to_c128(h, h->h_ostage[pend_b], h_out128 + 2 * pend_m0, (size_t)pend_mc);
        pend_m0 = -1;
        return DDCB200_OK;
    };

    for (int64_t c = 0; c < n_chunks; ++c) {
        const int b = (int)(c % ddcb200::kBufs);
        const int64_t s0 = (c / n_tchunks) * spc;                       // first stream of this chunk
        const int64_t ns = std::min<int64_t>(spc, n_streams - s0);      // its streams
        const int64_t m0 = (c % n_tchunks) * m_chunk;
        const int64_t mc = std::min<int64_t>(m_chunk, M - m0);
        const int64_t n0 = m0 * D;
        const int64_t nc = (mc - 1) * D + T;                  // samples this chunk needs
        const int64_t nc4 = std::min<int64_t>((nc + 3) / 4 * 4, n_samples - n0);  // copy whole groups when available
        const size_t src_pitch = packed ? (size_t)in_stride : (size_t)in_stride * 4;
        const int64_t ncopy = by_stream ? n_samples : nc4;       // by-stream chunks copy whole rows (flat when rows are contiguous)
        const size_t row_bytes = packed ? (size_t)(ncopy / 4 * 5) : (size_t)ncopy * 4;
        const bool flat_in = by_stream && src_pitch == row_bytes && row_bytes % 16 == 0;
        const size_t src_off = (packed ? (size_t)(n0 / 4 * 5) : (size_t)n0 * 4) + (size_t)s0 * src_pitch;
        // buffer b is free once the D2H of chunk c - kBufs finished (ev_out) -- wait on the copy-in stream
        if (c >= ddcb200::kBufs) CUDA_TRY(cudaStreamWaitEvent(h->copy_in, h->ev_k[b], 0));
        if (n_streams == 1 && pageable_in && row_bytes >= (16u << 20)) {   // below that the thread start-up costs more than it saves
            rc = staged_h2d(h, h->d_chunk_in[b], reinterpret_cast<const char*>(h_in) + src_off, row_bytes, h->copy_in);
            if (rc) return rc;
        } else {
            // whole contiguous rows (by-stream chunks): the device rows take the host pitch, so the group is ONE flat copy
            const size_t dpitch = flat_in ? row_bytes : in_row_bytes;
            CUDA_TRY(copy_rows_async(h->d_chunk_in[b], dpitch, reinterpret_cast<const char*>(h_in) + src_off, src_pitch,
                                     row_bytes, (size_t)ns, cudaMemcpyHostToDevice, h->copy_in));
        }
        CUDA_TRY(cudaEventRecord(h->ev_in[b], h->copy_in));
        CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_in[b], 0));
        if (c >= ddcb200::kBufs) CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_out[b], 0));
        const int64_t n_dev = packed ? (nc4 / 4 * 4) : nc;
        const size_t d_in_pitch = flat_in ? row_bytes : in_row_bytes;
        const int64_t d_out_pitch = (by_stream && out_stride == mc) ? mc : m_chunk;   // contiguous host rows: one flat D2H
        rc = run_device(h, h->d_chunk_in[b], packed, packed ? nc4 : n_dev, ns,
                        packed ? (int64_t)d_in_pitch : (int64_t)(d_in_pitch / 4), step, sample_offset + n0,
                        h->d_chunk_out[b], d_out_pitch, h->stream, mc);
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
            CUDA_TRY(copy_rows_async(h_out + s0 * out_stride + m0, (size_t)out_stride * sizeof(ddcb200_c64), h->d_chunk_out[b],
                                     (size_t)d_out_pitch * sizeof(ddcb200_c64), (size_t)mc * sizeof(ddcb200_c64), (size_t)ns,
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
const char* ddcb200_last_error(void) { return ddch::g_err; }

int64_t ddcb200_out_len(int64_t n_samples, int n_taps, int decimation) {
    if (n_samples <= 0 || n_taps <= 0 || decimation <= 0) return 0;
    const int64_t full = (n_samples >= n_taps ? n_samples - n_taps : n_taps - n_samples) + 1;
    return (full + decimation - 1) / decimation;
}

int ddcb200_plan(int n_taps, int decimation, int packed, int aligned, int variant, int packed_engine, char* name, int name_cap) {
    if (n_taps <= 0 || decimation <= 0 || !name || name_cap < 24) return fail(DDCB200_EINVAL, "plan: bad arguments");
    const bool tensor_ok = packed && ddch::tc10_supported(nullptr, n_taps, decimation);
    KernelPlan pl = plan_kernel(n_taps, decimation, packed != 0, aligned != 0, variant, packed_engine, tensor_ok);
    if (pl.family == KF_UNPACK_F32) {
        // the unpack stage writes 16-byte aligned float32 rows, whatever the packed rows were
        pl = plan_kernel(n_taps, decimation, false, true, variant, packed_engine, false);
        snprintf(name, (size_t)name_cap, "%s:%s", kFamilyName[KF_UNPACK_F32], kFamilyName[pl.family]);
    } else {
        snprintf(name, (size_t)name_cap, "%s", kFamilyName[pl.family]);
    }
    return pl.jt;
}

int ddcb200_tensor_engine_geometry(int n_taps, int decimation, int32_t* out12) {
    if (!out12) return fail(DDCB200_EINVAL, "tensor_engine_geometry: null output");
    return ddch::tc10_describe(n_taps, decimation, out12);
}

int ddcb200_set_taps(ddcb200_t* h, const double* taps, int n_taps) {
    if (!h || !taps || n_taps <= 0) return fail(DDCB200_EINVAL, "set_taps: bad arguments");
    double s = 0.0;
    for (int i = 0; i < n_taps; ++i) s += taps[i];  // same left-to-right float64 sum as Python's sum() (ddc.py:98)
    if (!(s != 0.0) || !std::isfinite(s)) return fail(DDCB200_EINVAL, "set_taps: sum of taps is %g", s);
    h->taps.assign(taps, taps + n_taps);
    h->taps_sum = s;
    h->wt_jt = 0;   // invalidate the folded-tap caches
    h->taps_version++;
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
    for (int i = 0; i < ddcb200::kTcRing; ++i) {
        if (h->d_tc_b[i]) cudaFree(h->d_tc_b[i]);
        if (h->h_tc_b[i]) cudaFreeHost(h->h_tc_b[i]);
        if (h->tc_ev[i]) cudaEventDestroy(h->tc_ev[i]);
    }
    if (h->tc_up_ev) cudaEventDestroy(h->tc_up_ev);
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
        widen_c64_to_c128(nullptr, tmp.data(), h_out, (size_t)m);
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

static constexpr size_t kDbgBytes = (64 + 10 * 96) * sizeof(unsigned long long);   // counters + the tensor engine's event trace

int ddcb200_set_option(ddcb200_t* h, const char* key, int64_t value) {
    if (!h || !key) return fail(DDCB200_EINVAL, "set_option: bad arguments");
    if (!strcmp(key, "variant")) {
        h->force_variant = (int)value;
        return DDCB200_OK;
    }
    if (!strcmp(key, "dbg_counters")) {   // 1: allocate + zero, 0: free, 2: print wait/total cycle ratio to stderr, 3: print the event trace
        DeviceGuard g(h->device);
        if (value == 1) {
            if (!h->d_dbg) CUDA_TRY(cudaMalloc(&h->d_dbg, kDbgBytes));
            CUDA_TRY(cudaMemset(h->d_dbg, 0, kDbgBytes));
        } else if (value == 2 && h->d_dbg) {
            unsigned long long v[2] = {0, 0};
            CUDA_TRY(cudaDeviceSynchronize());
            CUDA_TRY(cudaMemcpy(v, h->d_dbg, 16, cudaMemcpyDeviceToHost));
            fprintf(stderr, "dbg_counters: compute warps waited %llu of %llu cycles = %.2f %%\n", v[0], v[1], v[1] ? 100.0 * v[0] / v[1] : 0.0);
            unsigned long long w[12] = {};
            CUDA_TRY(cudaMemcpy(w, h->d_dbg + 2, sizeof(w), cudaMemcpyDeviceToHost));
            static const char* role[4] = {"producer (raw_empty, -)", "mma (a_full, acc_empty)", "epilogue (acc_full, -)", "unpack (raw_full, a_empty)"};
            for (int r = 0; r < 4; ++r)
                if (w[3 * r + 2])
                    fprintf(stderr, "dbg_counters: tensor engine %-28s waits %.1f %% + %.1f %% of %llu cycles\n", role[r],
                            100.0 * w[3 * r] / w[3 * r + 2], 100.0 * w[3 * r + 1] / w[3 * r + 2], w[3 * r + 2]);
        } else if (value == 3 && h->d_dbg) {   // event trace of the tensor engine (builds with -DDDCB200_TC_TRACE fill it)
            std::vector<unsigned long long> tr(10 * 96);
            CUDA_TRY(cudaDeviceSynchronize());
            CUDA_TRY(cudaMemcpy(tr.data(), h->d_dbg + 64, tr.size() * 8, cudaMemcpyDeviceToHost));
            unsigned long long t0 = ~0ull;
            for (auto v : tr) if (v && v < t0) t0 = v;
            fprintf(stderr, "tc_trace: tile  tma_issue unp_start unp0_done unpL_done raw_next mma_start mma_issued acc_full acc_freed epi_done  (clocks since the first event)\n");
            static const int order[10] = {0, 1, 2, 3, 9, 4, 5, 6, 7, 8};
            for (int k = 0; k < 96; ++k) {
                fprintf(stderr, "tc_trace: %4d", k);
                for (int e : order) fprintf(stderr, " %9lld", tr[96 * e + k] ? (long long)(tr[96 * e + k] - t0) : -1LL);
                fprintf(stderr, "\n");
            }
        } else if (value == 0 && h->d_dbg) {
            cudaFree(h->d_dbg);
            h->d_dbg = nullptr;
        }
        return DDCB200_OK;
    }
    if (!strcmp(key, "packed_engine")) {   // 1: tcgen05 tensor cores for packed 10-bit input (default); 0: CUDA cores
        if (value != 0 && value != 1) return fail(DDCB200_EINVAL, "packed_engine must be 0 or 1");
        h->packed_engine = (int)value;
        return DDCB200_OK;
    }
    if (!strcmp(key, "tc_ns")) {   // tuning: sub-streams of the tensor engine (0 = automatic)
        if (value != 0 && value != 8 && value != 16) return fail(DDCB200_EINVAL, "tc_ns must be 0, 8 or 16");
        h->tc_ns = (int)value;
        return DDCB200_OK;
    }
    if (!strcmp(key, "host_chunk_mode")) {   // 0: by stream when whole streams fit a chunk (default); 1: always time chunks
        if (value != 0 && value != 1) return fail(DDCB200_EINVAL, "host_chunk_mode must be 0 or 1");
        h->host_chunk_mode = (int)value;
        return DDCB200_OK;
    }
    if (!strcmp(key, "tc_na") || !strcmp(key, "tc_nraw")) {   // tuning: pipeline depths of the tensor engine (0 = automatic)
        if (value < 0 || value > 8) return fail(DDCB200_EINVAL, "%s must be 0 .. 8", key);
        (key[4] == 'a' ? h->tc_na : h->tc_nraw) = (int)value;
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
    // the previous device push may have come on another stream: its kernel and tail copy still own the work buffers
    if (s->user_pending) CUDA_TRY(cudaStreamWaitEvent(st, s->ev_user, 0));
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
