// Host-side internals shared by the translation units of libddcb200.so: the handle, error reporting, tap folding and the
// per-family kernel launchers (one .cu per kernel family so that the families compile in parallel).
#pragma once
#include "../../include/ddcb200.h"

#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <memory>
#include <string>
#include <vector>

#include "ddc_common.cuh"
#include "host_pool.h"

namespace ddch {
int fail(int code, const char* fmt, ...);   // records the thread-local message of ddcb200_last_error() and returns `code`
}

#define CUDA_TRY(expr)                                                                                      \
    do {                                                                                                    \
        cudaError_t e_ = (expr);                                                                            \
        if (e_ != cudaSuccess) return ddch::fail(DDCB200_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

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
    unsigned long long* d_dbg = nullptr;   // diagnostic counters (option "dbg_counters")
    std::string last_variant = "none";
    bool smem_attr_set = false;
    // folded fast-FIR taps of the last (step, jt, D): streaming callers repeat the same step call after call
    std::vector<float2> wt_cache;
    double wt_step = 0.0;
    int wt_jt = 0, wt_d = 0;
    // pinned staging for pageable host input (see staged_h2d)
    static constexpr int kStage = 3;
    static constexpr size_t kStageBytes = 8u << 20;
    void* h_stage[kStage] = {};
    cudaEvent_t ev_stage[kStage] = {};
    int stage_pos = 0;
    int copy_threads = 8;                  // option "copy_threads": host threads of the staging / widening (capped by the machine's cores)
    std::unique_ptr<ddch::HostPool> pool;  // copy_threads - 1 parked workers (host_pool.h), created on first use
    float* d_unpack_ws = nullptr;          // float32 workspace of the two-launch packed path (unpack, then a float32 kernel)
    size_t unpack_ws_cap = 0;
    cudaEvent_t unpack_ev = nullptr;       // end of the last kernel that read the workspace (calls may come on different streams)
    // tensor-core engine for packed input (ddc_kernel_tc.cuh): fp16 hi / lo tap-matrix images, device ring + cache key
    static constexpr int kTcRing = 4;
    void* d_tc_b[kTcRing] = {};
    void* h_tc_b[kTcRing] = {};            // pinned
    size_t tc_cap[kTcRing] = {};
    cudaEvent_t tc_ev[kTcRing] = {};       // end of the last kernel that read each image
    cudaEvent_t tc_up_ev = nullptr;        // end of the last upload
    int tc_slot = -1, tc_d = 0;
    double tc_step = 0.0;
    unsigned long long taps_version = 0, tc_version = 0;
    float tc_inv_scale = 1.f, tc_lo_scale = 1.f;
    int tc_na = 0, tc_nraw = 0;            // options "tc_na" / "tc_nraw": force the A-stage / raw-slot counts (tuning)
    int tc_ns = 0, tc_ns_built = 0;        // option "tc_ns": force the sub-stream count (tuning); the one of the cached image
    int host_chunk_mode = 0;               // option "host_chunk_mode": 0 by-stream chunks for batches of short streams, 1 time chunks
    int packed_engine = 1;                 // option "packed_engine": 1 tensor cores where supported (default), 0 CUDA cores
    ddcb200_c64* h_ostage[kBufs] = {};   // pinned landing buffers for the complex128 host path (one per chunk buffer)
    size_t ostage_cap = 0;
};

namespace ddch {
using ddck::RunParams;

// c[k] = taps[T-1-k]/sum * exp(-j 2 pi k step), float64 -> float32; zero padded to n_pad
void make_ctaps(const ddcb200* h, double step, int n_pad, float2* out);
// fast-FIR tap sets (kernel W layout), cached per handle while (step, jt, D) repeat
const float2* cached_wtaps(ddcb200* h, double step, int jt, int D);
// tensor map of the tensor-staged / sliced kernels (ddc_kernel_ws.cuh)
int make_slice_tmap(const float* d_in, int D, bool whole, int slot_rows, long long n_rows, long long n_streams, long long in_stride,
                    CUtensorMap* out);

// ---- kernel launchers, one translation unit per family (k_*.cu); D / jt are checked by the dispatcher -----------------
int launch_tile(ddcb200* h, RunParams& p, const float2* ctaps, cudaStream_t st, int D, int ks);                       // k_tile.cu
int launch_pd(ddcb200* h, RunParams& p, const float2* ctaps, cudaStream_t st, int D, int jt);                         // k_pd.cu
int launch_w(ddcb200* h, RunParams& p, cudaStream_t st, double step, int D, int jt);                                  // k_w.cu
int launch_w10s(ddcb200* h, RunParams& p, cudaStream_t st, double step);                                              // k_w10.cu
bool tc10_supported(const ddcb200* h, int n_taps, int D);                                                                               // k_tc.cu
int launch_tc10(ddcb200* h, RunParams& p, cudaStream_t st, double step, int D);                                       // k_tc.cu
int tc10_describe(int n_taps, int D, int32_t out[12]);                                                                // k_tc.cu
int launch_ws(ddcb200* h, RunParams& p, const float* d_in, long long n_rows, cudaStream_t st, double step, int D, int jt);   // k_ws*.cu
int launch_ws4(ddcb200* h, RunParams& p, const float* d_in, long long n_rows, cudaStream_t st, double step, int jt);
int launch_ws8(ddcb200* h, RunParams& p, const float* d_in, long long n_rows, cudaStream_t st, double step, int jt);
int launch_ws16(ddcb200* h, RunParams& p, const float* d_in, long long n_rows, cudaStream_t st, double step, int jt);
int launch_ws32(ddcb200* h, RunParams& p, const float* d_in, long long n_rows, cudaStream_t st, double step, int jt);
int launch_ws64(ddcb200* h, RunParams& p, const float* d_in, long long n_rows, cudaStream_t st, double step, int jt);
}  // namespace ddch
