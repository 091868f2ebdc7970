/*
 * ddcb200.h -- C ABI of the B200-native fused digital down-converter (NCO mix -> FIR -> decimate).
 *
 * This is the drop-in boundary for the feng/ddc hot path of ska-sa/dc_sand.  The reference has no FFI layer
 * of its own: its boundary is the Python class feng/ddc/src/ddc.py:10-188.  Each entry point below names the
 * reference interface it replaces (paths relative to the reference checkout).  Plain pointers and sizes only;
 * no C++ or torch types cross this boundary; no exceptions: every function returns 0 on success or a negative
 * DDCB200_E* code and records a message retrievable with ddcb200_last_error().
 *
 * Arithmetic contract (reference: feng/ddc/src/cwg.py:31-36, ddc.py:66,98,119):
 *
 *     y[s][m] = rot(m) * sum_{k=0}^{T-1} c[k] * x[s][m*D + k]            m = 0 .. M-1,  M = (N-T)/D + 1
 *     c[k]    = taps[T-1-k] / sum(taps) * exp(-j 2 pi k step)           (float64 on the host, then float32)
 *     rot(m)  = exp(-j 2 pi frac((sample_offset + m*D) * step))          (64-bit fixed-point phase on device)
 *
 * which equals the reference's  convolve(x * exp(-j 2 pi n step), taps, "valid") / sum(taps) [0::D]  with
 * `step` = phase_step_cycles = int(N*fc/fs)/(N-1) computed by the caller exactly as cwg.py:31-33 does.
 * Accumulation is FP32 FMA on CUDA cores; output is interleaved (re, im) float32 ("complex64"), the layout of
 * the reference's own GPU prototype (feng/ddc/src/ddc_host_gpu.py:51,136).
 */
#ifndef DDCB200_H_
#define DDCB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DDCB200_VERSION 100 /* 0.1.0 */

/* status codes */
#define DDCB200_OK 0
#define DDCB200_EINVAL (-1)   /* bad argument (NULL, non-positive size, misaligned pointer ...) */
#define DDCB200_ECUDA (-2)    /* a CUDA runtime call or kernel failed; see ddcb200_last_error() */
#define DDCB200_ENOMEM (-3)   /* host or device allocation failed */
#define DDCB200_ETOOSHORT (-4) /* n_samples < n_taps on an entry point that needs N >= T */

typedef struct ddcb200 ddcb200_t;

/* complex64 element of the output, interleaved re/im float32 (ddc_host_gpu.py:51,136). */
typedef struct ddcb200_c64 {
    float re, im;
} ddcb200_c64;

/* ---- lifetime --------------------------------------------------------------------------------------------
 * Replaces DigitalDownConverter.__init__ (ddc.py:13-31) minus the CSV parsing, which stays in Python
 * (_import_ddc_filter_coeffs, ddc.py:33-48).  `taps` are the raw float64 coefficients in file order; the
 * library normalises by their float64 sum (ddc.py:98).  One handle is bound to one CUDA device and owns one
 * non-blocking CUDA stream plus a small workspace.  Handles are independent; calls on one handle must not
 * overlap in time (same rule as the reference object, which is not thread-safe either). */
int ddcb200_create(ddcb200_t** handle, int device, const double* taps, int n_taps, int decimation);
void ddcb200_destroy(ddcb200_t* handle);

/* Runtime coefficient reload (the reference re-reads a CSV into a new object; ddc.py:33-48). */
int ddcb200_set_taps(ddcb200_t* handle, const double* taps, int n_taps);
int ddcb200_set_decimation(ddcb200_t* handle, int decimation);

/* Output length of run(): ceil((|N - T| + 1) / D); scipy's "valid" convolution swaps operands when N < T
 * (ddc.py:98,119).  Returns 0 for n_samples <= 0. */
int64_t ddcb200_out_len(int64_t n_samples, int n_taps, int decimation);

/* Which kernel family the dispatcher chooses for a call -- a pure function of the filter length, the decimation, the input
 * type and alignment and the two selection options, so the selection table can be checked without a GPU (the reference has
 * one code path, ddc.py:121-188; this is the drop-in's map of it onto kernels).
 *   aligned        rows start on 16-byte boundaries (float32: pointer % 16 == 0 and stride % 4 == 0; packed: stride % 16 == 0)
 *   variant, packed_engine   the options of ddcb200_set_option (0 / 1 are the defaults)
 * Writes the family into `name` ("tensor10", "w10s", "ws", "w", "pd", "tile", "generic"; packed input without a fused-unpack
 * kernel: "unpack+f32:" followed by the float32 family) and returns the tap-block count the kernel is instantiated for (0 where
 * it has none), or a negative DDCB200_E* code. */
int ddcb200_plan(int n_taps, int decimation, int packed, int aligned, int variant, int packed_engine, char* name, int name_cap);

/* Shared-memory pipeline of the tensor-core engine for packed input (the fused form of the stub ddc.py:68-83 followed by
 * ddc.py:51-119) at (n_taps, decimation) with default options -- pure host arithmetic, so the engine's sizing rules can be
 * checked without a GPU.  Returns 1 and fills out12 = {samples per MMA row, accumulator columns N, K, sample stages, raw
 * slots, unpack teams, shared-memory bytes, 16-sample groups per tile, groups the unpack warps cover per tile, sub-stream
 * pitch in bytes, packed bytes copied per tile, tap-matrix bytes}; returns 0 where the engine is not built or does not fit
 * (the dispatcher then takes the CUDA-core kernels); DDCB200_EINVAL for a null pointer. */
int ddcb200_tensor_engine_geometry(int n_taps, int decimation, int32_t* out12);

/* ---- device-resident entry points (inputs and outputs already in HBM) ------------------------------------
 * Replace _mix + _bandpass_fir_filter + _decimate (ddc.py:51-66, 85-100, 102-119) and the NCO generation of
 * cwg.generate_carrier_wave(complex=True) (cwg.py:31-36) for `n_streams` independent 1-D streams.
 *   d_in               device pointer; stream s starts at d_in + s * in_stride (elements: float32 samples)
 *   n_samples          N per stream (>= n_taps, else DDCB200_ETOOSHORT -- see ddcb200_run_short_f32)
 *   phase_step_cycles  NCO cycles per sample, int(N*fc/fs)/(N-1) for a one-shot call (cwg.py:31-33)
 *   sample_offset      index of x[0] within the logical stream (chunked operation); 0 for a one-shot call
 *   d_out              device pointer to complex64; stream s at d_out + s * out_stride; M elements each
 *   cuda_stream        cudaStream_t to launch on, or NULL for the handle's own stream
 * Asynchronous: returns after enqueueing.  Use ddcb200_sync() (or your own stream sync) before reading. */
int ddcb200_run_f32(ddcb200_t* handle, const float* d_in, int64_t n_samples, int64_t n_streams,
                    int64_t in_stride, double phase_step_cycles, int64_t sample_offset, ddcb200_c64* d_out,
                    int64_t out_stride, void* cuda_stream);

/* Same, with the digitiser's packed 10-bit transport format fused into the load path.  Replaces the stub
 * _decode_8bit_to_10bit_to_float_data (ddc.py:68-83) + the three stages above.  Format (defined by this
 * library): sample k is a two's-complement 10-bit integer, MSB first, at bit offset 10*k of a
 * big-endian bit stream; 4 samples per 5 bytes.  `in_stride_bytes` separates streams; n_samples must be a
 * multiple of 4. */
int ddcb200_run_packed10(ddcb200_t* handle, const uint8_t* d_in, int64_t n_samples, int64_t n_streams,
                         int64_t in_stride_bytes, double phase_step_cycles, int64_t sample_offset,
                         ddcb200_c64* d_out, int64_t out_stride, void* cuda_stream);

/* Stand-alone unpack stage (ddc.py:68-83), bit-exact integer work: packed bytes -> int16 and/or float32
 * (either output pointer may be NULL). */
int ddcb200_unpack10(ddcb200_t* handle, const uint8_t* d_in, int64_t n_samples, int16_t* d_out_i16,
                     float* d_out_f32, void* cuda_stream);

/* Inverse of the unpack stage for test vectors built in HBM (the reference has neither a packer nor an unpacker, only the
 * stub ddc.py:68-83; the format is this library's, see above): float32 samples are rounded to nearest, clipped to
 * [-512, 511] and written 4 samples -> 5 bytes.  Rows of n_streams; strides in float32 elements / bytes. */
int ddcb200_pack10(ddcb200_t* handle, const float* d_in, int64_t n_samples, int64_t n_streams, int64_t in_stride,
                   uint8_t* d_out, int64_t out_stride_bytes, void* cuda_stream);

/* N < T corner of the reference: scipy swaps the operands, so run() returns
 *     y[i] = (1/sum(taps)) * sum_n mix[n] * taps[i*D' ... ]   -- precisely: full[i] = sum_n mix[n]*taps[i+N-1-n],
 * i = 0..T-N, then [0::D].  Kept for drop-in fidelity (ddc.py:98 with len(mix) < len(taps)); single stream. */
int ddcb200_run_short_f32(ddcb200_t* handle, const float* d_in, int64_t n_samples, double phase_step_cycles,
                          int64_t sample_offset, ddcb200_c64* d_out, void* cuda_stream);

/* ---- stage entry points (device pointers) ------------------------------------------------------------------------
 * The reference exposes its stages as methods; run() here is fused, these let the drop-in class keep the stage methods
 * on the GPU: _mix (ddc.py:51-66), _bandpass_fir_filter (ddc.py:85-100: full-rate "valid" FIR / sum(taps), complex64
 * in, n_in - n_taps + 1 complex64 out), _decimate (ddc.py:102-119: z[offset::D], ceil((n_in - offset) / D) out). */
int ddcb200_mix_f32(ddcb200_t* handle, const float* d_x, const ddcb200_c64* d_cw, ddcb200_c64* d_out, int64_t n,
                    void* cuda_stream);
int ddcb200_fir_c64(ddcb200_t* handle, const ddcb200_c64* d_in, int64_t n_in, ddcb200_c64* d_out, void* cuda_stream);
int ddcb200_decimate_c64(ddcb200_t* handle, const ddcb200_c64* d_in, int64_t n_in, int64_t offset, ddcb200_c64* d_out,
                         void* cuda_stream);

/* ---- test-signal generator on the device ----------------------------------------------------------------------------
 * Replaces cwg.generate_carrier_wave (feng/ddc/src/cwg.py:6-44), the reference's test-vector source, for inputs that
 * should never touch the host: sample n of stream s = cw_scale * exp(-j 2 pi (phase0_cycles + (sample_offset + n) *
 * phase_step_cycles)), float32 real part (is_complex = 0) or complex64.  noise_mode 0: none; 1: + noise_scale *
 * truncated normal on [-1, 1], sigma 0.5, on the real part (cwg.py:47-70; the reference's draw is unseeded, this one is
 * Philox4x32-10 keyed by (seed, stream), counter = sample index); 2: digitiser model, real part + noise_scale * N(0,1),
 * rounded and clipped to [-512, 511].  phase_step_cycles = int(N f / fs) / (N - 1) reproduces the reference's linspace. */
int ddcb200_cwg(ddcb200_t* handle, void* d_out, int64_t num_samples, int64_t n_streams, int64_t out_stride, int is_complex,
                double cw_scale, double phase_step_cycles, double phase0_cycles, int64_t sample_offset, int noise_mode,
                double noise_scale, uint64_t seed, void* cuda_stream);

/* ---- host-buffer entry points (what DigitalDownConverter.run binds to) -------------------------------------
 * Replace DigitalDownConverter.run (ddc.py:121-188) for host arrays: time-chunked, double-buffered
 * H2D -> fused kernel -> D2H on two CUDA streams, synchronous on return.  Pinned host memory (see
 * ddcb200_host_alloc) makes the copies truly asynchronous; pageable memory works but is slower.
 * N < T is routed to the short path when n_streams == 1, else DDCB200_ETOOSHORT. */
int ddcb200_run_host_f32(ddcb200_t* handle, const float* h_in, int64_t n_samples, int64_t n_streams,
                         int64_t in_stride, double phase_step_cycles, int64_t sample_offset,
                         ddcb200_c64* h_out, int64_t out_stride);
/* One stream, result as complex128 (interleaved re, im float64) -- the dtype DigitalDownConverter.run returns
 * (ddc.py:98,119: scipy.signal.convolve of a complex64 array with float64 taps).  The arithmetic is the complex64 path's;
 * the widening happens on host threads straight into the caller's array while the next chunk is in flight. */
int ddcb200_run_host_f32_c128(ddcb200_t* handle, const float* h_in, int64_t n_samples, double phase_step_cycles,
                              int64_t sample_offset, double* h_out);
int ddcb200_run_host_packed10(ddcb200_t* handle, const uint8_t* h_in, int64_t n_samples, int64_t n_streams,
                              int64_t in_stride_bytes, double phase_step_cycles, int64_t sample_offset,
                              ddcb200_c64* h_out, int64_t out_stride);

/* ---- streaming sessions -------------------------------------------------------------------------------------------
 * The reference processes one finite array per run() (ddc.py:121); a digitiser delivers an endless stream (cf. the
 * reference's ibverbs receiver, ibverbs_sample_project/ibverbs_rx.c:282-327).  A session carries the last T-D .. T-1
 * samples of every stream and the absolute sample index across pushes, so the concatenated outputs of any sequence of
 * pushes equal one run() over the concatenated input (same NCO phase law: pass the phase_step_cycles of that run()).
 * push_f32: device pointers, asynchronous on `cuda_stream`; push_host_f32: host pointers, pieces of max_chunk_samples
 * double-buffered (H2D of piece i+1 under the kernel of piece i), synchronous on return.  *n_out = outputs per stream. */
typedef struct ddcb200_session ddcb200_session_t;
int ddcb200_session_open(ddcb200_t* handle, int64_t n_streams, int64_t max_chunk_samples, double phase_step_cycles,
                        ddcb200_session_t** out);
void ddcb200_session_close(ddcb200_session_t* s);
int ddcb200_session_reset(ddcb200_session_t* s, int64_t first_sample_index);
int64_t ddcb200_session_pending(ddcb200_session_t* s);    /* samples carried per stream */
int64_t ddcb200_session_position(ddcb200_session_t* s);   /* absolute index of the next sample to be pushed */
int64_t ddcb200_session_out_len(ddcb200_session_t* s, int64_t n_samples);   /* outputs the next push of n_samples will give */
int ddcb200_session_push_f32(ddcb200_session_t* s, const float* d_in, int64_t n_samples, int64_t in_stride,
                            ddcb200_c64* d_out, int64_t out_stride, int64_t* n_out, void* cuda_stream);
int ddcb200_session_push_host_f32(ddcb200_session_t* s, const float* h_in, int64_t n_samples, int64_t in_stride,
                                 ddcb200_c64* h_out, int64_t out_stride, int64_t* n_out);
/* The same for packed 10-bit input (the stub ddc.py:68-83 fused in): pushes are whole groups of 4 samples (5 bytes),
 * strides are in bytes, the decimation must be a multiple of 4. */
int ddcb200_session_open_packed10(ddcb200_t* handle, int64_t n_streams, int64_t max_chunk_samples,
                                  double phase_step_cycles, ddcb200_session_t** out);
int ddcb200_session_push_packed10(ddcb200_session_t* s, const uint8_t* d_in, int64_t n_samples, int64_t in_stride_bytes,
                                  ddcb200_c64* d_out, int64_t out_stride, int64_t* n_out, void* cuda_stream);
int ddcb200_session_push_host_packed10(ddcb200_session_t* s, const uint8_t* h_in, int64_t n_samples,
                                       int64_t in_stride_bytes, ddcb200_c64* h_out, int64_t out_stride, int64_t* n_out);

/* Pinned host memory helpers (the reference's prototype uses cuda.pagelocked_empty, ddc_host_gpu.py:45-58). */
void* ddcb200_host_alloc(size_t bytes);
void ddcb200_host_free(void* p);

/* ---- misc ------------------------------------------------------------------------------------------------ */
int ddcb200_sync(ddcb200_t* handle);            /* waits for the handle's own stream */
void* ddcb200_stream(ddcb200_t* handle);        /* the handle's cudaStream_t */
const char* ddcb200_last_error(void);           /* thread-local message of the last failure */
int ddcb200_version(void);
/* Number of kernels of this library launched through `handle` so far (bench.py's gpu_launches). */
int64_t ddcb200_launch_count(ddcb200_t* handle);
/* Name of the kernel variant the last run on this handle dispatched to (e.g. "fused_tma<D16,R4,T256>"). */
const char* ddcb200_last_variant(ddcb200_t* handle);
/* Tuning/diagnostic knobs (all optional; 0 restores the default):
 *   "variant"        kernel choice: 0 auto, 1 generic, 2 / 3 tile kernel without / with the tap split, 7 fast FIR on 1-D bulk
 *                    copies (D = 16), 8 phase-major direct form (D = 32 / 64), 10 warp-specialised CUDA-core packed kernel,
 *                    11 tensor-staged / sliced fast FIR, 13 tensor-core engine for packed input (see DESIGN.md section 4
 *                    and tools/README.md); a choice that is not built for the call's (taps, decimation) falls through to auto;
 *   "chunk_samples"  time-chunk size of the host path (default 2^24 samples per launch);
 *   "copy_threads"   host threads that stage pageable input through pinned buffers and widen complex128 output (a pool
 *                    parked in the handle; default 8, never more than the machine's cores; 0 = leave pageable copies to the
 *                    driver);
 *   "packed_engine"  engine for packed 10-bit input: 1 (default) the tcgen05 tensor-core engine wherever it applies
 *                    (decimation 4 .. 64 a power of two, 16-byte aligned rows, filter fits shared memory), 0 the CUDA-core
 *                    kernels; float32 input always runs on the CUDA cores;
 *   "tc_ns", "tc_na", "tc_nraw"   tuning of the tensor engine (row width, A stages, raw slots; 0 = automatic);
 *   "debug_mode", "dbg_counters"   measurement aids of the kernels (compute-only / memory-only
 *                    ceilings, ring wait-time counters). */
int ddcb200_set_option(ddcb200_t* handle, const char* key, int64_t value);

#ifdef __cplusplus
}
#endif
#endif /* DDCB200_H_ */
