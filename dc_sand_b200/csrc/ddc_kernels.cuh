// Kernels of the fused digital down-converter (sm_100a).
//
//   ddc_fused_kernel   persistent, warp-specialised: one producer warp stages input tiles (+ tap halo) in shared
//                      memory with 1-D TMA bulk copies behind an mbarrier ring; compute warps run the polyphase
//                      decimating FIR with the NCO folded into complex taps (FFMA2, taps through uniform
//                      registers), rotate each output by the NCO phase of its first sample and store complex64.
//   ddc_generic_kernel one thread per output, any N/T/D, unaligned pointers, packed or float input: stream tails,
//                      odd decimation factors, and the cross-check of the fused kernel in the tests.
//   ddc_short_kernel   N < T corner of the reference (scipy swaps the operands), per-sample NCO.
//   unpack10_kernel    stand-alone unpack stage (bit-exact integer work).
#pragma once
#include "ddc_common.cuh"

namespace ddck {

// =============================================================================================================
// Generic kernel
// =============================================================================================================
template <bool PACKED>
__device__ __forceinline__ float load_sample(const void* __restrict__ base, long long n) {
    if (PACKED) return (float)unpack10_at(reinterpret_cast<const uint8_t*>(base), n);
    return __ldg(reinterpret_cast<const float*>(base) + n);
}

template <bool PACKED>
__global__ void __launch_bounds__(128) ddc_generic_kernel(const RunParams p, const float2* __restrict__ ctaps, int decim) {
    const long long m = p.m_begin + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int s = blockIdx.y;
    if (m >= p.n_out) return;
    const char* in_b = reinterpret_cast<const char*>(p.in) + (PACKED ? p.in_stride : p.in_stride * 4) * s;
    const long long n0 = m * decim;
    float re = 0.f, im = 0.f;
    for (int k = 0; k < p.n_taps; ++k) {
        const float x = load_sample<PACKED>(in_b, n0 + k);
        const float2 c = __ldg(ctaps + k);
        re = fmaf(x, c.x, re);
        im = fmaf(x, c.y, im);
    }
    const float2 rot = nco_rot(p.phase0_fx + (unsigned long long)n0 * p.step_fx);
    p.out[(long long)s * p.out_stride + m] = cmul(make_float2(re, im), rot);
}

// N < T: full[i] = sum_n mix[n] * h[i + N-1-n], i = 0..T-N, decimated [0::D]; h = taps / sum(taps) (float32)
__global__ void ddc_short_kernel(const RunParams p, const float* __restrict__ h, int n_taps_real, int decim) {
    const long long i_out = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i_out >= p.n_out) return;
    const long long i = i_out * decim;
    const float* x = reinterpret_cast<const float*>(p.in);
    const int n_samp = (int)p.n_samples;
    float re = 0.f, im = 0.f;
    for (int n = 0; n < n_samp; ++n) {
        const float2 cw = nco_rot(p.phase0_fx + (unsigned long long)n * p.step_fx);
        const float xv = x[n];
        const float hv = h[i + n_samp - 1 - n];
        re = fmaf(xv * cw.x, hv, re);
        im = fmaf(xv * cw.y, hv, im);
    }
    p.out[i_out] = make_float2(re, im);
}

__global__ void unpack10_kernel(const uint8_t* __restrict__ in, long long n_groups, int16_t* __restrict__ o16,
                                float* __restrict__ of32) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    const uint8_t* b = in + g * 5;
    const uint32_t b0 = b[0];
    const uint32_t lo = ((uint32_t)b[1] << 24) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 8) | (uint32_t)b[4];
    int v[4];
    unpack10_word(b0, lo, v);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (o16) o16[g * 4 + k] = (int16_t)v[k];
        if (of32) of32[g * 4 + k] = (float)v[k];
    }
}

// unpack stage for [streams][5 N / 4] packed rows -> [streams][pitch] float32 rows: the first half of the two-launch path that
// serves packed input for tap / decimation combinations without a fused-unpack kernel
__global__ void unpack10_rows_kernel(const uint8_t* __restrict__ in, long long in_stride_bytes, long long n_groups,
                                     float* __restrict__ out, long long out_stride) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    const uint8_t* b = in + (long long)blockIdx.y * in_stride_bytes + g * 5;
    const uint32_t b0 = b[0];
    const uint32_t lo = ((uint32_t)b[1] << 24) | ((uint32_t)b[2] << 16) | ((uint32_t)b[3] << 8) | (uint32_t)b[4];
    int v[4];
    unpack10_word(b0, lo, v);
    *reinterpret_cast<float4*>(out + (long long)blockIdx.y * out_stride + g * 4) =
        make_float4((float)v[0], (float)v[1], (float)v[2], (float)v[3]);
}

// Inverse of the unpack stage, for test-vector generation in HBM (ddcb200_pack10): float32 samples are rounded, clipped to
// [-512, 511] and written in the transport format, 4 samples -> 5 bytes, rows of [streams].
__global__ void pack10_rows_kernel(const float* __restrict__ in, long long in_stride, long long n_groups, uint8_t* __restrict__ out,
                                   long long out_stride_bytes) {
    const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    const float* x = in + (long long)blockIdx.y * in_stride + g * 4;
    uint32_t u[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        int v = __float2int_rn(x[k]);
        v = v < -512 ? -512 : (v > 511 ? 511 : v);
        u[k] = (uint32_t)v & 0x3FFu;
    }
    uint8_t* b = out + (long long)blockIdx.y * out_stride_bytes + g * 5;
    b[0] = (uint8_t)(u[0] >> 2);
    b[1] = (uint8_t)(((u[0] & 0x3u) << 6) | (u[1] >> 4));
    b[2] = (uint8_t)(((u[1] & 0xFu) << 4) | (u[2] >> 6));
    b[3] = (uint8_t)(((u[2] & 0x3Fu) << 2) | (u[3] >> 8));
    b[4] = (uint8_t)(u[3] & 0xFFu);
}

// Stage kernels: the reference exposes its three stages as separate methods (ddc.py:51-66, 85-100, 102-119).  The fused
// kernels above are what run() uses; these exist so that the stage methods of the drop-in class also execute on the GPU.
__global__ void mix_kernel(const float* __restrict__ x, const float2* __restrict__ cw, float2* __restrict__ out, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const float xv = x[i];
        const float2 c = cw[i];
        out[i] = make_float2(xv * c.x, xv * c.y);   // float32 * complex64 -> complex64 (ddc.py:66)
    }
}

// full-rate "valid" FIR of a complex64 sequence with real taps: y[n] = sum_k h[k] z[n + k], h = reversed taps / sum
__global__ void fir_c64_kernel(const float2* __restrict__ z, const float* __restrict__ h, int n_taps, float2* __restrict__ out,
                               long long n_out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_out) return;
    float re = 0.f, im = 0.f;
    for (int k = 0; k < n_taps; ++k) {
        const float2 v = __ldg(z + i + k);
        const float hk = __ldg(h + k);
        re = fmaf(v.x, hk, re);
        im = fmaf(v.y, hk, im);
    }
    out[i] = make_float2(re, im);
}

__global__ void decimate_c64_kernel(const float2* __restrict__ z, long long offset, int decim, float2* __restrict__ out,
                                    long long n_out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_out) out[i] = z[offset + i * decim];
}

// =============================================================================================================
// Carrier-wave / test-signal generator on the device (reference: feng/ddc/src/cwg.py:6-70, the test-vector source of the
// reference's own tests).  sample n = cw_scale * exp(-j 2 pi (phase0 + n step)) [+ noise on the real part]
//   noise_mode 0: none
//   noise_mode 1: noise_scale * truncated normal on [-1, 1], sigma 0.5        (cwg._generate_noise, cwg.py:47-70)
//   noise_mode 2: noise_scale * N(0, 1), then round-to-nearest and clip to the 10-bit range [-512, 511]   (digitiser model)
// Noise comes from Philox4x32-10 keyed by (seed, stream) with the sample index as counter: reproducible, order-free.
// =============================================================================================================
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
    constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
        const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += W0;
        key.y += W1;
    }
    return ctr;
}

// two independent N(0,1) from two 32-bit words (Box-Muller); u1 in (0, 1]
__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
    const float u1 = ((float)(a >> 8) + 1.0f) * (1.0f / 16777216.0f);
    const float u2 = (float)(b >> 8) * (1.0f / 16777216.0f);
    const float rad = sqrtf(-2.0f * __logf(u1));
    float sn, cs;
    sincospif(2.0f * u2, &sn, &cs);
    return make_float2(rad * cs, rad * sn);
}

template <bool COMPLEX>
__global__ void cwg_kernel(void* __restrict__ out, long long n, long long out_stride, float cw_scale, unsigned long long step_fx,
                           unsigned long long phase0_fx, int noise_mode, float noise_scale, unsigned long long seed) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned stream = blockIdx.y;
    if (i >= n) return;
    const float2 cw = ddck::nco_rot(phase0_fx + (unsigned long long)i * step_fx);   // exp(-j 2 pi phase)
    float re = cw_scale * cw.x, im = cw_scale * cw.y;
    if (noise_mode != 0) {
        const uint2 key = make_uint2((uint32_t)seed ^ (stream * 0x9E3779B9u), (uint32_t)(seed >> 32) + stream);
        float g = 0.f;
        if (noise_mode == 1) {
            // truncated normal by rejection (95.4 % acceptance); attempt number in counter word w
            bool ok = false;
            for (uint32_t att = 0; att < 16 && !ok; ++att) {
                const uint4 r = philox4x32_10(make_uint4((uint32_t)i, (uint32_t)((unsigned long long)i >> 32), 0x7A11u, att), key);
                const float2 z0 = box_muller(r.x, r.y), z1 = box_muller(r.z, r.w);
                const float c[4] = {0.5f * z0.x, 0.5f * z0.y, 0.5f * z1.x, 0.5f * z1.y};
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (!ok && fabsf(c[k]) <= 1.0f) {
                        g = c[k];
                        ok = true;
                    }
            }
            re += noise_scale * g;
        } else {
            const uint4 r = philox4x32_10(make_uint4((uint32_t)i, (uint32_t)((unsigned long long)i >> 32), 0xD161u, 0u), key);
            g = box_muller(r.x, r.y).x;
            re = fminf(fmaxf(rintf(re + noise_scale * g), -512.f), 511.f);
            im = 0.f;
        }
    }
    if (COMPLEX)
        reinterpret_cast<float2*>(out)[(long long)stream * out_stride + i] = make_float2(re, im);
    else
        reinterpret_cast<float*>(out)[(long long)stream * out_stride + i] = re;
}

// =============================================================================================================
// Fused persistent kernel
// =============================================================================================================
// Shared-memory layout of one pipeline stage.  A "thread-row" is the ROW = R*D float32 samples that produce R
// consecutive outputs (256 B for every supported D).  S consecutive thread-rows form a "super-row" that is staged by
// ONE bulk copy (S*ROW*4 bytes, 8 KB for S = 32) and is followed by a 16-byte pad, i.e. super-row pitch = S*ROW + 4
// floats.  A tile is NROWS = 256 thread-rows (+ halo).  Within a warp, lane bits [0,3) = i select one of eight
// consecutive super-rows, so the eight lanes of a quarter warp issue LDS.128 at addresses that differ by
// (S*ROW + 4) floats = 16 B mod 128 B: eight distinct bank groups, no bank conflicts, with dense TMA-friendly rows
// (a tile + halo is 9 bulk copies).
//
// KS = tap split: the T taps of one thread-row are shared by KS threads, lane l of warps w and w + 8 (the half must be
// warp-uniform so that taps stay uniform-register operands); each accumulates J/KS tap blocks for the same R outputs,
// the halves are exchanged through a small double-buffered shared-memory area under a 64-thread named barrier, and
// each thread then rotates and stores R/KS outputs.  KS = 2 doubles the resident compute warps (16 per SM) for the
// same input staging footprint.
template <int D, int R, int S, int KS>
struct FusedCfg {
    static constexpr int ROW = R * D;            // samples per thread-row
    static constexpr int SRP = S * ROW + 4;      // super-row pitch in floats
    static constexpr int V = D / 4;              // float4 per tap block
    static constexpr int NROWS = 256;            // thread-rows per tile
    static constexpr int QN = 4;                 // thread-rows per warp at the same super-row
    static constexpr int WPG = S / QN;           // warps per group of 8 super-rows
    static constexpr int SENDN = (R >= 2) ? R / 2 : 1;  // float2 exchanged per thread when KS == 2
    static constexpr size_t XBUF_BYTES = (KS == 2) ? (size_t)2 * 2 * NROWS * SENDN * sizeof(float2) : 0;
    static constexpr int NT = NROWS * KS;        // compute threads
    static constexpr int TILE_OUT = NROWS * R;   // outputs per tile
    static_assert(D % 4 == 0, "D must be a multiple of 4");
    static_assert(KS == 1 || KS == 2, "tap split 1 or 2");
    static_assert(S == 4 || S == 8 || S == 16 || S == 32, "S must be 4, 8, 16 or 32");
    static_assert((S * ROW / 4) % 8 == 0, "super-row must be a whole number of 128-byte lines");
    static_assert(NROWS % (8 * S) == 0, "tile must be a whole number of 8-super-row groups");
    __host__ __device__ static constexpr int row_offset(int row) { return (row / S) * SRP + (row % S) * ROW; }
    __host__ __device__ static constexpr size_t stage_floats(int halo_rows) {
        // whole super-rows of the tile and of the halo, plus a last partial super-row (16-byte pad kept)
        return (size_t)(NROWS / S + halo_rows / S) * SRP + (size_t)((halo_rows % S) ? (halo_rows % S) * ROW + 4 : 0);
    }
};

template <int D, int R, int S, int KS, int STAGES, int MAXT, bool PACKED>
__global__ void __launch_bounds__(FusedCfg<D, R, S, KS>::NT + 32, 1)
ddc_fused_kernel(const __grid_constant__ RunParams p, const __grid_constant__ TapsParam<MAXT> taps) {
    using C = FusedCfg<D, R, S, KS>;
    constexpr int ROW = C::ROW, SRP = C::SRP, V = C::V, NT = C::NT, NROWS = C::NROWS;
    constexpr int TILE_OUT = C::TILE_OUT;
    constexpr int TILE_S = NROWS * ROW;  // samples per tile

    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw);
    uint64_t* empty_bar = full_bar + STAGES;
    float2* xbuf = reinterpret_cast<float2*>(smem_raw + 128);  // [parity][half][NROWS][SENDN], KS == 2 only
    float* buf = reinterpret_cast<float*>(smem_raw + 128 + C::XBUF_BYTES);
    const int stage_floats = (int)C::stage_floats(p.halo_rows);

    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // warp-uniform by construction: lets ptxas keep loop state in uniform registers
    const int lane = tid & 31;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], NT / 32);
        }
        mbar_fence_init();
    }
    __syncthreads();

    // Tiles are numbered stream-major; a CTA walks tile = blockIdx.x, += gridDim.x.  (stream, tile-in-stream) are
    // advanced incrementally (no 64-bit division on the per-tile path).
    const int tps = (int)p.tiles_per_stream;
    const int gstep_s = (int)(gridDim.x / tps), gstep_t = (int)(gridDim.x % tps);
    int cur_s = (int)(blockIdx.x / tps), cur_t = (int)(blockIdx.x % tps);
    const int n_iter = (int)((p.total_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);  // blockIdx.x < total_tiles

    if (warp == NT / 32) {
        // ------------------------------------------------ producer warp
        const int n_sr = NROWS / S + (p.halo_rows + S - 1) / S;  // super-rows per stage
        const long long want = (long long)(NROWS + p.halo_rows) * ROW;
        for (int it = 0; it < n_iter && p.debug_mode != 1 && p.debug_mode != 3; ++it) {
            const int stage = it % STAGES;
            const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
            mbar_wait(&empty_bar[stage], ph ^ 1u);
            const float* src = reinterpret_cast<const float*>(p.in) + (long long)cur_s * p.in_stride + (long long)cur_t * TILE_S;
            float* dst = buf + (size_t)stage * stage_floats;
            const long long valid = p.n_samples - (long long)cur_t * TILE_S;  // samples of this stream from the tile start
            if (valid >= want) {
                // full tile: one elected lane issues the TMA bulk copies (8 KB each)
                if (lane == 0) {
                    mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)want * 4u);
                    int left = NROWS + p.halo_rows;
#pragma unroll 1
                    for (int sr = 0; left > 0; ++sr, left -= S) {
                        const int nrow = left < S ? left : S;
                        bulk_g2s(dst + sr * SRP, src + (size_t)sr * S * ROW, (uint32_t)nrow * ROW * 4u, &full_bar[stage]);
                    }
                }
            } else {
                // ragged last tile of a stream: bulk-copy what is whole 16-byte groups, hand-copy the last 1-3 samples,
                // zero-fill the rest (zero taps of a padded tap set must not meet stale shared memory)
                uint32_t tx = 0;
                for (int sr = 0; sr < n_sr; ++sr) {
                    const int left = NROWS + p.halo_rows - sr * S;
                    const int cap = (left < S ? left : S) * ROW;          // floats this super-row holds
                    const long long s0 = (long long)sr * S * ROW;
                    long long cnt = valid - s0;
                    cnt = cnt < 0 ? 0 : (cnt > cap ? cap : cnt);
                    const int bulk = (int)cnt & ~3;
                    for (int k = bulk + lane; k < cap; k += 32) dst[sr * SRP + k] = (k < (int)cnt) ? src[s0 + k] : 0.f;
                    tx += (uint32_t)bulk * 4u;
                }
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive_expect_tx(&full_bar[stage], tx);  // release: the plain stores above become visible
                    for (int sr = 0; sr < n_sr; ++sr) {
                        const int left = NROWS + p.halo_rows - sr * S;
                        const int cap = (left < S ? left : S) * ROW;
                        const long long s0 = (long long)sr * S * ROW;
                        long long cnt = valid - s0;
                        cnt = cnt < 0 ? 0 : (cnt > cap ? cap : cnt);
                        const int bulk = (int)cnt & ~3;
                        if (bulk > 0) bulk_g2s(dst + sr * SRP, src + s0, (uint32_t)bulk * 4u, &full_bar[stage]);
                    }
                }
            }
            __syncwarp();
            cur_s += gstep_s;
            cur_t += gstep_t;
            if (cur_t >= tps) { cur_t -= tps; ++cur_s; }
        }
    } else {
        // ------------------------------------------------ compute warps
        const int JH = p.n_tap_blocks / KS;                  // tap blocks per thread
        const int li = lane & 7;
        const int lq = lane >> 3;
        const int half = (KS == 2) ? (warp / (NROWS / 32)) : 0;   // warp-uniform
        const int wr = warp % (NROWS / 32);
        const int g = (wr / C::WPG) * (8 * S) + li * S + (wr % C::WPG) * C::QN + lq;  // thread-row within the tile
        const int jbeg = half * JH;                          // first tap block of this thread (multiple of R)
        const int row0 = g + jbeg / R;
        constexpr int RO = (KS == 2 && R >= 2) ? R / 2 : R;  // outputs this thread finishes
        const int rbeg = (KS == 2 && R >= 2) ? half * RO : 0;
        // NCO rotation of output m = tile*TILE_OUT + g*R + rbeg + r is rot_tile(tile) * rot_thr[r]: the per-thread factor
        // is computed once per kernel, the per-tile factor once per tile (one sincospif instead of R per tile).
        float2 rot_thr[RO];
#pragma unroll
        for (int r = 0; r < RO; ++r)
            rot_thr[r] = nco_rot((unsigned long long)((long long)(g * R + rbeg + r) * D) * p.step_fx);
        const unsigned long long tile_dph = (unsigned long long)((long long)TILE_OUT * D) * p.step_fx;

        // The two (KS = 1) or four (KS = 2) compute warps that share an SM sub-partition do identical work and would reach
        // their epilogues together, leaving the FMA pipe idle; starting every other one half a tile late keeps one
        // warp in its FMA loop while its neighbour rotates and stores.
        if (p.stagger_cycles > 0 && ((warp >> 2) & 1)) {
            const long long t0 = clock64();
            while (clock64() - t0 < p.stagger_cycles) {}
        }

        for (int it = 0; it < n_iter; ++it) {
            const int stage = it % STAGES;
            const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
            if (p.debug_mode != 1 && p.debug_mode != 3) mbar_wait(&full_bar[stage], ph);
            const float* sbuf = buf + (size_t)stage * stage_floats;

#ifndef DDCB200_SPLIT_ACC
#define DDCB200_SPLIT_ACC 0
#endif
            constexpr int NA = DDCB200_SPLIT_ACC ? 2 : 1;  // partial sums per output (more independent FMA chains)
            float2 accp[R][NA];
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int a = 0; a < NA; ++a) accp[r][a] = make_float2(0.f, 0.f);
            float4 xw[R][V];  // rotating window: slot (j + r) % R holds tap-block-sized sample block j + r
            // Row pointers advance by a thread-local recurrence (not a function of j0) so that the compiler keeps the
            // tap-block counter j0 in a uniform register and fetches taps with LDCU -> UR operands of FFMA2.
            const float* p0 = sbuf + C::row_offset(row0);
            int sub = row0 % S;
#pragma unroll
            for (int s = 0; s < R - 1; ++s)
#pragma unroll
                for (int v = 0; v < V; ++v) xw[s][v] = *reinterpret_cast<const float4*>(p0 + s * D + 4 * v);
            const float4* tbase = &taps.c2[(size_t)jbeg * (D / 2)];
            const int jend = (p.debug_mode == 2) ? 0 : JH;
            for (int j0 = 0; j0 < jend; j0 += R) {
                const bool wrap = (sub == S - 1);
                const float* p1 = p0 + ROW + (wrap ? 4 : 0);
                sub = wrap ? 0 : sub + 1;
#pragma unroll
                for (int jj = 0; jj < R; ++jj) {
                    const int srel = jj + R - 1;  // newest block of this step, relative to block j0 (row p0)
                    const float* src = (srel / R) ? p1 : p0;
                    if (p.debug_mode != 3) {  // 3: compute only AND no shared-memory loads in the loop (FMA ceiling)
#pragma unroll
                        for (int v = 0; v < V; ++v)
                            xw[srel % R][v] = *reinterpret_cast<const float4*>(src + (srel % R) * D + 4 * v);
                    }
                    const float4* tp = tbase + (size_t)(j0 + jj) * (D / 2);
#pragma unroll
                    for (int v = 0; v < V; ++v) {
                        const float4 ta = tp[2 * v], tb = tp[2 * v + 1];
#pragma unroll
                        for (int r = 0; r < R; ++r) accp[r][0] = ffma2(xw[(jj + r) % R][v].x, make_float2(ta.x, ta.y), accp[r][0]);
#pragma unroll
                        for (int r = 0; r < R; ++r) accp[r][NA - 1] = ffma2(xw[(jj + r) % R][v].y, make_float2(ta.z, ta.w), accp[r][NA - 1]);
#pragma unroll
                        for (int r = 0; r < R; ++r) accp[r][0] = ffma2(xw[(jj + r) % R][v].z, make_float2(tb.x, tb.y), accp[r][0]);
#pragma unroll
                        for (int r = 0; r < R; ++r) accp[r][NA - 1] = ffma2(xw[(jj + r) % R][v].w, make_float2(tb.z, tb.w), accp[r][NA - 1]);
                    }
                }
                p0 = p1;
            }
            float2 acc[R];
#pragma unroll
            for (int r = 0; r < R; ++r)
                acc[r] = (NA == 2) ? make_float2(accp[r][0].x + accp[r][NA - 1].x, accp[r][0].y + accp[r][NA - 1].y) : accp[r][0];
            // all shared-memory reads of this stage are done -> hand the slot back to the producer
            __syncwarp();
            if (lane == 0 && p.debug_mode != 1 && p.debug_mode != 3) mbar_arrive(&empty_bar[stage]);

            // epilogue: combine tap halves, rotate each output by the NCO phase of its first input sample, store
            float2 y[RO];
            bool writer = true;
            if (KS == 2) {
                constexpr int SN = C::SENDN;
                const int xt = wr * 32 + lane;
                float2* xs = xbuf + ((size_t)((it & 1) * 2 + half) * NROWS + xt) * SN;         // what I send
                const float2* xr = xbuf + ((size_t)((it & 1) * 2 + (half ^ 1)) * NROWS + xt) * SN;  // what my partner sent
                if (R >= 2) {
                    // this thread keeps outputs [half*RO, half*RO + RO) and receives the partner's partial sums for them
                    if (SN % 2 == 0) {
#pragma unroll
                        for (int r = 0; r < SN; r += 2) {
                            const float2 a = half ? acc[r] : acc[r + RO], b = half ? acc[r + 1] : acc[r + 1 + RO];
                            *reinterpret_cast<float4*>(xs + r) = make_float4(a.x, a.y, b.x, b.y);
                        }
                    } else {
#pragma unroll
                        for (int r = 0; r < SN; ++r) xs[r] = half ? acc[r] : acc[r + RO];
                    }
                    asm volatile("bar.sync %0, 64;" ::"r"(1 + wr) : "memory");
                    if (SN % 2 == 0) {
#pragma unroll
                        for (int r = 0; r < SN; r += 2) {
                            const float4 v = *reinterpret_cast<const float4*>(xr + r);
                            const float2 k0 = half ? acc[r + RO] : acc[r], k1 = half ? acc[r + 1 + RO] : acc[r + 1];
                            y[r] = make_float2(k0.x + v.x, k0.y + v.y);
                            y[r + 1] = make_float2(k1.x + v.z, k1.y + v.w);
                        }
                    } else {
#pragma unroll
                        for (int r = 0; r < SN; ++r) {
                            const float2 v = xr[r];
                            const float2 k0 = half ? acc[r + RO] : acc[r];
                            y[r] = make_float2(k0.x + v.x, k0.y + v.y);
                        }
                    }
                } else {
                    if (half) xs[0] = acc[0];
                    asm volatile("bar.sync %0, 64;" ::"r"(1 + wr) : "memory");
                    const float2 v = half ? make_float2(0.f, 0.f) : xr[0];
                    y[0] = make_float2(acc[0].x + v.x, acc[0].y + v.y);
                    writer = (half == 0);
                }
            } else {
#pragma unroll
                for (int r = 0; r < RO; ++r) y[r] = acc[r];
            }
            const float2 rot_tile = nco_rot(p.phase0_fx + (unsigned long long)cur_t * tile_dph);
            const long long m0 = (long long)cur_t * TILE_OUT + (g * R + rbeg);
            float2* o = p.out + (long long)cur_s * p.out_stride + m0;
#pragma unroll
            for (int r = 0; r < RO; ++r) y[r] = cmul(cmul(y[r], rot_thr[r]), rot_tile);
            if (writer) {
                if (m0 + RO <= p.n_out) {
                    if (p.vec_store && (RO % 2 == 0)) {
#pragma unroll
                        for (int r = 0; r < RO; r += 2)
                            __stcs(reinterpret_cast<float4*>(o + r), make_float4(y[r].x, y[r].y, y[r + 1].x, y[r + 1].y));
                    } else {
#pragma unroll
                        for (int r = 0; r < RO; ++r) __stcs(o + r, y[r]);
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < RO; ++r)
                        if (m0 + r < p.n_out) __stcs(o + r, y[r]);
                }
            }
            cur_s += gstep_s;
            cur_t += gstep_t;
            if (cur_t >= tps) { cur_t -= tps; ++cur_s; }
        }
    }
}

}  // namespace ddck
