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

// =============================================================================================================
// Fused persistent kernel
// =============================================================================================================
// Shared-memory layout of one pipeline stage: (NT + halo_rows) rows, one row per compute thread, each row holds the
// ROW = R*D float32 samples that produce that thread's R outputs, row pitch = ROW + 4 floats.  The 16-byte skew
// makes the per-thread LDS.128 of a quarter warp hit 8 distinct bank groups (lane stride 272 B), i.e. zero bank
// conflicts, while every row is still one 16-byte-aligned bulk copy.
template <int D, int R>
struct FusedCfg {
    static constexpr int ROW = R * D;       // samples per thread-row
    static constexpr int PITCH = ROW + 4;   // floats
    static constexpr int V = D / 4;         // float4 per tap block
    static_assert(D % 4 == 0, "D must be a multiple of 4");
    static_assert((ROW / 4) % 2 == 0, "row must be an even number of 16-byte chunks so that the skewed pitch is odd");
};

template <int D, int R, int NT, int STAGES, int MAXT, bool PACKED>
__global__ void __launch_bounds__(NT + 32, 1)
ddc_fused_kernel(const __grid_constant__ RunParams p, const __grid_constant__ TapsParam<MAXT> taps) {
    using C = FusedCfg<D, R>;
    constexpr int ROW = C::ROW, PITCH = C::PITCH, V = C::V;
    constexpr int TILE_OUT = NT * R;                // outputs per tile
    constexpr long long TILE_S = (long long)NT * ROW;  // samples per tile

    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw);
    uint64_t* empty_bar = full_bar + STAGES;
    float* buf = reinterpret_cast<float*>(smem_raw + 128);
    const int rows = NT + p.halo_rows;
    const int stage_floats = rows * PITCH;

    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], NT / 32);
        }
        mbar_fence_init();
    }
    __syncthreads();

    if (warp == NT / 32) {
        // ------------------------------------------------ producer warp: TMA bulk copies, one per row
        int it = 0;
        for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
            const int stage = it % STAGES;
            const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
            mbar_wait(&empty_bar[stage], ph ^ 1u);
            const long long s = tile / p.tiles_per_stream;
            const long long t = tile - s * p.tiles_per_stream;
            const float* src = reinterpret_cast<const float*>(p.in) + s * p.in_stride + t * TILE_S;
            float* dst = buf + (size_t)stage * stage_floats;
            if (lane == 0) mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)rows * ROW * 4u);
            __syncwarp();
            for (int r = lane; r < rows; r += 32) bulk_g2s(dst + r * PITCH, src + (size_t)r * ROW, ROW * 4u, &full_bar[stage]);
        }
    } else {
        // ------------------------------------------------ compute warps
        int it = 0;
        const int J = p.n_tap_blocks;
        for (long long tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
            const int stage = it % STAGES;
            const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
            mbar_wait(&full_bar[stage], ph);
            const float* base = buf + (size_t)stage * stage_floats + tid * PITCH;

            float2 acc[R];
#pragma unroll
            for (int r = 0; r < R; ++r) acc[r] = make_float2(0.f, 0.f);
            float4 xw[R][V];  // rotating window: slot (j + r) % R holds tap-block-sized sample block j + r
#pragma unroll
            for (int s = 0; s < R - 1; ++s)
#pragma unroll
                for (int v = 0; v < V; ++v) xw[s][v] = *reinterpret_cast<const float4*>(base + s * D + 4 * v);

            const float* rowp = base;
            for (int j0 = 0; j0 < J; j0 += R, rowp += PITCH) {
#pragma unroll
                for (int jj = 0; jj < R; ++jj) {
                    constexpr int dummy = 0; (void)dummy;
                    const int srel = jj + R - 1;  // newest block of this step, relative to block j0
#pragma unroll
                    for (int v = 0; v < V; ++v)
                        xw[srel % R][v] = *reinterpret_cast<const float4*>(rowp + (srel / R) * PITCH + (srel % R) * D + 4 * v);
                    const float4* tp = &taps.c2[(size_t)(j0 + jj) * (D / 2)];
#pragma unroll
                    for (int v = 0; v < V; ++v) {
                        const float4 ta = tp[2 * v], tb = tp[2 * v + 1];
#pragma unroll
                        for (int r = 0; r < R; ++r) acc[r] = ffma2(xw[(jj + r) % R][v].x, make_float2(ta.x, ta.y), acc[r]);
#pragma unroll
                        for (int r = 0; r < R; ++r) acc[r] = ffma2(xw[(jj + r) % R][v].y, make_float2(ta.z, ta.w), acc[r]);
#pragma unroll
                        for (int r = 0; r < R; ++r) acc[r] = ffma2(xw[(jj + r) % R][v].z, make_float2(tb.x, tb.y), acc[r]);
#pragma unroll
                        for (int r = 0; r < R; ++r) acc[r] = ffma2(xw[(jj + r) % R][v].w, make_float2(tb.z, tb.w), acc[r]);
                    }
                }
            }
            // all shared-memory reads of this stage are done -> hand the slot back to the producer
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[stage]);

            // epilogue: NCO rotation of each output by the phase of its first input sample, vectorised store
            const long long s = tile / p.tiles_per_stream;
            const long long t = tile - s * p.tiles_per_stream;
            const long long m0 = t * TILE_OUT + (long long)tid * R;
            float2* o = p.out + s * p.out_stride + m0;
            const unsigned long long ph0 = p.phase0_fx + (unsigned long long)(m0 * D) * p.step_fx;
            const unsigned long long dph = (unsigned long long)D * p.step_fx;
            float2 y[R];
#pragma unroll
            for (int r = 0; r < R; ++r) y[r] = cmul(acc[r], nco_rot(ph0 + (unsigned long long)r * dph));
            if (p.vec_store && (R % 2 == 0)) {
#pragma unroll
                for (int r = 0; r < R; r += 2)
                    __stcs(reinterpret_cast<float4*>(o + r), make_float4(y[r].x, y[r].y, y[r + 1].x, y[r + 1].y));
            } else {
#pragma unroll
                for (int r = 0; r < R; ++r) __stcs(o + r, y[r]);
            }
        }
    }
}

}  // namespace ddck
