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

}  // namespace ddck
