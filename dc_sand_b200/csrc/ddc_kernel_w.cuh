// Fast-FIR fused DDC kernel ("kernel W"): the phase-major kernel of ddc_kernel_p.cuh with 25 % fewer multiplies.
//
// At T = 256, D = 16 the direct form needs 2T/D = 32 FP32 FMAs per input sample, which puts the path on the FP32
// roofline (231 us at 74.4 TFLOP/s for 2^28 samples) instead of the HBM one (185 us).  The polyphase branch of phase d
// is a J-tap FIR over the block index b (x_d[b] = x[bD + d]):   y[m] = sum_d sum_j c[jD + d] x_d[m + j].
// Splitting outputs and taps into even/odd (the 2-by-2 fast FIR algorithm, a.k.a. Winograd F(2,2)):
//
//      y[2r]   = M0[r] + M1[r]            M0[r] = sum_d sum_i (x_d[2(r+i)]   - x_d[2(r+i)+1]) * c[2i D + d]
//      y[2r+1] = M1[r] - M2[r]            M1[r] = sum_d sum_i  x_d[2(r+i)+1]                  * (c[2i D + d] + c[(2i+1) D + d])
//                                         M2[r] = sum_d sum_i (x_d[2(r+i)+1] - x_d[2(r+i)+2]) * c[(2i+1) D + d]
//
// i.e. three half-length FIRs at half the output rate: 3/4 of the FMAs, plus one subtraction per window sample that a
// thread forms in registers (2 (R/2 + J/2 - 1) per phase against 3 (R/2)(J/2) FFMA2).  The sums M0..M2 are linear, so
// they accumulate over all phases and tap pairs and are combined once per chunk.  For integer-valued digitiser samples
// the differences are exact; in general the rounding error is the same order as the direct form (tools/winograd_error.py:
// 4.5e-7 against 7.9e-7 of max|y| at T = 256, D = 16, because the accumulation chains are half as long).
//
// Everything else -- chunk ring, two elect-based TMA producer warps, slot sequence guard, taps as kernel parameters read
// through uniform registers, FFMA2 with a broadcast sample operand, deferred epilogue -- is kernel P's (ddc_kernel_p.cuh).
#pragma once
#include "ddc_kernel_p.cuh"

namespace ddck {

// taps for kernel W: index ((3 i + seq) D + d), seq 0: c[2i D + d], 1: c[2i D + d] + c[(2i+1) D + d], 2: c[(2i+1) D + d]
template <int D, int JT>
struct WCfg : PCfg<D, JT, 1> {
    using B = PCfg<D, JT, 1>;
    static_assert(JT % 2 == 0 && B::R % 2 == 0, "fast FIR needs an even number of tap blocks and outputs per thread");
    // Long filters (JT = 32, 64, ...) run as NJG passes of JP = 16 tap blocks over the same staged chunk: pass jg reads the
    // window that starts 16 jg blocks (= 16 jg / R thread-rows) further on; the ring geometry (halo rows, slot size) is
    // PCfg's for the whole filter, the register window is the 16-block one.
    static constexpr int JP = JT < 16 ? JT : 16;   // tap blocks per pass
    static constexpr int NJG = JT / JP;            // passes
    static_assert(JT % JP == 0, "long filters are padded to a multiple of 16 tap blocks");
    static_assert(NJG == 1 || (JP % B::R == 0), "a pass must advance the window by whole thread-rows");
    static constexpr int JH = JP / 2;              // tap pairs per pass
    static constexpr int RH = B::R / 2;            // output pairs per thread
    static constexpr int NS = RH + JH - 1;         // entries of each derived sequence a thread touches per pass
    static constexpr int NWP = B::R + JP - 1;      // blocks of the register window of one pass
    static constexpr int WROWS = (NWP - 1) / B::R + 1;   // thread-rows a pass window spans
    static constexpr int NTW = 3 * (JT / 2) * D;   // complex taps passed to the kernel (one level)
    static constexpr int NTW2 = 9 * ((JT + 3) / 4) * D;   // two nested levels
};

#ifndef DDCB200_W10_UNPACK_UNROLL
#define DDCB200_W10_UNPACK_UNROLL 1   // 3 was measured slower (1.174 against 1.130 ms on 64 x 2^24 samples)
#endif
#ifndef DDCB200_W_PACKED_SUB
#define DDCB200_W_PACKED_SUB 1
#endif

__device__ __forceinline__ float4 sub4(const float4 a, const float4 b) {
#if DDCB200_W_PACKED_SUB
    // two packed subtractions (SASS: FADD2 with a negated operand) instead of four FADD; measured 7 % faster end to end
    const float2 lo = __fadd2_rn(make_float2(a.x, a.y), make_float2(-b.x, -b.y));
    const float2 hi = __fadd2_rn(make_float2(a.z, a.w), make_float2(-b.z, -b.w));
    return make_float4(lo.x, lo.y, hi.x, hi.y);
#else
    return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
#endif
}

// One phase group (4 phases) of the three half-rate FIRs: 3 * RH * JH * 4 FFMA2 on the window w[0 .. R+JT-2] (float4 =
// the 4 phases of one D-sample block); tp points at the float4 pair of this phase group inside tap set (i = 0, seq = 0).
template <int D, int JT, int R>
__device__ __forceinline__ void w_fir_pg(const float4 (&w)[R + JT - 1], const float4* tp, float2 (&m0a)[R / 2],
                                         float2 (&m1a)[R / 2], float2 (&m2a)[R / 2]) {
    constexpr int JH = JT / 2, RH = R / 2;
#pragma unroll
    for (int i = 0; i < JH; ++i) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            {
                const float4 t = tp[(3 * i + 0) * (D / 2) + half];
#pragma unroll
                for (int r = 0; r < RH; ++r) {
                    const float4 s = sub4(w[2 * (r + i)], w[2 * (r + i) + 1]);
                    m0a[r] = ffma2(half ? s.z : s.x, make_float2(t.x, t.y), m0a[r]);
                }
#pragma unroll
                for (int r = 0; r < RH; ++r) {
                    const float4 s = sub4(w[2 * (r + i)], w[2 * (r + i) + 1]);
                    m0a[r] = ffma2(half ? s.w : s.y, make_float2(t.z, t.w), m0a[r]);
                }
            }
            {
                const float4 t = tp[(3 * i + 1) * (D / 2) + half];
#pragma unroll
                for (int r = 0; r < RH; ++r) {
                    const float4 s = w[2 * (r + i) + 1];
                    m1a[r] = ffma2(half ? s.z : s.x, make_float2(t.x, t.y), m1a[r]);
                }
#pragma unroll
                for (int r = 0; r < RH; ++r) {
                    const float4 s = w[2 * (r + i) + 1];
                    m1a[r] = ffma2(half ? s.w : s.y, make_float2(t.z, t.w), m1a[r]);
                }
            }
            {
                const float4 t = tp[(3 * i + 2) * (D / 2) + half];
#pragma unroll
                for (int r = 0; r < RH; ++r) {
                    const float4 s = sub4(w[2 * (r + i) + 1], w[2 * (r + i) + 2]);
                    m2a[r] = ffma2(half ? s.z : s.x, make_float2(t.x, t.y), m2a[r]);
                }
#pragma unroll
                for (int r = 0; r < RH; ++r) {
                    const float4 s = sub4(w[2 * (r + i) + 1], w[2 * (r + i) + 2]);
                    m2a[r] = ffma2(half ? s.w : s.y, make_float2(t.z, t.w), m2a[r]);
                }
            }
        }
    }
}

// Two nested levels of the 2-by-2 fast FIR (9/16 of the multiplies): each of the three half-rate FIRs of w_fir_pg
// (sequence s_a, taps g_a) is split once more into three quarter-rate FIRs
//      A_ab[rho] = sum_iota t_ab[rho + iota] * h_ab[iota],       rho = 0 .. R/4-1,  iota = 0 .. JP/4-1,
//      t_a0[v] = s_a[2v] - s_a[2v+1],  t_a1[v] = s_a[2v+1],  t_a2[v] = s_a[2v+1] - s_a[2v+2]          (data, in registers)
//      h_a0[i] = g_a[2i],              h_a1[i] = g_a[2i] + g_a[2i+1],   h_a2[i] = g_a[2i+1]            (taps, from the host)
// and M_a[2 rho] = A_a0 + A_a1, M_a[2 rho + 1] = A_a1 - A_a2 once per chunk.  Per phase group: 9 (R/4)(JP/4) 4 = 288 FFMA2
// and 104 FADD2 at R = 8, JP = 16, against 384 + 44 for one level and 512 for the direct form.  Rounding error of the
// nested form on the reference filter: 3.0e-7 of max|y| (tools/winograd_error.py), below the one-level and direct forms.
// Tap index: ((9 iota + 3 a + b) D + d).
// MEASURED: slower than one level at R = 8 (0.252 against 0.229 ms compute-only, T = 256, 2^28 samples): every tap fetch
// (LDCU.64) feeds only R / 4 = 2 FFMA2 and the uniform datapath becomes the limit.  It would need R = 16 outputs per thread,
// i.e. 32 KB chunks that do not fit an 8-warp ring.  Kept behind option "variant" = 9 for that experiment only.
template <int D, int JP, int R>
__device__ __forceinline__ void w2_fir_pg(const float4 (&w)[R + JP - 1], const float4* tp, float2 (&acc)[9][R / 4]) {
    static_assert(JP % 4 == 0 && R % 4 == 0, "nested fast FIR needs multiples of four");
    constexpr int JQ = JP / 4, RQ = R / 4;
    constexpr int NS1 = R / 2 + JP / 2 - 1;     // level-1 sequence length
    constexpr int NS2 = RQ + JQ - 1;            // level-2 sequence length
    float4 t[9][NS2];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float4 s[NS1];
#pragma unroll
        for (int n = 0; n < NS1; ++n)
            s[n] = a == 0 ? sub4(w[2 * n], w[2 * n + 1]) : (a == 1 ? w[2 * n + 1] : sub4(w[2 * n + 1], w[2 * n + 2]));
#pragma unroll
        for (int v = 0; v < NS2; ++v) {
            t[3 * a + 0][v] = sub4(s[2 * v], s[2 * v + 1]);
            t[3 * a + 1][v] = s[2 * v + 1];
            t[3 * a + 2][v] = sub4(s[2 * v + 1], s[2 * v + 2]);
        }
    }
#pragma unroll
    for (int i = 0; i < JQ; ++i) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float4 tt[9];
#pragma unroll
            for (int ab = 0; ab < 9; ++ab) tt[ab] = tp[(9 * i + ab) * (D / 2) + half];
#pragma unroll
            for (int ab = 0; ab < 9; ++ab)
#pragma unroll
                for (int r = 0; r < RQ; ++r)
                    acc[ab][r] = ffma2(half ? t[ab][r + i].z : t[ab][r + i].x, make_float2(tt[ab].x, tt[ab].y), acc[ab][r]);
#pragma unroll
            for (int ab = 0; ab < 9; ++ab)
#pragma unroll
                for (int r = 0; r < RQ; ++r)
                    acc[ab][r] = ffma2(half ? t[ab][r + i].w : t[ab][r + i].y, make_float2(tt[ab].z, tt[ab].w), acc[ab][r]);
        }
    }
}

// y[0 .. R-1] from the nine quarter-rate sums
template <int R>
__device__ __forceinline__ void w2_combine(const float2 (&acc)[9][R / 4], float2 (&y)[R]) {
    constexpr int RQ = R / 4;
    float2 m[3][2 * RQ];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int r = 0; r < RQ; ++r) {
            m[a][2 * r] = make_float2(acc[3 * a][r].x + acc[3 * a + 1][r].x, acc[3 * a][r].y + acc[3 * a + 1][r].y);
            m[a][2 * r + 1] = make_float2(acc[3 * a + 1][r].x - acc[3 * a + 2][r].x, acc[3 * a + 1][r].y - acc[3 * a + 2][r].y);
        }
#pragma unroll
    for (int r = 0; r < 2 * RQ; ++r) {
        y[2 * r] = make_float2(m[0][r].x + m[1][r].x, m[0][r].y + m[1][r].y);
        y[2 * r + 1] = make_float2(m[1][r].x - m[2][r].x, m[1][r].y - m[2][r].y);
    }
}

// Branch-free epilogue (so that it can live inside the FIR's basic block, deferred by one chunk): polynomial NCO and
// predicated stores.  A thread's R outputs start at a multiple of R elements of its output row; the row itself may start
// on an odd complex64 element ([streams, M] arrays with odd M), so the 16-byte pairing is chosen per thread from the
// address: (0,1)(2,3).. when it is 16-byte aligned, 0 | (1,2)(3,4).. | R-1 otherwise.  `left` = how many of my outputs
// exist (ragged stream tail); nout <= m0 disables every store.
template <int R>
__device__ __forceinline__ void w_epilogue(const float2 (&y)[R], const float2 (&rot_thr)[R], unsigned long long chunk_phase,
                                           long long m0, float2* o, long long nout) {
    const float2 rot_chunk = nco_rot_bf(chunk_phase);
    const int left = (int)(nout - m0 < (long long)R ? (nout - m0 < 0 ? 0 : nout - m0) : (long long)R);
    const bool odd = (reinterpret_cast<unsigned long long>(o) & 8ull) != 0;
    float2 z[R];
#pragma unroll
    for (int r = 0; r < R; ++r) z[r] = cmul(cmul(y[r], rot_thr[r]), rot_chunk);
#pragma unroll
    for (int r = 0; r < R; r += 2) {
        st_cs_v4_if(o + r, z[r].x, z[r].y, z[r + 1].x, z[r + 1].y, !odd && left >= r + 2);
        st_cs_v2_if(o + r, z[r].x, z[r].y, (!odd && left == r + 1) || (odd && r == 0 && left >= 1));
    }
#pragma unroll
    for (int r = 1; r < R; r += 2) {
        if (r + 1 < R) st_cs_v4_if(o + r, z[r].x, z[r].y, z[r + 1].x, z[r + 1].y, odd && left >= r + 2);
        st_cs_v2_if(o + r, z[r].x, z[r].y, odd && (left == r + 1 || (r == R - 1 && left >= R)));
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// float32 input
// ---------------------------------------------------------------------------------------------------------------------
// NEST = 0: one fast-FIR level (w_fir_pg, 3 (JT/2) D complex taps); NEST = 1: two nested levels (w2_fir_pg, 9 (JT/4) D taps)
template <int D, int JT, int NEST>
__global__ void __launch_bounds__(WCfg<D, JT>::NWARPS * 32 + 32 * WCfg<D, JT>::NPROD, 1)
ddc_fused_w_kernel(const __grid_constant__ RunParams p,
                   const __grid_constant__ TapsParam<(NEST ? WCfg<D, JT>::NTW2 : WCfg<D, JT>::NTW)> taps) {
    using C = WCfg<D, JT>;
    constexpr int ROW = C::ROW, R = C::R, NW = C::NWP, NWARPS = C::NWARPS, NG = C::NGROUPS;
    constexpr int NSLOT = C::NSLOT, RH = C::RH, NS = C::NS, JP = C::JP, NJG = C::NJG;
    constexpr int WANT = C::TOT_ROWS * ROW;
    static_assert(2 * (NS - 1) + 2 == NW - 1, "window bookkeeping");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw);
    uint64_t* empty_bar = full_bar + 16;
    volatile int* slot_seq = reinterpret_cast<volatile int*>(smem_raw + 384);
    float* buf = reinterpret_cast<float*>(smem_raw + C::HDR_BYTES);

    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int lane = tid & 31;
    if (tid == 0) {
#pragma unroll 1
        for (int s = 0; s < NSLOT; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
            slot_seq[s] = -1;
        }
        mbar_fence_init();
    }
    __syncthreads();

    const int cps = (int)p.tiles_per_stream;
    const int n_k = (int)((p.total_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);
    const unsigned long long chunk_dph = (unsigned long long)((long long)C::CHUNK_OUT * D) * p.step_fx;

    if (warp >= NWARPS) {
        // ------------------------------------------------------------------ producer warps (as in ddc_fused_pd_kernel)
        constexpr int NP = C::NPROD;
        const int pid = warp - NWARPS;
        const long long pstride = (long long)NP * gridDim.x;
        const int gs = (int)(pstride / cps), gc = (int)(pstride % cps);
        const long long pfirst = blockIdx.x + (long long)pid * gridDim.x;
        int cs = (int)(pfirst / cps), cc = (int)(pfirst % cps);
        const int sbase = C::sub_base(pid), scnt = C::sub_count(pid);
        int sidx = 0;
        uint32_t par = 1;
        for (int k = pid; k < n_k && (p.debug_mode & 255) != 1; k += NP) {
            const int slot = sbase + sidx;
            const bool leader = elect_one();
            if (leader) {
                mbar_wait(&empty_bar[slot], par);
                slot_seq[slot] = k;
            }
            __syncwarp();
            const float* src = reinterpret_cast<const float*>(p.in) + (long long)cs * p.in_stride + (long long)cc * C::CHUNK_S;
            float* dst = buf + (size_t)slot * C::SLOT_FLOATS;
            const long long valid = p.n_samples - (long long)cc * C::CHUNK_S;
            if (valid >= WANT) {
                if (leader) {
                    mbar_arrive_expect_tx(&full_bar[slot], (uint32_t)WANT * 4u);
#pragma unroll
                    for (int sr = 0; sr < C::NSR; ++sr) {
                        constexpr int SR4 = C::SROWS;
                        const int nrow = (C::TOT_ROWS - sr * SR4) < SR4 ? (C::TOT_ROWS - sr * SR4) : SR4;
                        bulk_g2s(dst + sr * C::SRP, src + sr * SR4 * ROW, (uint32_t)nrow * ROW * 4u, &full_bar[slot]);
                    }
                }
            } else {
                // ragged last chunk of a stream: whole 16-byte groups by TMA, the last 1-3 samples by hand, zeros after
                uint32_t tx = 0;
                for (int sr = 0; sr < C::NSR; ++sr) {
                    const int cap = ((C::TOT_ROWS - sr * C::SROWS) < C::SROWS ? (C::TOT_ROWS - sr * C::SROWS) : C::SROWS) * ROW;
                    const long long s0 = (long long)sr * C::SROWS * ROW;
                    long long cnt = valid - s0;
                    cnt = cnt < 0 ? 0 : (cnt > cap ? cap : cnt);
                    const int bulk = (int)cnt & ~3;
                    for (int e = bulk + lane; e < cap; e += 32) dst[sr * C::SRP + e] = (e < (int)cnt) ? src[s0 + e] : 0.f;
                    tx += (uint32_t)bulk * 4u;
                }
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive_expect_tx(&full_bar[slot], tx);
                    for (int sr = 0; sr < C::NSR; ++sr) {
                        const int cap = ((C::TOT_ROWS - sr * C::SROWS) < C::SROWS ? (C::TOT_ROWS - sr * C::SROWS) : C::SROWS) * ROW;
                        const long long s0 = (long long)sr * C::SROWS * ROW;
                        long long cnt = valid - s0;
                        cnt = cnt < 0 ? 0 : (cnt > cap ? cap : cnt);
                        const int bulk = (int)cnt & ~3;
                        if (bulk > 0) bulk_g2s(dst + sr * C::SRP, src + s0, (uint32_t)bulk * 4u, &full_bar[slot]);
                    }
                }
            }
            __syncwarp();
            if (++sidx == scnt) { sidx = 0; par ^= 1u; }
            cs += gs;
            cc += gc;
            if (cc >= cps) { cc -= cps; ++cs; }
        }
    } else {
        // ------------------------------------------------------------------ compute warps
        const int grp = warp;
        const int g = (lane & 7) * C::SROWS + (lane >> 3);
        int rowoff[C::HALO_ROWS + 1];
#pragma unroll
        for (int h = 0; h <= C::HALO_ROWS; ++h) rowoff[h] = C::row_offset(g + h);
        float2 rot_thr[R];
#pragma unroll
        for (int r = 0; r < R; ++r) rot_thr[r] = nco_rot((unsigned long long)((long long)(g * R + r) * D) * p.step_fx);

        const long long kstride = (long long)NG * gridDim.x;
        const int gs = (int)(kstride / cps), gc = (int)(kstride % cps);
        const long long first = blockIdx.x + (long long)grp * gridDim.x;
        int cs = (int)(first / cps), cc = (int)(first % cps);
        const int sbase = C::sub_base(grp % C::NPROD), scnt = C::sub_count(grp % C::NPROD);
        int sidx = (grp / C::NPROD) % scnt;
        uint32_t par = (uint32_t)((grp / C::NPROD) / scnt) & 1u;

        // deferred epilogue state (see ddc_fused_pd_kernel): the previous chunk's sums and where they go
        float2 yprev[R];
#pragma unroll
        for (int r = 0; r < R; ++r) yprev[r] = make_float2(0.f, 0.f);
        long long prev_m0 = 0;
        float2* prev_o = p.out;
        int prev_cc = 0;
        long long prev_nout = 0;   // 0 disables the stores

        float2 m0a[RH], m1a[RH], m2a[RH];   // NEST = 0
        constexpr int RQN = NEST ? R / 4 : 1;
        float2 acc9[9][RQN];                // NEST = 1
        constexpr int TPG = NEST ? 9 * (JP / 4) * (D / 2) : 3 * (JP / 2) * (D / 2);   // float4 per tap group of JP blocks
        auto fir = [&](const float4(&w)[NW], const float4* tpp) {
            if constexpr (NEST) w2_fir_pg<D, JP, R>(w, tpp, acc9);
            else w_fir_pg<D, JP, R>(w, tpp, m0a, m1a, m2a);
        };
        long long t_wait = 0;
        const long long t_begin = clock64();
        const bool memonly = (p.debug_mode & 255) == 2;   // tuning aid: ring traffic without the FIR
        for (int k = grp; k < n_k; k += NG) {
            const int slot = sbase + sidx;
            if ((p.debug_mode & 255) != 1) {
                const long long tw0 = p.dbg ? clock64() : 0;
                while (slot_seq[slot] != k) {}
                mbar_wait(&full_bar[slot], par);
                if (p.dbg) t_wait += clock64() - tw0;
            }
            const float* sbuf = buf + (size_t)slot * C::SLOT_FLOATS;
#pragma unroll
            for (int r = 0; r < RH; ++r) m0a[r] = m1a[r] = m2a[r] = make_float2(0.f, 0.f);
#pragma unroll
            for (int ab = 0; ab < 9; ++ab)
#pragma unroll
                for (int r = 0; r < RQN; ++r) acc9[ab][r] = make_float2(0.f, 0.f);

            if (memonly) {
                if (sbuf[rowoff[0]] == 123.456f) m0a[0].x = 1.f;
            } else {
                // Fully unrolling the phase groups (to overlap the next group's LDS and release the slot a quarter chunk
                // earlier) was measured SLOWER (0.256 against 0.242 ms), so the loop stays rolled: group 0 carries the
                // previous chunk's epilogue in its basic block, groups 1 .. V-1 share one loop body.
                int xoff = 0;
                const float4* tp = &taps.c2[0];
                {
                    float4 w[NW];
#pragma unroll
                    for (int b = 0; b < NW; ++b)
                        w[b] = *reinterpret_cast<const float4*>(sbuf + rowoff[b / R] + (b % R) * D);
                    w_epilogue<R>(yprev, rot_thr, p.phase0_fx + (unsigned long long)prev_cc * chunk_dph, prev_m0, prev_o, prev_nout);
                    fir(w, tp);
                    xoff = 4;
                    tp += 2;
                }
                if constexpr (NJG == 1) {
#pragma unroll 1
                    for (int pg = 1; pg < C::V; ++pg, tp += 2) {
                        asm volatile("" : "+r"(xoff));
                        float4 w[NW];
#pragma unroll
                        for (int b = 0; b < NW; ++b)
                            w[b] = *reinterpret_cast<const float4*>(sbuf + xoff + rowoff[b / R] + (b % R) * D);
                        xoff += 4;
                        fir(w, tp);
                    }
                } else {
                    // passes 1 .. NJG V - 1 share one loop body: pass = jg V + pg; the per-thread row offsets of tap group jg
                    // are recomputed per pass (a dozen integer instructions against 384 FFMA2)
                    int grow = g;   // first thread-row of the current pass window (opaque to the induction-variable optimiser)
                    int pgi = 1;
#pragma unroll 1
                    for (int pass = 1; pass < NJG * C::V; ++pass) {
                        asm volatile("" : "+r"(xoff), "+r"(grow));
                        int ro[C::WROWS];
#pragma unroll
                        for (int h = 0; h < C::WROWS; ++h) ro[h] = ((grow + h) / C::SROWS) * C::SRP + ((grow + h) % C::SROWS) * ROW;
                        float4 w[NW];
#pragma unroll
                        for (int b = 0; b < NW; ++b)
                            w[b] = *reinterpret_cast<const float4*>(sbuf + xoff + ro[b / R] + (b % R) * D);
                        fir(w, tp);
                        // next pass: next phase group, or phase group 0 of the next tap group
                        ++pgi;
                        xoff += 4;
                        tp += 2;
                        if (pgi == C::V) {
                            pgi = 0;
                            xoff = 0;
                            grow += JP / R;
                            tp += TPG - 2 * C::V;   // first tap set of the next group, phase group 0
                        }
                    }
                }
            }
            __syncwarp();
            if (lane == 0 && (p.debug_mode & 255) != 1) mbar_arrive(&empty_bar[slot]);

            // combine the three half-rate sums and hand them to the next iteration's deferred epilogue
            if constexpr (NEST) {
                w2_combine<R>(acc9, yprev);
            } else {
#pragma unroll
                for (int r = 0; r < RH; ++r) {
                    yprev[2 * r] = make_float2(m0a[r].x + m1a[r].x, m0a[r].y + m1a[r].y);
                    yprev[2 * r + 1] = make_float2(m1a[r].x - m2a[r].x, m1a[r].y - m2a[r].y);
                }
            }
            prev_cc = cc;
            prev_m0 = (long long)cc * C::CHUNK_OUT + g * R;
            prev_o = p.out + (long long)cs * p.out_stride + prev_m0;
            prev_nout = p.n_out;

            sidx += NG / C::NPROD;
            if (sidx >= scnt) { sidx -= scnt; par ^= 1u; }
            cs += gs;
            cc += gc;
            if (cc >= cps) { cc -= cps; ++cs; }
        }
        w_epilogue<R>(yprev, rot_thr, p.phase0_fx + (unsigned long long)prev_cc * chunk_dph, prev_m0, prev_o, prev_nout);
        if (p.dbg && lane == 0) {
            atomicAdd(p.dbg, (unsigned long long)t_wait);
            atomicAdd(p.dbg + 1, (unsigned long long)(clock64() - t_begin));
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Sixteen compute warps: two warps per chunk, each taking half of the phase groups ("kernel W2X").
//
// With one warp per chunk only 8 compute warps fit the 13-slot ring, i.e. two per scheduler, and the FMA pipe idles ~11 % of
// the time (pass boundaries, dependency bubbles) however the loop is arranged: stagger, prefetch and tap-fetch experiments
// all left the compute-only time at 0.229 ms.  Here warps w and w + 8 share chunk slot and work: w runs phase groups 0, 1 of
// the chunk, w + 8 phase groups 2, 3, both over all eight outputs of every thread-row.  The partial sums meet through the
// (by then dead) head of the chunk slot: each warp finishes four of the eight outputs (and their deferred epilogue).
// Four warps per scheduler at 110 registers is exactly what the register file holds (576 threads x 112).
// MEASURED: 0.248 ms against 0.244 ms for one warp per chunk -- the two named-barrier rendezvous and the doubled per-chunk
// bookkeeping eat what the extra warps gain.  Kept behind option "variant" = 12 as a documented experiment.
// ---------------------------------------------------------------------------------------------------------------------
template <int D, int JT>
__global__ void __launch_bounds__(2 * WCfg<D, JT>::NWARPS * 32 + 32 * WCfg<D, JT>::NPROD, 1)
ddc_fused_w2x_kernel(const __grid_constant__ RunParams p, const __grid_constant__ TapsParam<WCfg<D, JT>::NTW> taps) {
    using C = WCfg<D, JT>;
    constexpr int ROW = C::ROW, R = C::R, NW = C::NWP, NWARPS = C::NWARPS, NG = C::NGROUPS;
    constexpr int NSLOT = C::NSLOT, RH = C::RH, JP = C::JP;
    constexpr int WANT = C::TOT_ROWS * ROW;
    constexpr int RO = R / 2;          // outputs a warp finishes
    static_assert(C::NJG == 1 && C::V == 4 && R == 8, "two-warps-per-chunk variant: D = 16, at most 16 tap blocks");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw);
    uint64_t* empty_bar = full_bar + 16;
    volatile int* slot_seq = reinterpret_cast<volatile int*>(smem_raw + 384);
    float* buf = reinterpret_cast<float*>(smem_raw + C::HDR_BYTES);

    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int lane = tid & 31;
    if (tid == 0) {
#pragma unroll 1
        for (int s = 0; s < NSLOT; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 2);      // both warps of a pair hand the slot back
            slot_seq[s] = -1;
        }
        mbar_fence_init();
    }
    __syncthreads();

    const int cps = (int)p.tiles_per_stream;
    const int n_k = (int)((p.total_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);
    const unsigned long long chunk_dph = (unsigned long long)((long long)C::CHUNK_OUT * D) * p.step_fx;

    if (warp >= 2 * NWARPS) {
        // ------------------------------------------------------------------ producer warps (as in ddc_fused_w_kernel)
        constexpr int NP = C::NPROD;
        const int pid = warp - 2 * NWARPS;
        const long long pstride = (long long)NP * gridDim.x;
        const int gs = (int)(pstride / cps), gc = (int)(pstride % cps);
        const long long pfirst = blockIdx.x + (long long)pid * gridDim.x;
        int cs = (int)(pfirst / cps), cc = (int)(pfirst % cps);
        const int sbase = C::sub_base(pid), scnt = C::sub_count(pid);
        int sidx = 0;
        uint32_t par = 1;
        for (int k = pid; k < n_k; k += NP) {
            const int slot = sbase + sidx;
            const bool leader = elect_one();
            if (leader) {
                mbar_wait(&empty_bar[slot], par);
                slot_seq[slot] = k;
            }
            __syncwarp();
            const float* src = reinterpret_cast<const float*>(p.in) + (long long)cs * p.in_stride + (long long)cc * C::CHUNK_S;
            float* dst = buf + (size_t)slot * C::SLOT_FLOATS;
            const long long valid = p.n_samples - (long long)cc * C::CHUNK_S;
            if (valid >= WANT) {
                if (leader) {
                    mbar_arrive_expect_tx(&full_bar[slot], (uint32_t)WANT * 4u);
#pragma unroll
                    for (int sr = 0; sr < C::NSR; ++sr) {
                        constexpr int SR4 = C::SROWS;
                        const int nrow = (C::TOT_ROWS - sr * SR4) < SR4 ? (C::TOT_ROWS - sr * SR4) : SR4;
                        bulk_g2s(dst + sr * C::SRP, src + sr * SR4 * ROW, (uint32_t)nrow * ROW * 4u, &full_bar[slot]);
                    }
                }
            } else {
                uint32_t tx = 0;
                for (int sr = 0; sr < C::NSR; ++sr) {
                    const int cap = ((C::TOT_ROWS - sr * C::SROWS) < C::SROWS ? (C::TOT_ROWS - sr * C::SROWS) : C::SROWS) * ROW;
                    const long long s0 = (long long)sr * C::SROWS * ROW;
                    long long cnt = valid - s0;
                    cnt = cnt < 0 ? 0 : (cnt > cap ? cap : cnt);
                    const int bulk = (int)cnt & ~3;
                    for (int e = bulk + lane; e < cap; e += 32) dst[sr * C::SRP + e] = (e < (int)cnt) ? src[s0 + e] : 0.f;
                    tx += (uint32_t)bulk * 4u;
                }
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive_expect_tx(&full_bar[slot], tx);
                    for (int sr = 0; sr < C::NSR; ++sr) {
                        const int cap = ((C::TOT_ROWS - sr * C::SROWS) < C::SROWS ? (C::TOT_ROWS - sr * C::SROWS) : C::SROWS) * ROW;
                        const long long s0 = (long long)sr * C::SROWS * ROW;
                        long long cnt = valid - s0;
                        cnt = cnt < 0 ? 0 : (cnt > cap ? cap : cnt);
                        const int bulk = (int)cnt & ~3;
                        if (bulk > 0) bulk_g2s(dst + sr * C::SRP, src + s0, (uint32_t)bulk * 4u, &full_bar[slot]);
                    }
                }
            }
            __syncwarp();
            if (++sidx == scnt) { sidx = 0; par ^= 1u; }
            cs += gs;
            cc += gc;
            if (cc >= cps) { cc -= cps; ++cs; }
        }
    } else {
        // ------------------------------------------------------------------ compute warps: pair grp = warp % 8, half hw = warp / 8
        const int grp = warp % NG;
        const int hw = warp / NG;
        const int g = (lane & 7) * C::SROWS + (lane >> 3);
        int rowoff[C::HALO_ROWS + 1];
#pragma unroll
        for (int h = 0; h <= C::HALO_ROWS; ++h) rowoff[h] = C::row_offset(g + h);
        float2 rot_thr[RO];   // my outputs of the thread-row: r = RO hw .. RO hw + RO - 1
#pragma unroll
        for (int r = 0; r < RO; ++r) rot_thr[r] = nco_rot((unsigned long long)((long long)(g * R + RO * hw + r) * D) * p.step_fx);

        const long long kstride = (long long)NG * gridDim.x;
        const int gs = (int)(kstride / cps), gc = (int)(kstride % cps);
        const long long first = blockIdx.x + (long long)grp * gridDim.x;
        int cs = (int)(first / cps), cc = (int)(first % cps);
        const int sbase = C::sub_base(grp % C::NPROD), scnt = C::sub_count(grp % C::NPROD);
        int sidx = (grp / C::NPROD) % scnt;
        uint32_t par = (uint32_t)((grp / C::NPROD) / scnt) & 1u;
        const int bar_id = 1 + grp;   // named barrier of the pair (64 threads)

        float2 yprev[RO];
#pragma unroll
        for (int r = 0; r < RO; ++r) yprev[r] = make_float2(0.f, 0.f);
        long long prev_m0 = 0;
        float2* prev_o = p.out;
        int prev_cc = 0;
        long long prev_nout = 0;
        float2 m0a[RH], m1a[RH], m2a[RH];

        for (int k = grp; k < n_k; k += NG) {
            const int slot = sbase + sidx;
            while (slot_seq[slot] != k) {}
            mbar_wait(&full_bar[slot], par);
            float* sbuf = buf + (size_t)slot * C::SLOT_FLOATS;
#pragma unroll
            for (int r = 0; r < RH; ++r) m0a[r] = m1a[r] = m2a[r] = make_float2(0.f, 0.f);
            // my two phase groups: 2 hw and 2 hw + 1
            const float4* tp = &taps.c2[4 * hw];
            int xoff = 8 * hw;
            asm volatile("" : "+r"(xoff));
            {
                float4 w[NW];
#pragma unroll
                for (int b = 0; b < NW; ++b)
                    w[b] = *reinterpret_cast<const float4*>(sbuf + xoff + rowoff[b / R] + (b % R) * D);
                w_epilogue<RO>(yprev, rot_thr, p.phase0_fx + (unsigned long long)prev_cc * chunk_dph, prev_m0, prev_o, prev_nout);
                w_fir_pg<D, JP, R>(w, tp, m0a, m1a, m2a);
            }
            {
                float4 w[NW];
#pragma unroll
                for (int b = 0; b < NW; ++b)
                    w[b] = *reinterpret_cast<const float4*>(sbuf + xoff + 4 + rowoff[b / R] + (b % R) * D);
                w_fir_pg<D, JP, R>(w, tp + 2, m0a, m1a, m2a);
            }
            // both warps are done reading the chunk: its head becomes the exchange area (48 bytes per lane and direction)
            asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
            float4* xs = reinterpret_cast<float4*>(sbuf) + (hw * 32 + lane) * 3;              // what I send
            const float4* xr = reinterpret_cast<const float4*>(sbuf) + ((hw ^ 1) * 32 + lane) * 3;   // what my partner sent
            // the pairs my partner finishes: 2, 3 if I am half 0, else 0, 1 (selects, not dynamic indexing: the sums stay in registers)
            const bool h1 = hw != 0;
            auto sel = [&](const float2& lo, const float2& hi) { return h1 ? lo : hi; };
            {
                const float2 s0 = sel(m0a[0], m0a[2]), s1 = sel(m0a[1], m0a[3]);
                const float2 t0 = sel(m1a[0], m1a[2]), t1 = sel(m1a[1], m1a[3]);
                const float2 u0 = sel(m2a[0], m2a[2]), u1 = sel(m2a[1], m2a[3]);
                xs[0] = make_float4(s0.x, s0.y, s1.x, s1.y);
                xs[1] = make_float4(t0.x, t0.y, t1.x, t1.y);
                xs[2] = make_float4(u0.x, u0.y, u1.x, u1.y);
            }
            asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory");
            const float4 r0 = xr[0], r1 = xr[1], r2 = xr[2];
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[slot]);   // the second arrival returns the slot to the producer

            // my output pairs: 2 hw, 2 hw + 1 -> outputs RO hw .. RO hw + 3
            auto mine = [&](const float2& lo, const float2& hi) { return h1 ? hi : lo; };
            const float2 k0 = mine(m0a[0], m0a[2]), k1 = mine(m0a[1], m0a[3]);
            const float2 l0 = mine(m1a[0], m1a[2]), l1 = mine(m1a[1], m1a[3]);
            const float2 n0 = mine(m2a[0], m2a[2]), n1 = mine(m2a[1], m2a[3]);
            const float2 a0 = make_float2(k0.x + r0.x, k0.y + r0.y), a1 = make_float2(k1.x + r0.z, k1.y + r0.w);
            const float2 b0 = make_float2(l0.x + r1.x, l0.y + r1.y), b1 = make_float2(l1.x + r1.z, l1.y + r1.w);
            const float2 c0 = make_float2(n0.x + r2.x, n0.y + r2.y), c1 = make_float2(n1.x + r2.z, n1.y + r2.w);
            yprev[0] = make_float2(a0.x + b0.x, a0.y + b0.y);
            yprev[1] = make_float2(b0.x - c0.x, b0.y - c0.y);
            yprev[2] = make_float2(a1.x + b1.x, a1.y + b1.y);
            yprev[3] = make_float2(b1.x - c1.x, b1.y - c1.y);
            prev_cc = cc;
            prev_m0 = (long long)cc * C::CHUNK_OUT + g * R + RO * hw;
            prev_o = p.out + (long long)cs * p.out_stride + prev_m0;
            prev_nout = p.n_out;

            sidx += NG / C::NPROD;
            if (sidx >= scnt) { sidx -= scnt; par ^= 1u; }
            cs += gs;
            cc += gc;
            if (cc >= cps) { cc -= cps; ++cs; }
        }
        w_epilogue<RO>(yprev, rot_thr, p.phase0_fx + (unsigned long long)prev_cc * chunk_dph, prev_m0, prev_o, prev_nout);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Small decimations (D = 4, 8) on the same machinery: output m = NQ m' + q (NQ = 16 / D) is
//      y[NQ m' + q] = sum_k c[k] x[16 m' + D q + k] = sum_k' c_q[k'] x[16 m' + k'],      c_q[k'] = c[k' - D q],
// i.e. NQ interleaved decimate-by-16 filters with SHIFTED tap sets over the SAME staged chunk (T' = T + D (NQ - 1) taps
// each, the same total flops as the direct D-decimating form).  A warp runs the fast FIR NQ times over its chunk, once per
// tap set, and stores run q's outputs at stride NQ.  The tap count is 16 NJG + JL blocks: NJG groups of 16 blocks and a
// last group of JL (even, <= 18) blocks, so that T' = 264 costs 18 blocks, not 32.
// ---------------------------------------------------------------------------------------------------------------------
template <int JT, int NQ>
struct WQCfg : PCfg<16, JT, 1> {
    using B = PCfg<16, JT, 1>;
    static_assert(JT % 2 == 0, "even number of tap blocks");
    // tap groups: NJG groups of 16 blocks, then a last group of JL blocks.  A remainder of 2 blocks is merged into the last
    // group (JL = 18): a 2-block pass would load 9 window entries for 48 FFMA2 and run at a quarter of the FMA rate.
    static constexpr int JL = JT <= 18 ? JT : (JT % 16 == 0 ? 16 : (JT % 16 == 2 ? 18 : JT % 16));
    static constexpr int NJG = (JT - JL) / 16;
    static constexpr int JPA = NJG > 0 ? 16 : JL;  // blocks of the first pass
    static constexpr int RH = B::R / 2;
    static constexpr int TQ4 = 3 * (JT / 2) * 8;   // float4 per tap set (3 (JT/2) 16 complex)
    static constexpr int NTW = NQ * 3 * (JT / 2) * 16;
    static constexpr int WROWS_MAX = 4;            // thread-rows a pass window can span (R + 18 - 1 = 25 blocks)
};

template <int R>
__device__ __forceinline__ void wq_epilogue(const float2 (&y)[R], const float2 (&rot_thr)[R], unsigned long long chunk_phase,
                                            float2* o, int stride, int left) {
    const float2 rot_chunk = nco_rot_bf(chunk_phase);
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const float2 z = cmul(cmul(y[r], rot_thr[r]), rot_chunk);
        st_cs_v2_if(o + (long long)r * stride, z.x, z.y, r < left);
    }
}

template <int JT, int NQ>
__global__ void __launch_bounds__(WQCfg<JT, NQ>::NWARPS * 32 + 32 * WQCfg<JT, NQ>::NPROD, 1)
ddc_fused_wq_kernel(const __grid_constant__ RunParams p, const __grid_constant__ TapsParam<WQCfg<JT, NQ>::NTW> taps) {
    using C = WQCfg<JT, NQ>;
    constexpr int D = 16;   // data geometry: 16-sample blocks, whatever the true decimation
    constexpr int ROW = C::ROW, R = C::R, NWARPS = C::NWARPS, NG = C::NGROUPS;
    constexpr int NSLOT = C::NSLOT, RH = C::RH, NJG = C::NJG, JL = C::JL, JPA = C::JPA;
    constexpr int WANT = C::TOT_ROWS * ROW;
    constexpr int NWA = R + JPA - 1;   // first-pass window
    static_assert(R == 8 && C::V == 4, "geometry");

    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw);
    uint64_t* empty_bar = full_bar + 16;
    volatile int* slot_seq = reinterpret_cast<volatile int*>(smem_raw + 384);
    float* buf = reinterpret_cast<float*>(smem_raw + C::HDR_BYTES);

    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int lane = tid & 31;
    if (tid == 0) {
#pragma unroll 1
        for (int s = 0; s < NSLOT; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
            slot_seq[s] = -1;
        }
        mbar_fence_init();
    }
    __syncthreads();

    const int cps = (int)p.tiles_per_stream;
    const int n_k = (int)((p.total_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);
    const unsigned long long chunk_dph = (unsigned long long)((long long)C::CHUNK_OUT * D) * p.step_fx;

    if (warp >= NWARPS) {
        // ------------------------------------------------------------------ producer warps (as in ddc_fused_w_kernel)
        constexpr int NP = C::NPROD;
        const int pid = warp - NWARPS;
        const long long pstride = (long long)NP * gridDim.x;
        const int gs = (int)(pstride / cps), gc = (int)(pstride % cps);
        const long long pfirst = blockIdx.x + (long long)pid * gridDim.x;
        int cs = (int)(pfirst / cps), cc = (int)(pfirst % cps);
        const int sbase = C::sub_base(pid), scnt = C::sub_count(pid);
        int sidx = 0;
        uint32_t par = 1;
        for (int k = pid; k < n_k; k += NP) {
            const int slot = sbase + sidx;
            const bool leader = elect_one();
            if (leader) {
                mbar_wait(&empty_bar[slot], par);
                slot_seq[slot] = k;
            }
            __syncwarp();
            const float* src = reinterpret_cast<const float*>(p.in) + (long long)cs * p.in_stride + (long long)cc * C::CHUNK_S;
            float* dst = buf + (size_t)slot * C::SLOT_FLOATS;
            const long long valid = p.n_samples - (long long)cc * C::CHUNK_S;
            if (valid >= WANT) {
                if (leader) {
                    mbar_arrive_expect_tx(&full_bar[slot], (uint32_t)WANT * 4u);
#pragma unroll
                    for (int sr = 0; sr < C::NSR; ++sr) {
                        constexpr int SR4 = C::SROWS;
                        const int nrow = (C::TOT_ROWS - sr * SR4) < SR4 ? (C::TOT_ROWS - sr * SR4) : SR4;
                        bulk_g2s(dst + sr * C::SRP, src + sr * SR4 * ROW, (uint32_t)nrow * ROW * 4u, &full_bar[slot]);
                    }
                }
            } else {
                uint32_t tx = 0;
                for (int sr = 0; sr < C::NSR; ++sr) {
                    const int cap = ((C::TOT_ROWS - sr * C::SROWS) < C::SROWS ? (C::TOT_ROWS - sr * C::SROWS) : C::SROWS) * ROW;
                    const long long s0 = (long long)sr * C::SROWS * ROW;
                    long long cnt = valid - s0;
                    cnt = cnt < 0 ? 0 : (cnt > cap ? cap : cnt);
                    const int bulk = (int)cnt & ~3;
                    for (int e = bulk + lane; e < cap; e += 32) dst[sr * C::SRP + e] = (e < (int)cnt) ? src[s0 + e] : 0.f;
                    tx += (uint32_t)bulk * 4u;
                }
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive_expect_tx(&full_bar[slot], tx);
                    for (int sr = 0; sr < C::NSR; ++sr) {
                        const int cap = ((C::TOT_ROWS - sr * C::SROWS) < C::SROWS ? (C::TOT_ROWS - sr * C::SROWS) : C::SROWS) * ROW;
                        const long long s0 = (long long)sr * C::SROWS * ROW;
                        long long cnt = valid - s0;
                        cnt = cnt < 0 ? 0 : (cnt > cap ? cap : cnt);
                        const int bulk = (int)cnt & ~3;
                        if (bulk > 0) bulk_g2s(dst + sr * C::SRP, src + s0, (uint32_t)bulk * 4u, &full_bar[slot]);
                    }
                }
            }
            __syncwarp();
            if (++sidx == scnt) { sidx = 0; par ^= 1u; }
            cs += gs;
            cc += gc;
            if (cc >= cps) { cc -= cps; ++cs; }
        }
    } else {
        // ------------------------------------------------------------------ compute warps
        const int grp = warp;
        const int g = (lane & 7) * C::SROWS + (lane >> 3);
        int rowoff[C::WROWS_MAX];
#pragma unroll
        for (int h = 0; h < C::WROWS_MAX; ++h) rowoff[h] = C::row_offset(g + h);
        float2 rot_thr[R];
#pragma unroll
        for (int r = 0; r < R; ++r) rot_thr[r] = nco_rot((unsigned long long)((long long)(g * R + r) * D) * p.step_fx);

        const long long kstride = (long long)NG * gridDim.x;
        const int gs = (int)(kstride / cps), gc = (int)(kstride % cps);
        const long long first = blockIdx.x + (long long)grp * gridDim.x;
        int cs = (int)(first / cps), cc = (int)(first % cps);
        const int sbase = C::sub_base(grp % C::NPROD), scnt = C::sub_count(grp % C::NPROD);
        int sidx = (grp / C::NPROD) % scnt;
        uint32_t par = (uint32_t)((grp / C::NPROD) / scnt) & 1u;

        // deferred epilogue state: the previous run's sums and where they go
        float2 yprev[R];
#pragma unroll
        for (int r = 0; r < R; ++r) yprev[r] = make_float2(0.f, 0.f);
        float2* prev_o = p.out;
        int prev_cc = 0, prev_left = 0;   // left = 0 disables the stores

        float2 m0a[RH], m1a[RH], m2a[RH];
        for (int k = grp; k < n_k; k += NG) {
            const int slot = sbase + sidx;
            while (slot_seq[slot] != k) {}
            mbar_wait(&full_bar[slot], par);
            const float* sbuf = buf + (size_t)slot * C::SLOT_FLOATS;
            const float4* tq = &taps.c2[0];
#pragma unroll 1
            for (int q = 0; q < NQ; ++q, tq += C::TQ4) {
#pragma unroll
                for (int r = 0; r < RH; ++r) m0a[r] = m1a[r] = m2a[r] = make_float2(0.f, 0.f);
                int xoff = 0;
                const float4* tp = tq;
                {   // first pass (tap group 0, phase group 0) with the previous run's epilogue in its basic block
                    float4 w[NWA];
#pragma unroll
                    for (int b = 0; b < NWA; ++b)
                        w[b] = *reinterpret_cast<const float4*>(sbuf + rowoff[b / R] + (b % R) * D);
                    wq_epilogue<R>(yprev, rot_thr, p.phase0_fx + (unsigned long long)prev_cc * chunk_dph, prev_o, NQ, prev_left);
                    w_fir_pg<D, JPA, R>(w, tp, m0a, m1a, m2a);
                    xoff = 4;
                    tp += 2;
                }
                int grow = g;   // first thread-row of the current pass window
                if constexpr (NJG > 0) {
                    int pgi = 1;
#pragma unroll 1
                    for (int pass = 1; pass < NJG * C::V; ++pass) {
                        asm volatile("" : "+r"(xoff), "+r"(grow));
                        int ro[C::WROWS_MAX];
#pragma unroll
                        for (int h = 0; h < C::WROWS_MAX; ++h) ro[h] = ((grow + h) / C::SROWS) * C::SRP + ((grow + h) % C::SROWS) * ROW;
                        float4 w[R + 15];
#pragma unroll
                        for (int b = 0; b < R + 15; ++b)
                            w[b] = *reinterpret_cast<const float4*>(sbuf + xoff + ro[b / R] + (b % R) * D);
                        w_fir_pg<D, 16, R>(w, tp, m0a, m1a, m2a);
                        ++pgi;
                        xoff += 4;
                        tp += 2;
                        if (pgi == C::V) {
                            pgi = 0;
                            xoff = 0;
                            grow += 16 / R;
                            tp += 3 * 8 * (D / 2) - 2 * C::V;
                        }
                    }
                    // after the last full group the loop leaves tp / grow / xoff at (group NJG, phase group 0)
                }
                {
                    // last tap group: JL blocks starting at block 16 NJG; phase groups 0 .. 3 (1 .. 3 when it was also the
                    // first pass)
                    constexpr int NWL = R + JL - 1;
                    constexpr int WROWS_L = (NWL - 1) / R + 1;
#pragma unroll 1
                    for (int pg = (NJG > 0 ? 0 : 1); pg < C::V; ++pg, tp += 2) {
                        asm volatile("" : "+r"(xoff), "+r"(grow));
                        int ro[WROWS_L];
#pragma unroll
                        for (int h = 0; h < WROWS_L; ++h) ro[h] = ((grow + h) / C::SROWS) * C::SRP + ((grow + h) % C::SROWS) * ROW;
                        float4 w[NWL];
#pragma unroll
                        for (int b = 0; b < NWL; ++b)
                            w[b] = *reinterpret_cast<const float4*>(sbuf + xoff + ro[b / R] + (b % R) * D);
                        w_fir_pg<D, JL, R>(w, tp, m0a, m1a, m2a);
                        xoff += 4;
                    }
                }
                // the chunk goes back to the producer after the last tap set
                __syncwarp();
                mbar_arrive_if(&empty_bar[slot], lane == 0 && q == NQ - 1);

#pragma unroll
                for (int r = 0; r < RH; ++r) {
                    yprev[2 * r] = make_float2(m0a[r].x + m1a[r].x, m0a[r].y + m1a[r].y);
                    yprev[2 * r + 1] = make_float2(m1a[r].x - m2a[r].x, m1a[r].y - m2a[r].y);
                }
                // outputs of this run: m = NQ m' + q for m' = m0 .. m0 + R - 1; those with m < n_out exist
                const long long m0 = (long long)cc * C::CHUNK_OUT + g * R;
                const long long nq = (p.n_out - q + NQ - 1) / NQ;
                const long long lf = nq - m0;
                prev_left = (int)(lf < 0 ? 0 : (lf > R ? R : lf));
                prev_cc = cc;
                prev_o = p.out + (long long)cs * p.out_stride + m0 * NQ + q;
            }
            sidx += NG / C::NPROD;
            if (sidx >= scnt) { sidx -= scnt; par ^= 1u; }
            cs += gs;
            cc += gc;
            if (cc >= cps) { cc -= cps; ++cs; }
        }
        wq_epilogue<R>(yprev, rot_thr, p.phase0_fx + (unsigned long long)prev_cc * chunk_dph, prev_o, NQ, prev_left);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// packed 10-bit input (BASELINE configs[2]; reference stub ddc.py:68-83): raw TMA ring + in-warp unpack as in
// ddc_fused_p10_kernel, with the fast FIR and the deferred branch-free epilogue.  The unpack avoids I2F (quarter-rate
// conversion pipe): the 10 bits are placed in the mantissa of 2^23 with the sign bit flipped,
//      as_float(((word >> s) & 0x3FF) ^ 0x4B000200) = 2^23 + (v + 512),      v = that value - (2^23 + 512)   (exact),
// i.e. one shift, one LOP3 and one FADD per sample -- bit-exact for all 1024 codes (tests/test_gpu_parity.py).
//
// Private float buffer layout: a lane unpacks 16 consecutive samples (one D = 16 block: four 16-byte units) per step, so
// the eight lanes of a quarter warp store to addresses 64 bytes apart -- only two distinct bank groups, a 4-way conflict
// that cost 0.37 ms of 1.35 ms (shared-memory store bandwidth is per SM).  The four units of block c are therefore
// stored ROTATED by (c >> 1): logical unit u lives at physical unit (u + (c >> 1)) & 3.  Writers become conflict-free;
// FIR readers (all lanes read the same block and unit of different rows) only see a different constant offset.
// ---------------------------------------------------------------------------------------------------------------------
template <int D, int JT>
struct W10Cfg : P10Cfg<D, JT> {
    using B = PCfg<D, JT, 1>;
    static_assert(JT % 2 == 0 && B::R % 2 == 0, "fast FIR needs an even number of tap blocks and outputs per thread");
    static constexpr int JH = JT / 2, RH = B::R / 2;
    static constexpr int NTW = 3 * JH * D;
    // physical float offset inside a row of the 16-byte unit that logically starts at float offset fo (multiple of 4)
    __host__ __device__ static constexpr int rot_off(int fo) { return (fo & ~15) + 4 * ((((fo >> 2) & 3) + ((fo >> 5) & 3)) & 3); }
};

__device__ __forceinline__ float unpack10_bits(uint32_t field, uint32_t k4b000200) {   // field: the 10 bits in [9:0], anything above
    uint32_t bits;   // (field & 0x3FF) ^ 0x4B000200 as ONE LOP3 (the constant must sit in a register for that)
    asm("lop3.b32 %0, %1, 0x3FF, %2, 0x6A;" : "=r"(bits) : "r"(field), "r"(k4b000200));
    return __uint_as_float(bits);   // = 2^23 + 512 + v; the caller subtracts 8389120 (two samples per FADD2)
}

template <int D, int JT>
__global__ void __launch_bounds__(W10Cfg<D, JT>::NWARPS * 32 + 32 * W10Cfg<D, JT>::NPROD, 1)
ddc_fused_w10_kernel(const __grid_constant__ RunParams p, const __grid_constant__ TapsParam<W10Cfg<D, JT>::NTW> taps) {
    using C = W10Cfg<D, JT>;
    constexpr int ROW = C::ROW, R = C::R, NW = C::NW, NWARPS = C::NWARPS, NG = C::NGROUPS;
    constexpr int NRAW = C::NRAW, RAWB = C::RAW_BYTES, RH = C::RH;
    constexpr int WANT = C::TOT_ROWS * ROW;   // samples staged per chunk

    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw);          // [16]
    uint64_t* empty_bar = full_bar + 16;                                 // [16]
    volatile int* slot_seq = reinterpret_cast<volatile int*>(smem_raw + 384);
    float* fbuf = reinterpret_cast<float*>(smem_raw + 512);              // NG private float buffers
    unsigned char* rbuf = smem_raw + 512 + C::FLOAT_BYTES;               // NRAW raw slots

    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int lane = tid & 31;
    if (tid == 0) {
#pragma unroll 1
        for (int s = 0; s < NRAW; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
            slot_seq[s] = -1;
        }
        mbar_fence_init();
    }
    __syncthreads();

    const int cps = (int)p.tiles_per_stream;
    const int n_k = (int)((p.total_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);
    const unsigned long long chunk_dph = (unsigned long long)((long long)C::CHUNK_OUT * D) * p.step_fx;

    if (warp >= NWARPS) {
        // ------------------------------------------------------------------ producer warps (one bulk copy per chunk)
        constexpr int NP = C::NPROD;
        const int pid = warp - NWARPS;
        const long long pstride = (long long)NP * gridDim.x;
        const int gs = (int)(pstride / cps), gc = (int)(pstride % cps);
        const long long pfirst = blockIdx.x + (long long)pid * gridDim.x;
        int cs = (int)(pfirst / cps), cc = (int)(pfirst % cps);
        const int sbase = C::sub_base(pid), scnt = C::sub_count(pid);
        int sidx = 0;
        uint32_t par = 1;
        for (int k = pid; k < n_k; k += NP) {
            const int slot = sbase + sidx;
            if (lane == 0) {
                mbar_wait(&empty_bar[slot], par);
                slot_seq[slot] = k;
            }
            __syncwarp();
            const unsigned char* src = reinterpret_cast<const unsigned char*>(p.in) + (long long)cs * p.in_stride +
                                       (long long)cc * (C::CHUNK_S / 4 * 5);
            unsigned char* dst = rbuf + (size_t)slot * RAWB;
            const long long valid = p.n_samples - (long long)cc * C::CHUNK_S;   // samples (multiple of 4)
            if (valid >= WANT) {
                if (lane == 0) {
                    mbar_arrive_expect_tx(&full_bar[slot], (uint32_t)RAWB);
                    bulk_g2s(dst, src, (uint32_t)RAWB, &full_bar[slot]);
                }
            } else {
                const int vb = (int)(valid > 0 ? valid / 4 * 5 : 0);   // valid bytes
                const int bulk = vb & ~15;
                for (int e = bulk + lane; e < RAWB; e += 32) dst[e] = (e < vb) ? src[e] : (unsigned char)0;
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive_expect_tx(&full_bar[slot], (uint32_t)bulk);
                    if (bulk > 0) bulk_g2s(dst, src, (uint32_t)bulk, &full_bar[slot]);
                }
            }
            __syncwarp();
            if (++sidx == scnt) { sidx = 0; par ^= 1u; }
            cs += gs;
            cc += gc;
            if (cc >= cps) { cc -= cps; ++cs; }
        }
    } else {
        // ------------------------------------------------------------------ compute warps
        const int grp = warp;
        const int g = (lane & 7) * C::SROWS + (lane >> 3);
        int rowoff[C::HALO_ROWS + 1];
#pragma unroll
        for (int h = 0; h <= C::HALO_ROWS; ++h) rowoff[h] = C::row_offset(g + h);
        float2 rot_thr[R];
#pragma unroll
        for (int r = 0; r < R; ++r) rot_thr[r] = nco_rot((unsigned long long)((long long)(g * R + r) * D) * p.step_fx);
        float* sbuf = fbuf + (size_t)grp * C::SLOT_FLOATS;

        const long long kstride = (long long)NG * gridDim.x;
        const int gs = (int)(kstride / cps), gc = (int)(kstride % cps);
        const long long first = blockIdx.x + (long long)grp * gridDim.x;
        int cs = (int)(first / cps), cc = (int)(first % cps);
        const int sbase = C::sub_base(grp % C::NPROD), scnt = C::sub_count(grp % C::NPROD);
        int sidx = (grp / C::NPROD) % scnt;
        uint32_t par = (uint32_t)((grp / C::NPROD) / scnt) & 1u;

        float2 yprev[R];
#pragma unroll
        for (int r = 0; r < R; ++r) yprev[r] = make_float2(0.f, 0.f);
        long long prev_m0 = 0;
        float2* prev_o = p.out;
        int prev_cc = 0;
        long long prev_nout = 0;
        float2 m0a[RH], m1a[RH], m2a[RH];
        static_assert(D == 16 || D == 32 || D == 64, "unit rotation assumes whole 16-sample groups per block");
        // where this lane's four 16-byte units of a 16-sample group go (floats, relative to the group): rotation by the
        // index of the group within its 128-sample row, halved
        int wr_unit[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) wr_unit[q] = 4 * ((q + ((lane & 7) >> 1)) & 3);
        uint32_t kmagic = 0x4B000200u;
        asm volatile("" : "+r"(kmagic));   // keep it in a register (see unpack10_f)

        long long t_wait = 0;
        const long long t_begin = clock64();
        const int dm = p.debug_mode & 255;   // tuning aids: 3 = unpack only (no FIR), 4 = FIR only (no unpack)
        for (int k = grp; k < n_k; k += NG) {
            const int slot = sbase + sidx;
            {
                const long long tw0 = p.dbg ? clock64() : 0;
                while (slot_seq[slot] != k) {}
                mbar_wait(&full_bar[slot], par);
                if (p.dbg) t_wait += clock64() - tw0;
            }
            // ---- unpack: 16 samples (20 bytes = 5 words) per step and lane; integer work, bit-exact
            const uint32_t* rw = reinterpret_cast<const uint32_t*>(rbuf + (size_t)slot * RAWB);
            constexpr int NSG = WANT / 16;   // 16-sample groups per chunk (272)
            constexpr int UNPACK_UNROLL = DDCB200_W10_UNPACK_UNROLL;
#pragma unroll UNPACK_UNROLL
            for (int sg = lane; sg < (dm == 4 ? 0 : NSG); sg += 32) {
                uint32_t w[5];
#pragma unroll
                for (int i = 0; i < 5; ++i) w[i] = __byte_perm(rw[sg * 5 + i], 0, 0x0123);   // big-endian words
                float v[16];
#pragma unroll
                for (int s16 = 0; s16 < 16; ++s16) {
                    // sample s16 occupies bits [10 s16, 10 s16 + 10) of the 160-bit big-endian group: right-align it
                    const int bit = 10 * s16, wi = bit >> 5, sh = bit & 31;
                    const uint32_t fld = (sh <= 22) ? (w[wi] >> (22 - sh)) : __funnelshift_r(w[wi + 1 > 4 ? 4 : wi + 1], w[wi], 54 - sh);
                    v[s16] = unpack10_bits(fld, kmagic);
                }
#pragma unroll
                for (int s16 = 0; s16 < 16; s16 += 2) {
                    const float2 d = __fadd2_rn(make_float2(v[s16], v[s16 + 1]), make_float2(-8389120.0f, -8389120.0f));
                    v[s16] = d.x;
                    v[s16 + 1] = d.y;
                }
                // group sg = block (sg & 7) of row (sg >> 3); sg & 7 == lane & 7 in every step
                float* blk = sbuf + C::row_offset(sg >> 3) + (lane & 7) * 16;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    *reinterpret_cast<float4*>(blk + wr_unit[q]) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[slot]);   // raw slot back to the producer

#pragma unroll
            for (int r = 0; r < RH; ++r) m0a[r] = m1a[r] = m2a[r] = make_float2(0.f, 0.f);
            int xoff = 0;
            const float4* tp = &taps.c2[0];
            if (dm != 3) {
                float4 w[NW];
#pragma unroll
                for (int b = 0; b < NW; ++b)
                    w[b] = *reinterpret_cast<const float4*>(sbuf + rowoff[b / R] + C::rot_off((b % R) * D));
                w_epilogue<R>(yprev, rot_thr, p.phase0_fx + (unsigned long long)prev_cc * chunk_dph, prev_m0, prev_o, prev_nout);
                w_fir_pg<D, JT, R>(w, tp, m0a, m1a, m2a);
                xoff = 4;
                tp += 2;
            }
#pragma unroll 1
            for (int pg = 1; pg < (dm == 3 ? 0 : C::V); ++pg, tp += 2) {
                asm volatile("" : "+r"(xoff));
                int xrot[4];   // D = 16: physical float offset of logical unit xoff / 4 under rotation 0 .. 3
#pragma unroll
                for (int rt = 0; rt < 4; ++rt) xrot[rt] = (xoff + 4 * rt) & 12;
                float4 w[NW];
#pragma unroll
                for (int b = 0; b < NW; ++b) {
                    // float offset (b % R) D + xoff inside the row: 16-sample group ((b % R) D + xoff) / 16, unit (xoff / 4) & 3
                    const int grp16 = ((b % R) * D) / 16;          // + xoff / 16, which is 0 for D = 16 (xoff < 16)
                    if (D == 16) {
                        w[b] = *reinterpret_cast<const float4*>(sbuf + rowoff[b / R] + grp16 * 16 + xrot[(grp16 >> 1) & 3]);
                    } else {
                        const int fo = (b % R) * D + xoff;
                        w[b] = *reinterpret_cast<const float4*>(sbuf + rowoff[b / R] + (fo & ~15) + 4 * ((((fo >> 2) & 3) + ((fo >> 5) & 3)) & 3));
                    }
                }
                xoff += 4;
                w_fir_pg<D, JT, R>(w, tp, m0a, m1a, m2a);
            }
            __syncwarp();   // every lane is done with the private buffer before the next unpack overwrites it

#pragma unroll
            for (int r = 0; r < RH; ++r) {
                yprev[2 * r] = make_float2(m0a[r].x + m1a[r].x, m0a[r].y + m1a[r].y);
                yprev[2 * r + 1] = make_float2(m1a[r].x - m2a[r].x, m1a[r].y - m2a[r].y);
            }
            prev_cc = cc;
            prev_m0 = (long long)cc * C::CHUNK_OUT + g * R;
            prev_o = p.out + (long long)cs * p.out_stride + prev_m0;
            prev_nout = p.n_out;

            sidx += NG / C::NPROD;
            if (sidx >= scnt) { sidx -= scnt; par ^= 1u; }
            cs += gs;
            cc += gc;
            if (cc >= cps) { cc -= cps; ++cs; }
        }
        w_epilogue<R>(yprev, rot_thr, p.phase0_fx + (unsigned long long)prev_cc * chunk_dph, prev_m0, prev_o, prev_nout);
        if (p.dbg && lane == 0) {
            atomicAdd(p.dbg, (unsigned long long)t_wait);
            atomicAdd(p.dbg + 1, (unsigned long long)(clock64() - t_begin));
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Packed input, warp-specialised: in ddc_fused_w10_kernel every compute warp unpacks its own chunk and then filters it, so
// the FMA pipe idles during the latency-bound unpack (0.19 of 1.13 ms).  Here FOUR UNPACK WARPS turn raw chunks into a
// ring of float chunks (same rotated-unit layout) and EIGHT FIR WARPS consume them exactly like the float32 kernel; the
// integer / LSU work of the unpackers runs under the FIR warps' FFMA2 stream.
//      TMA producer warp -> raw ring (NR x 5440 B) -> unpack warps -> float ring (NF x 17.1 KB) -> FIR warps -> HBM
// Float-ring hand-over uses sequence words in shared memory (ready[slot] = chunk, done[slot] = chunk) rather than
// mbarrier parities: successive uses of a slot are produced and consumed by different warps, which may arrive more than
// one phase early, and a parity cannot tell that from "done".
// ---------------------------------------------------------------------------------------------------------------------
#ifndef DDCB200_W10S_NUNP
#define DDCB200_W10S_NUNP 4
#endif
#ifndef DDCB200_W10S_NR
#define DDCB200_W10S_NR 6
#endif
#ifndef DDCB200_W10S_UNROLL
#define DDCB200_W10S_UNROLL 1
#endif
template <int D, int JT>
struct W10SCfg : PCfg<D, JT, 1> {
    using B = PCfg<D, JT, 1>;
    static_assert(JT % 2 == 0 && B::R % 2 == 0 && D == 16, "fast FIR, 16-sample blocks");
    static constexpr int RH = B::R / 2;
    static constexpr int NTW = 3 * (JT / 2) * D;
    static constexpr int NFIR = 8, NUNP = DDCB200_W10S_NUNP;                       // FIR warps, unpack warps (+ 1 producer warp)
    static constexpr int RAW_BYTES = B::TOT_ROWS * B::ROW / 4 * 5;                 // 5440 for 34 rows
    static constexpr int SLOT_BYTES = B::SLOT_FLOATS * 4;
    static constexpr int HDR = 1024;
    static constexpr int NR = DDCB200_W10S_NR;                                     // raw slots
    static constexpr int NF = (227 * 1024 - HDR - NR * RAW_BYTES) / SLOT_BYTES;    // float slots (11 for J = 16)
    static_assert(NF >= NFIR + 2, "float ring too small");
    static_assert(RAW_BYTES % 16 == 0, "raw chunk must be a whole number of 16-byte groups");
    static constexpr int SMEM = HDR + NF * SLOT_BYTES + NR * RAW_BYTES;
    __host__ __device__ static constexpr int rot_off(int fo) { return (fo & ~15) + 4 * ((((fo >> 2) & 3) + ((fo >> 5) & 3)) & 3); }
};

template <int D, int JT>
__global__ void __launch_bounds__((W10SCfg<D, JT>::NFIR + W10SCfg<D, JT>::NUNP + 1) * 32, 1)
ddc_fused_w10s_kernel(const __grid_constant__ RunParams p, const __grid_constant__ TapsParam<W10SCfg<D, JT>::NTW> taps) {
    using C = W10SCfg<D, JT>;
    constexpr int ROW = C::ROW, R = C::R, NW = C::NW, NFIR = C::NFIR, NUNP = C::NUNP, NR = C::NR, NF = C::NF;
    constexpr int RAWB = C::RAW_BYTES, RH = C::RH;
    constexpr int WANT = C::TOT_ROWS * ROW;   // samples staged per chunk

    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* raw_full = reinterpret_cast<uint64_t*>(smem_raw);                   // [8]
    uint64_t* raw_empty = raw_full + 8;                                           // [8]
    volatile int* raw_seq = reinterpret_cast<volatile int*>(smem_raw + 128);      // [8]   chunk staged in each raw slot
    volatile int* f_ready = reinterpret_cast<volatile int*>(smem_raw + 256);      // [16]  chunk unpacked into each float slot
    volatile int* f_done = reinterpret_cast<volatile int*>(smem_raw + 384);       // [16]  chunk last consumed from each float slot
    float* fbuf = reinterpret_cast<float*>(smem_raw + C::HDR);
    unsigned char* rbuf = smem_raw + C::HDR + NF * C::SLOT_BYTES;

    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int lane = tid & 31;
    if (tid == 0) {
#pragma unroll 1
        for (int s = 0; s < NR; ++s) {
            mbar_init(&raw_full[s], 1);
            mbar_init(&raw_empty[s], 1);
            raw_seq[s] = -1;
        }
#pragma unroll 1
        for (int s = 0; s < 16; ++s) {
            f_ready[s] = -1;
            f_done[s] = -1;
        }
        mbar_fence_init();
    }
    __syncthreads();

    const int cps = (int)p.tiles_per_stream;
    const int n_k = (int)((p.total_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);
    const unsigned long long chunk_dph = (unsigned long long)((long long)C::CHUNK_OUT * D) * p.step_fx;

    if (warp == NFIR + NUNP) {
        // ------------------------------------------------------------------ TMA producer warp: one bulk copy per chunk
        const int gs = (int)((long long)gridDim.x / cps), gc = (int)((long long)gridDim.x % cps);
        int cs = (int)(blockIdx.x / cps), cc = (int)(blockIdx.x % cps);
        for (int k = 0; k < n_k; ++k) {
            const int slot = k % NR;
            if (lane == 0) {
                if (k >= NR) mbar_wait(&raw_empty[slot], (uint32_t)((k / NR) - 1) & 1u);   // single waiter, in order
                raw_seq[slot] = k;
            }
            __syncwarp();
            const unsigned char* src = reinterpret_cast<const unsigned char*>(p.in) + (long long)cs * p.in_stride +
                                       (long long)cc * (C::CHUNK_S / 4 * 5);
            unsigned char* dst = rbuf + (size_t)slot * RAWB;
            const long long valid = p.n_samples - (long long)cc * C::CHUNK_S;   // samples (multiple of 4)
            if (valid >= WANT) {
                if (lane == 0) {
                    mbar_arrive_expect_tx(&raw_full[slot], (uint32_t)RAWB);
                    bulk_g2s(dst, src, (uint32_t)RAWB, &raw_full[slot]);
                }
            } else {
                const int vb = (int)(valid > 0 ? valid / 4 * 5 : 0);   // valid bytes
                const int bulk = vb & ~15;
                for (int e = bulk + lane; e < RAWB; e += 32) dst[e] = (e < vb) ? src[e] : (unsigned char)0;
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive_expect_tx(&raw_full[slot], (uint32_t)bulk);
                    if (bulk > 0) bulk_g2s(dst, src, (uint32_t)bulk, &raw_full[slot]);
                }
            }
            __syncwarp();
            cs += gs;
            cc += gc;
            if (cc >= cps) { cc -= cps; ++cs; }
        }
    } else if (warp >= NFIR) {
        // ------------------------------------------------------------------ unpack warps: chunk k by warp k % NUNP
        const int u = warp - NFIR;
        int wr_unit[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) wr_unit[q] = 4 * ((q + ((lane & 7) >> 1)) & 3);
        uint32_t kmagic = 0x4B000200u;
        asm volatile("" : "+r"(kmagic));
        for (int k = u; k < n_k; k += NUNP) {
            const int rs = k % NR, fs = k % NF;
            while (raw_seq[rs] != k) {}
            mbar_wait(&raw_full[rs], (uint32_t)(k / NR) & 1u);
            if (k >= NF)
                while (f_done[fs] != k - NF) {}     // the FIR warp has finished the chunk that lived in this float slot
            __threadfence_block();
            __syncwarp();
            const uint32_t* rw = reinterpret_cast<const uint32_t*>(rbuf + (size_t)rs * RAWB);
            float* sbuf = fbuf + (size_t)fs * C::SLOT_FLOATS;
            constexpr int NSG = WANT / 16;   // 16-sample groups per chunk (272)
            constexpr int UNR = DDCB200_W10S_UNROLL;
#pragma unroll UNR
            for (int sg = lane; sg < NSG; sg += 32) {
                uint32_t w[5];
#pragma unroll
                for (int i = 0; i < 5; ++i) w[i] = __byte_perm(rw[sg * 5 + i], 0, 0x0123);   // big-endian words
                float v[16];
#pragma unroll
                for (int s16 = 0; s16 < 16; ++s16) {
                    const int bit = 10 * s16, wi = bit >> 5, sh = bit & 31;
                    const uint32_t fld = (sh <= 22) ? (w[wi] >> (22 - sh)) : __funnelshift_r(w[wi + 1 > 4 ? 4 : wi + 1], w[wi], 54 - sh);
                    v[s16] = unpack10_bits(fld, kmagic);
                }
#pragma unroll
                for (int s16 = 0; s16 < 16; s16 += 2) {
                    const float2 d = __fadd2_rn(make_float2(v[s16], v[s16 + 1]), make_float2(-8389120.0f, -8389120.0f));
                    v[s16] = d.x;
                    v[s16 + 1] = d.y;
                }
                float* blk = sbuf + C::row_offset(sg >> 3) + (lane & 7) * 16;
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    *reinterpret_cast<float4*>(blk + wr_unit[q]) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            }
            __syncwarp();
            __threadfence_block();
            if (lane == 0) {
                mbar_arrive(&raw_empty[rs]);   // raw slot back to the producer
                f_ready[fs] = k;               // float chunk published
            }
        }
    } else {
        // ------------------------------------------------------------------ FIR warps: chunk k by warp k % NFIR
        const int grp = warp;
        const int g = (lane & 7) * C::SROWS + (lane >> 3);
        int rowoff[C::HALO_ROWS + 1];
#pragma unroll
        for (int h = 0; h <= C::HALO_ROWS; ++h) rowoff[h] = C::row_offset(g + h);
        float2 rot_thr[R];
#pragma unroll
        for (int r = 0; r < R; ++r) rot_thr[r] = nco_rot((unsigned long long)((long long)(g * R + r) * D) * p.step_fx);

        const long long kstride = (long long)NFIR * gridDim.x;
        const int gs = (int)(kstride / cps), gc = (int)(kstride % cps);
        const long long first = blockIdx.x + (long long)grp * gridDim.x;
        int cs = (int)(first / cps), cc = (int)(first % cps);

        float2 yprev[R];
#pragma unroll
        for (int r = 0; r < R; ++r) yprev[r] = make_float2(0.f, 0.f);
        long long prev_m0 = 0;
        float2* prev_o = p.out;
        int prev_cc = 0;
        long long prev_nout = 0;
        float2 m0a[RH], m1a[RH], m2a[RH];
        long long t_wait = 0;
        const long long t_begin = clock64();

        for (int k = grp; k < n_k; k += NFIR) {
            const int fs = k % NF;
            {
                const long long tw0 = p.dbg ? clock64() : 0;
                while (f_ready[fs] != k) {}
                if (p.dbg) t_wait += clock64() - tw0;
            }
            __threadfence_block();   // acquire: the unpackers' stores to the slot are visible before my loads
            __syncwarp();
            const float* sbuf = fbuf + (size_t)fs * C::SLOT_FLOATS;
#pragma unroll
            for (int r = 0; r < RH; ++r) m0a[r] = m1a[r] = m2a[r] = make_float2(0.f, 0.f);
            int xoff = 0;
            const float4* tp = &taps.c2[0];
            {
                float4 w[NW];
#pragma unroll
                for (int b = 0; b < NW; ++b)
                    w[b] = *reinterpret_cast<const float4*>(sbuf + rowoff[b / R] + C::rot_off((b % R) * D));
                w_epilogue<R>(yprev, rot_thr, p.phase0_fx + (unsigned long long)prev_cc * chunk_dph, prev_m0, prev_o, prev_nout);
                w_fir_pg<D, JT, R>(w, tp, m0a, m1a, m2a);
                xoff = 4;
                tp += 2;
            }
#pragma unroll 1
            for (int pg = 1; pg < C::V; ++pg, tp += 2) {
                asm volatile("" : "+r"(xoff));
                int xrot[4];
#pragma unroll
                for (int rt = 0; rt < 4; ++rt) xrot[rt] = (xoff + 4 * rt) & 12;
                float4 w[NW];
#pragma unroll
                for (int b = 0; b < NW; ++b) {
                    const int grp16 = b % R;   // D = 16: block index within the row
                    w[b] = *reinterpret_cast<const float4*>(sbuf + rowoff[b / R] + grp16 * 16 + xrot[(grp16 >> 1) & 3]);
                }
                xoff += 4;
                w_fir_pg<D, JT, R>(w, tp, m0a, m1a, m2a);
            }
            __syncwarp();                       // every lane has read its last window
            if (lane == 0) f_done[fs] = k;      // float slot back to the unpackers

#pragma unroll
            for (int r = 0; r < RH; ++r) {
                yprev[2 * r] = make_float2(m0a[r].x + m1a[r].x, m0a[r].y + m1a[r].y);
                yprev[2 * r + 1] = make_float2(m1a[r].x - m2a[r].x, m1a[r].y - m2a[r].y);
            }
            prev_cc = cc;
            prev_m0 = (long long)cc * C::CHUNK_OUT + g * R;
            prev_o = p.out + (long long)cs * p.out_stride + prev_m0;
            prev_nout = p.n_out;
            cs += gs;
            cc += gc;
            if (cc >= cps) { cc -= cps; ++cs; }
        }
        w_epilogue<R>(yprev, rot_thr, p.phase0_fx + (unsigned long long)prev_cc * chunk_dph, prev_m0, prev_o, prev_nout);
        if (p.dbg && lane == 0) {
            atomicAdd(p.dbg, (unsigned long long)t_wait);
            atomicAdd(p.dbg + 1, (unsigned long long)(clock64() - t_begin));
        }
    }
}

}  // namespace ddck
