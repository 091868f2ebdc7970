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
    static constexpr int NTW = 3 * (JT / 2) * D;   // complex taps passed to the kernel
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
template <int D, int JT>
__global__ void __launch_bounds__(WCfg<D, JT>::NWARPS * 32 + 32 * WCfg<D, JT>::NPROD, 1)
ddc_fused_w_kernel(const __grid_constant__ RunParams p,
                   const __grid_constant__ TapsParam<WCfg<D, JT>::NTW> taps) {
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

        float2 m0a[RH], m1a[RH], m2a[RH];
        constexpr int TPG = 3 * (JP / 2) * (D / 2);   // float4 per tap group of JP blocks
        auto fir = [&](const float4(&w)[NW], const float4* tpp) { w_fir_pg<D, JP, R>(w, tpp, m0a, m1a, m2a); };
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

}  // namespace ddck
