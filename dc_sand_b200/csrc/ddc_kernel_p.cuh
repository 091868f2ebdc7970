// Phase-major fused DDC kernel ("kernel P") for short polyphase branches (J = T / D <= 16 tap blocks).
//
// Why a second kernel: micro-benchmarks on B200 (tools/microbench/fma_bench*.cu, results in profiles/) show that an
// FFMA2 stream only stays near the FP32 peak if (a) each tap fetched through the uniform datapath (LDCU -> UR) feeds
// >= 8 FFMA2, (b) each LDS.128 feeds >= 32 FFMA2, (c) every other vector instruction is rare (each one costs ~0.7 FMA
// issue cycles) and (d) 16 warps per SM are resident to cover LDCU / LDS bubbles.  The rotating-window kernel
// (ddc_fused_kernel, R = 4 outputs per thread at D = 16) has one tap fetch per 4 FFMA2 and one LDS.128 per 16 and tops
// out near 72 % of the FP32 peak.  Kernel P computes R = 128 / D outputs per thread (8 at D = 16) with fully static
// register indexing:
//
//      for pg in my phase groups:                     (phase group = one float4 of every D-sample block)
//          w[b] = LDS.128(block b, phase group pg)    b = 0 .. R+J-2   (the compiler streams this window, ~9 live)
//          for j in 0 .. J-1, ph in 0 .. 3:  tap = c[j*D + 4*pg + ph]       (LDCU -> uniform register pair)
//              for r in 0 .. R-1:  acc[r] += w[r + j].ph * tap              (FFMA2, broadcast sample x complex tap)
//
// Work distribution: a "chunk" is 32 thread-rows (one per lane) of ROW = R*D = 128 samples = 32*R outputs, staged by
// TMA as eight 2 KB super-rows (+16 B pad each, so a quarter warp's LDS.128 hit eight distinct bank groups) plus a
// halo.  NSLOT chunk slots form a ring filled by the producer warp; a chunk is computed by KS warps (w, w + 8), each
// taking D/4/KS phase groups, which exchange partial sums through the (by then dead) chunk slot and finish R/KS
// outputs each.  Warp groups run decoupled from each other: there is no CTA-wide barrier after start-up.
#pragma once
#include "ddc_common.cuh"
#ifndef DDCB200_NPROD
#define DDCB200_NPROD 2
#endif
#ifndef DDCB200_PROD_ROT
#define DDCB200_PROD_ROT 0
#endif

namespace ddck {

template <int D, int JT, int KS>
struct PCfg {
    static constexpr int ROW = 128;                  // samples per thread-row
    static constexpr int R = ROW / D;                // outputs per thread-row
    static constexpr int V = D / 4;                  // phase groups
    static constexpr int VPW = V / KS;               // phase groups per warp
    static constexpr int RO = (R >= KS) ? R / KS : 1;  // outputs a thread finishes
    static constexpr int NW = R + JT - 1;            // window: sample blocks a thread touches
    static constexpr int SROWS = 4;                  // thread-rows per super-row
    static constexpr int SRP = SROWS * ROW + 4;      // super-row pitch (floats): 2 KB + 16 B
    static constexpr int HALO_ROWS = (NW - 1) / R;   // rows past a thread's own that it reads
    static constexpr int CHUNK_ROWS = 32;
    static constexpr int CHUNK_OUT = CHUNK_ROWS * R; // outputs per chunk
    static constexpr int CHUNK_S = CHUNK_ROWS * ROW; // 4096 samples
    static constexpr int NGROUPS = 8;                // chunks in flight in the compute warps
    static constexpr int NWARPS = NGROUPS * KS;      // compute warps
    static constexpr int NPROD = DDCB200_NPROD;                  // producer warps (chunk k is staged by producer k % NPROD)
    static constexpr int TOT_ROWS = CHUNK_ROWS + HALO_ROWS;
    static constexpr int NSR = (TOT_ROWS + SROWS - 1) / SROWS;   // super-rows staged per chunk (last may be partial)
    static constexpr int SLOT_FLOATS = (TOT_ROWS / SROWS) * SRP + ((TOT_ROWS % SROWS) ? (TOT_ROWS % SROWS) * ROW + 4 : 0);
    static constexpr int XBUF_BYTES = (KS == 2) ? NGROUPS * 2 * 32 * RO * 8 : 0;   // partial-sum exchange between partner warps
    static constexpr int HDR_BYTES = 512 + XBUF_BYTES;  // barriers + per-slot chunk rotation + exchange buffer
    static constexpr int NSLOT_MAX = (227 * 1024 - HDR_BYTES) / (SLOT_FLOATS * 4);
    static constexpr int NSLOT = NSLOT_MAX > 13 ? 13 : NSLOT_MAX;  // 8 chunks in use + up to 5 in flight
    static_assert(D % 4 == 0 && ROW % D == 0, "D must divide 128 and be a multiple of 4");
    static_assert(V % KS == 0, "phase groups must split evenly over the warps of a group");
    static_assert(KS == 1 || (R % KS == 0), "outputs must split evenly over the warps of a group");
    static_assert(NSLOT >= NGROUPS + 2, "not enough shared memory for the chunk ring");
    static_assert(NSLOT <= 16, "header layout");
    static_assert(NPROD == 1 || NPROD == 2, "one or two producer warps");
    static_assert(NGROUPS % NPROD == 0 && NSLOT / NPROD > NGROUPS / NPROD, "each sub-ring needs look-ahead slots");
    __host__ __device__ static constexpr int row_offset(int row) { return (row / SROWS) * SRP + (row % SROWS) * ROW; }
    // producer p owns slots [sub_base(p), sub_base(p) + sub_count(p))
    __host__ __device__ static constexpr int sub_count(int pr) { return NPROD == 1 ? NSLOT : (pr == 0 ? (NSLOT + 1) / 2 : NSLOT / 2); }
    __host__ __device__ static constexpr int sub_base(int pr) { return (NPROD == 1 || pr == 0) ? 0 : (NSLOT + 1) / 2; }
};

// =============================================================================================================
// The kernel: ring and FIR as described above, and the epilogue of chunk i (NCO rotation,
// address arithmetic, stores: ~160 mostly dependent instructions) is DEFERRED into the first phase-group pass of chunk
// i + 1, where it sits in the same basic block as 512 independent FFMA2 and the scheduler hides its latency chains
// between them.  Stand-alone, every epilogue left the FMA pipe to the other warp of the sub-partition for ~500 cycles.
// =============================================================================================================
template <int D, int JT, int MAXT>
__global__ void __launch_bounds__(PCfg<D, JT, 1>::NWARPS * 32 + 32 * PCfg<D, JT, 1>::NPROD, 1)
ddc_fused_pd_kernel(const __grid_constant__ RunParams p, const __grid_constant__ TapsParam<MAXT> taps) {
    using C = PCfg<D, JT, 1>;
    constexpr int ROW = C::ROW, R = C::R, NW = C::NW, SRP = C::SRP, NWARPS = C::NWARPS, NG = C::NGROUPS;
    constexpr int NSLOT = C::NSLOT;
    constexpr int WANT = C::TOT_ROWS * ROW;

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
        // ------------------------------------------------------------------ producer warps
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
                        bulk_g2s(dst + sr * SRP, src + sr * SR4 * ROW, (uint32_t)nrow * ROW * 4u, &full_bar[slot]);
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
                    for (int e = bulk + lane; e < cap; e += 32) dst[sr * SRP + e] = (e < (int)cnt) ? src[s0 + e] : 0.f;
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
                        if (bulk > 0) bulk_g2s(dst + sr * SRP, src + s0, (uint32_t)bulk * 4u, &full_bar[slot]);
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

        // deferred epilogue state: the previous chunk's sums and where they go (n_prev = 0: nothing pending)
        float2 yprev[R];
#pragma unroll
        for (int r = 0; r < R; ++r) yprev[r] = make_float2(0.f, 0.f);
        long long prev_m0 = 0;
        float2* prev_o = p.out;
        int prev_cc = 0;
        long long prev_nout = 0;   // 0 disables the stores

        auto epilogue = [&](const float2* y, int ecc, long long m0, float2* o, long long nout) {
            const float2 rot_chunk = nco_rot(p.phase0_fx + (unsigned long long)ecc * chunk_dph);
            float2 z[R];
#pragma unroll
            for (int r = 0; r < R; ++r) z[r] = cmul(cmul(y[r], rot_thr[r]), rot_chunk);
            if (m0 + R <= nout) {
                if (p.vec_store && (R % 2 == 0)) {
#pragma unroll
                    for (int r = 0; r < R; r += 2)
                        __stcs(reinterpret_cast<float4*>(o + r), make_float4(z[r].x, z[r].y, z[r + 1].x, z[r + 1].y));
                } else {
#pragma unroll
                    for (int r = 0; r < R; ++r) __stcs(o + r, z[r]);
                }
            } else {
#pragma unroll
                for (int r = 0; r < R; ++r)
                    if (m0 + r < nout) __stcs(o + r, z[r]);
            }
        };

        long long t_wait = 0;
        const long long t_begin = clock64();
        for (int k = grp; k < n_k; k += NG) {
            const int slot = sbase + sidx;
            if ((p.debug_mode & 255) != 1) {
                const long long tw0 = p.dbg ? clock64() : 0;
                while (slot_seq[slot] != k) {}
                mbar_wait(&full_bar[slot], par);
                if (p.dbg) t_wait += clock64() - tw0;
            }
            const float* sbuf = buf + (size_t)slot * C::SLOT_FLOATS;
            float2 acc[R];
#pragma unroll
            for (int r = 0; r < R; ++r) acc[r] = make_float2(0.f, 0.f);

            int xoff = 0;
            const float4* tp = &taps.c2[0];
            // ---- phase group 0, with the previous chunk's epilogue in the same basic block
            {
                float4 w[NW];
#pragma unroll
                for (int b = 0; b < NW; ++b)
                    w[b] = *reinterpret_cast<const float4*>(sbuf + rowoff[b / R] + (b % R) * D);
                epilogue(yprev, prev_cc, prev_m0, prev_o, prev_nout);
#pragma unroll
                for (int j = 0; j < JT; ++j) {
                    const float4 ta = tp[j * (D / 2)], tb = tp[j * (D / 2) + 1];
#pragma unroll
                    for (int r = 0; r < R; ++r) acc[r] = ffma2(w[r + j].x, make_float2(ta.x, ta.y), acc[r]);
#pragma unroll
                    for (int r = 0; r < R; ++r) acc[r] = ffma2(w[r + j].y, make_float2(ta.z, ta.w), acc[r]);
#pragma unroll
                    for (int r = 0; r < R; ++r) acc[r] = ffma2(w[r + j].z, make_float2(tb.x, tb.y), acc[r]);
#pragma unroll
                    for (int r = 0; r < R; ++r) acc[r] = ffma2(w[r + j].w, make_float2(tb.z, tb.w), acc[r]);
                }
                xoff = 4;
                tp += 2;
            }
            // ---- remaining phase groups
#pragma unroll 1
            for (int pg = 1; pg < C::V; ++pg, tp += 2) {
                asm volatile("" : "+r"(xoff));
                float4 w[NW];
#pragma unroll
                for (int b = 0; b < NW; ++b)
                    w[b] = *reinterpret_cast<const float4*>(sbuf + xoff + rowoff[b / R] + (b % R) * D);
                xoff += 4;
#pragma unroll
                for (int j = 0; j < JT; ++j) {
                    const float4 ta = tp[j * (D / 2)], tb = tp[j * (D / 2) + 1];
#pragma unroll
                    for (int r = 0; r < R; ++r) acc[r] = ffma2(w[r + j].x, make_float2(ta.x, ta.y), acc[r]);
#pragma unroll
                    for (int r = 0; r < R; ++r) acc[r] = ffma2(w[r + j].y, make_float2(ta.z, ta.w), acc[r]);
#pragma unroll
                    for (int r = 0; r < R; ++r) acc[r] = ffma2(w[r + j].z, make_float2(tb.x, tb.y), acc[r]);
#pragma unroll
                    for (int r = 0; r < R; ++r) acc[r] = ffma2(w[r + j].w, make_float2(tb.z, tb.w), acc[r]);
                }
            }
            __syncwarp();
            if (lane == 0 && (p.debug_mode & 255) != 1) mbar_arrive(&empty_bar[slot]);

            // hand this chunk's sums to the next iteration
#pragma unroll
            for (int r = 0; r < R; ++r) yprev[r] = acc[r];
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
        epilogue(yprev, prev_cc, prev_m0, prev_o, prev_nout);   // the last chunk
        if (p.dbg && lane == 0) {
            atomicAdd(p.dbg, (unsigned long long)t_wait);
            atomicAdd(p.dbg + 1, (unsigned long long)(clock64() - t_begin));
        }
    }
}

}  // namespace ddck
