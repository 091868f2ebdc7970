// Packed 10-bit input on the fast-FIR machinery of ddc_kernel_w.cuh, CUDA cores: the warp-specialised kernel (unpack warps ->
// float ring -> FIR warps).  Since round 2 the default engine for packed input is the tensor-core one (ddc_kernel_tc.cuh); this
// kernel is what option packed_engine = 0 selects at D = 16, 129 .. 256 taps (other cells: unpack stage + float32 kernel).  The
// in-warp-unpack variants of round 1 (fast FIR and direct form, D = 16 / 32 / 64) were measured slower than both and are gone.
#pragma once
#include "ddc_kernel_w.cuh"

namespace ddck {

// ---------------------------------------------------------------------------------------------------------------------
// packed 10-bit input (BASELINE configs[2]; reference stub ddc.py:68-83): raw TMA ring, unpack warps, the fast FIR and the
// deferred branch-free epilogue of ddc_kernel_w.cuh.  The unpack avoids I2F (quarter-rate
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
__device__ __forceinline__ float unpack10_bits(uint32_t field, uint32_t k4b000200) {   // field: the 10 bits in [9:0], anything above
    uint32_t bits;   // (field & 0x3FF) ^ 0x4B000200 as ONE LOP3 (the constant must sit in a register for that)
    asm("lop3.b32 %0, %1, 0x3FF, %2, 0x6A;" : "=r"(bits) : "r"(field), "r"(k4b000200));
    return __uint_as_float(bits);   // = 2^23 + 512 + v; the caller subtracts 8389120 (two samples per FADD2)
}

// ---------------------------------------------------------------------------------------------------------------------
// Packed input, warp-specialised: if every compute warp unpacks its own chunk and then filters it, the FMA pipe idles during
// the latency-bound unpack (round 1: 0.19 of 1.13 ms).  Here FOUR UNPACK WARPS turn raw chunks into a
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
