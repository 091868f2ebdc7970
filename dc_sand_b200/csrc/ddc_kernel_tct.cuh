// Tensor-core engine for packed 10-bit input, second form ("kernel TCT"): the TAP matrix lives in tensor memory as the A
// operand and the samples are the B operand.
//
// Why: the first form (ddc_kernel_tc.cuh) is bound by shared-memory bandwidth -- every tcgen05.mma there re-reads 128 sample
// rows x 32 B plus the tap slice from shared memory (7.2 B per sample at T = 256, D = 16, on top of the unpack's 4.5 B), and
// its MMA warp spends 70 % of the time blocked on a full tensor pipe.  An A operand in TENSOR MEMORY is read for free, so the
// roles are swapped:
//      Y^T[(r, c), row] = sum_k  A[(r, c), k] * X[row, k]       A = part c of tap(k - D r) e^{-j 2 pi step k}   (M = 4 R = 128 rows)
//                                                               X[row, k] = x[ROW_S row + k]                    (N = NR sample rows)
// with R = 32 outputs per sample row (ROW_S = 32 D samples, K = ROW_S - D + T).  A is written once per CTA with tcgen05.st
// (K / 2 columns of packed fp16 pairs next to the two accumulators: 2 NR + K / 2 <= 512 columns bounds T), the sample stream
// is laid out in shared memory exactly as before (NS = ROW_S / 8 sub-streams of 16-byte units, overlapping rows addressed by
// the no-swizzle K-major descriptor), and an MMA now reads only NR x 32 B of shared memory: 2.9 B per sample at T = 256,
// D = 16 for 47 MMAs of M128 x N32 x K16 per 16384-sample tile (16 clocks each on the tensor pipe).  The structured zeros of A
// cost more MACs than in the first form (T / K = 34 % useful), which the tensor pipe has to spare.
//
// Epilogue: the accumulator is transposed (TMEM lane = tap-matrix row, column = sample row).  The tap-matrix rows are ordered
// m = 32 w + 8 c + i for output r = 8 w + i, part c, so that the 16x256b fragment loads of lanes 32 w .. 32 w + 15 (parts
// re_hi, re_lo) and 32 w + 16 .. 32 w + 31 (im_hi, im_lo) hand thread (i, j) = (lane / 4, lane % 4) of warp w all four parts
// of output r for the eight sample rows 8 q + 2 j + e: one rotation per sample row, four FMA-pipe instructions per output, and
// a warp's store covers 8 consecutive outputs (64 bytes = two whole sectors) of four sample rows.
#pragma once
#include "ddc_kernel_tc.cuh"

namespace ddck {

template <int D_, int NR_>
struct TctShape {
    static constexpr int D = D_;
    static constexpr int R = 32;                          // outputs per sample row: M = 4 R rows of the tap matrix
    static constexpr int ROW_S = R * D;                   // samples per sample row
    static constexpr int NS = ROW_S / 8;                  // sub-streams
    static constexpr int LOG_NS = NS == 32 ? 5 : 6;
    static constexpr int NR = NR_;                        // sample rows per tile = N of the MMA
    static constexpr int TILE_S = NR * ROW_S;
    static constexpr int TILE_OUT = NR * R;
    // The packed bytes arrive in CHUNKS of 16384 samples (one raw-ring slot, one unpack batch per lane); a tile is NCHUNK of them
    static constexpr int CHUNK_S = 16384;
    static constexpr int NCHUNK = TILE_S / CHUNK_S;
    static constexpr int CHUNK_PACKED = CHUNK_S / 4 * 5;
    static constexpr int CHUNK_ST = CHUNK_S / 8 / NS * 16;   // bytes a chunk advances inside every sub-stream
    static constexpr int NUNP = DDCB200_TC_NUNP;
    static constexpr int UNP_BATCH = DDCB200_TC_UB;
    static constexpr int NEPI = NR / 32;                  // epilogue warp sets: one per 32-column block of the accumulator
    static constexpr int NTHREADS = (4 + 4 * NEPI + NUNP) * 32;
    static constexpr int ACC_COLS = 2 * NR;               // two accumulators, then the tap matrix
    static constexpr int HDR = 1024;
    static_assert(NS == 32 || NS == 64, "decimation 8 or 16");
    static_assert(NR % 32 == 0 && NR >= 32 && NR <= 128, "sample rows per tile: 32, 64 or 128");
    static_assert(TILE_S % CHUNK_S == 0 && NCHUNK >= 1, "a tile is a whole number of raw chunks");
    static_assert(NUNP % 2 == 0, "the unpack warps pair up over the groups of a row");
};

// A from tensor memory, B from shared memory; issued by ONE elected lane (predicate `issue`)
__device__ __forceinline__ void tct_mma_f16(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, bool accumulate,
                                            bool issue) {
    asm volatile(
        "{\n"
        ".reg .pred p, q;\n"
        ".reg .b64 db;\n"
        "setp.ne.b32 p, %5, 0;\n"
        "setp.ne.b32 q, %6, 0;\n"
        "mov.b64 db, {%2, %3};\n"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"((uint32_t)accumulate), "r"((uint32_t)issue)
        : "memory");
}
__device__ __forceinline__ void tct_st8(uint32_t taddr, const uint4 a, const uint4 b) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(taddr), "r"(a.x), "r"(a.y), "r"(a.z),
                 "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
                 : "memory");
}
__device__ __forceinline__ void tct_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <int D, int NR>
__global__ void __launch_bounds__(TctShape<D, NR>::NTHREADS, 1) ddc_tc10t_kernel(const __grid_constant__ RunParams p,
                                                                                 const __grid_constant__ TcParams tc) {
    using S = TctShape<D, NR>;
    constexpr int R = S::R, NS = S::NS, NU = S::NUNP;
    constexpr int H = NS / 2;                 // MMAs per row shift of the window (one per pair of sub-streams)

    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* raw_full = reinterpret_cast<uint64_t*>(smem);     // [8]
    uint64_t* raw_empty = raw_full + 8;                         // [8]
    uint64_t* a_full = raw_full + 16;                           // [8]   (sample stages; the names follow ddc_kernel_tc.cuh)
    uint64_t* a_empty = raw_full + 24;                          // [8]
    uint64_t* acc_full = raw_full + 32;                         // [2]
    uint64_t* acc_empty = raw_full + 34;                        // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 512);
    unsigned char* asm_ = smem + S::HDR;
    unsigned char* rsm = asm_ + (size_t)tc.n_a * tc.a_stage_bytes;

    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int lane = tid & 31;

    if (tid == 0) {
#pragma unroll 1
        for (int s = 0; s < 8; ++s) {
            mbar_init(&raw_full[s], 1);
            mbar_init(&raw_empty[s], NU);
            mbar_init(&a_full[s], NU);
            mbar_init(&a_empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&acc_full[s], 1);
            mbar_init(&acc_empty[s], 4 * S::NEPI);
        }
        mbar_fence_init();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // Tap matrix -> tensor memory, once per CTA: image [K / 16][128 rows][8 words] (a warp reads 1 KB contiguous per step); warp
    // w can only address lanes 32 (w % 4) .., so the warps of a lane quadrant share the K-steps
    {
        const int q = warp & 3;
        const uint4* img = reinterpret_cast<const uint4*>(tc.b_mat);
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)S::ACC_COLS;
        for (int ch = warp >> 2; ch < tc.k16; ch += (S::NTHREADS / 32) / 4) {
            const uint4* src = img + ((size_t)ch * 128 + q * 32 + lane) * 2;
            tct_st8(lane_addr + (uint32_t)(8 * ch), src[0], src[1]);
        }
        tct_wait_st();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    const int cps = (int)p.tiles_per_stream;
    const int n_k = (int)((p.total_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);
    const int gs = (int)((long long)gridDim.x / cps), gc = (int)((long long)gridDim.x % cps);
    int cs = (int)(blockIdx.x / cps), cc = (int)(blockIdx.x % cps);
    long long tw0 = 0, tw1 = 0;                    // diagnostic: cycles this warp spent in its two waits (option dbg_counters)
    const long long t_begin = p.dbg ? clock64() : 0;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer: one bulk copy per chunk
        int slot = 0;
        uint32_t par = 1;   // first pass: the slots are free
        for (int k = 0; k < n_k; ++k) {
#pragma unroll 1
            for (int hc = 0; hc < S::NCHUNK; ++hc) {
                const long long t0 = p.dbg ? clock64() : 0;
                mbar_wait_uni(&raw_empty[slot], par);
                if (p.dbg) tw0 += clock64() - t0;
                const long long chunk = (long long)cc * S::NCHUNK + hc;
                const unsigned char* src = reinterpret_cast<const unsigned char*>(p.in) + (long long)cs * p.in_stride + chunk * S::CHUNK_PACKED;
                unsigned char* dst = rsm + (size_t)slot * tc.raw_slot_bytes;
                const long long valid_s = p.n_samples - chunk * S::CHUNK_S;   // samples of this stream from the chunk start
                const long long valid_b = valid_s / 4 * 5;
                if (valid_b >= tc.raw_bytes) {
                    if (lane == 0) {
                        mbar_arrive_expect_tx(&raw_full[slot], (uint32_t)tc.raw_bytes);
                        bulk_g2s(dst, src, (uint32_t)tc.raw_bytes, &raw_full[slot]);
                    }
                } else {
                    // ragged end of a stream: whole 16-byte pieces by TMA, the rest by hand, zero bytes (= zero samples) after
                    const int vb = (int)(valid_b > 0 ? valid_b : 0);
                    const int bulk = vb & ~15;
                    for (int e = bulk + lane; e < tc.raw_bytes; e += 32) dst[e] = (e < vb) ? src[e] : (unsigned char)0;
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive_expect_tx(&raw_full[slot], (uint32_t)bulk);
                        if (bulk > 0) bulk_g2s(dst, src, (uint32_t)bulk, &raw_full[slot]);
                    }
                }
                __syncwarp();
                if (++slot == tc.n_raw) { slot = 0; par ^= 1u; }
            }
            cs += gs;
            cc += gc;
            if (cc >= cps) { cc -= cps; ++cs; }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        const bool leader = elect_one();
        const uint32_t idesc = (1u << 4) | ((uint32_t)(NR >> 3) << 17) | (8u << 24);   // f16 x f16 -> f32, M = 128, N = NR, K-major A and B
        const uint32_t b_hi = tc_desc_hi(128u);
        const uint32_t b_lo0 = tc_desc_lo(smem_u32(asm_), (uint32_t)tc.a_pitch);
        const uint32_t pair16 = (uint32_t)(2 * tc.a_pitch) >> 4;      // two sub-streams on, in 16-byte units
        const uint32_t stage16 = (uint32_t)tc.a_stage_bytes >> 4;
        const uint32_t a_tmem0 = tmem_base + (uint32_t)S::ACC_COLS;
        int as = 0, acc = 0;
        uint32_t apar = 0, cpar = 1;
        for (int k = 0; k < n_k; ++k) {
            const long long t0 = p.dbg ? clock64() : 0;
            mbar_wait_uni(&a_full[as], apar);
            const long long t1 = p.dbg ? clock64() : 0;
            mbar_wait_uni(&acc_empty[acc], cpar);
            if (p.dbg) { tw0 += t1 - t0; tw1 += clock64() - t1; }
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * NR);
            uint32_t b_lo = b_lo0 + (uint32_t)as * stage16;   // + one 16-byte row per H MMAs
            uint32_t a_t = a_tmem0;                           // + 8 columns (16 fp16) per MMA
#pragma unroll 1
            for (int i0 = 0; i0 < ((p.debug_mode & 0x100) ? 0 : tc.k16); i0 += 8) {   // debug_mode bit 8: no MMAs (tuning ceiling)
                const uint32_t sub = (uint32_t)(i0 & (H - 1));   // pair of sub-streams inside the row (H is a multiple of 8)
#pragma unroll
                for (int ii = 0; ii < 8; ++ii)
                    tct_mma_f16(d_tmem, a_t + (uint32_t)(8 * ii), b_lo + (sub + (uint32_t)ii) * pair16, b_hi, idesc, (i0 + ii) > 0,
                                leader && (i0 + ii) < tc.k16);
                a_t += 64u;
                if (sub + 8 == (uint32_t)H) b_lo += 1u;
            }
            tc_commit_if(&a_empty[as], leader);      // the sample stage may be overwritten once these MMAs have read it
            tc_commit_if(&acc_full[acc], leader);    // and the accumulator is complete
            __syncwarp();
            if (++as == tc.n_a) { as = 0; apar ^= 1u; }
            acc ^= 1;
            if (acc == 0) cpar ^= 1u;
        }
    } else if (warp >= 4 && warp < 4 + 4 * S::NEPI) {
        // ------------------------------------------------------------------ epilogue warps: one set of four per 32-column block
        // cb of the accumulator (an epilogue warp is ONE serial chain per tile -- wait, tensor-memory loads, stores -- so the sets
        // halve it); warp w of a set owns outputs 8 w .. 8 w + 7 of every sample row, thread (i, j) output 8 w + i of the sample
        // rows 32 cb + 8 q + 2 j + e
        const int w = warp & 3, i = lane >> 2, j = lane & 3, cb = (warp >> 2) - 1;
        const unsigned long long row_dph = (unsigned long long)S::ROW_S * p.step_fx;
        const unsigned long long tile_dph = (unsigned long long)S::TILE_S * p.step_fx;
        // NCO: the tap matrix carries the rotation by the sample's position INSIDE its row (k_tc.cu: build_a), so all outputs of
        // a sample row take the residual rotation e^{-j 2 pi step (first sample of the row)}: one polynomial rotation per thread
        // and tile, one complex multiplication per sample row by a constant that also carries the 512 / S of the unpack and taps
        float2 rowrot[8];
#pragma unroll
        for (int qe = 0; qe < 8; ++qe) {
            const float2 t = nco_rot((unsigned long long)(32 * cb + 8 * (qe >> 1) + 2 * j + (qe & 1)) * row_dph);
            rowrot[qe] = make_float2(t.x * tc.inv_scale, t.y * tc.inv_scale);
        }
        const float lo_w = tc.lo_scale / tc.inv_scale;   // 2^-11
        int acc = 0;
        uint32_t fpar = 0;
        for (int k = 0; k < n_k; ++k) {
            // everything that does not depend on the accumulator comes BEFORE the wait: this warp is one serial chain per tile
            const float2 rot0 = nco_rot_bf(p.phase0_fx + (unsigned long long)cc * tile_dph);
            float2* o = p.out + (long long)cs * p.out_stride + (long long)cc * S::TILE_OUT + 8 * w + i;
            const long long n_left = p.n_out - (long long)cc * S::TILE_OUT - (8 * w + i);   // my output of sample row n exists if n R < n_left
            float2 t[8];
#pragma unroll
            for (int qe = 0; qe < 8; ++qe) t[qe] = cmul(rot0, rowrot[qe]);
            const long long t0 = p.dbg ? clock64() : 0;
            mbar_wait_uni(&acc_full[acc], fpar);
            if (p.dbg) tw0 += clock64() - t0;
            tc_fence_after();
            {
                // lanes 32 w .. hold re_hi (rows i) and re_lo (rows i + 8), lanes 32 w + 16 .. im_hi and im_lo
                const uint32_t taddr = tmem_base + ((uint32_t)(w * 32) << 16) + (uint32_t)(acc * NR + 32 * cb);
                uint32_t v0[16], v1[16];
                tc_ld_16x256_x4(taddr, v0);
                tc_ld_16x256_x4(taddr + (16u << 16), v1);
                tc_wait_ld();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_empty[acc]);   // my part of the accumulator is back with the MMA warp
                if (!(p.debug_mode & 0x400)) {                 // debug_mode bit 10: no epilogue arithmetic or stores (tuning ceiling)
#pragma unroll
                    for (int qe = 0; qe < 8; ++qe) {
                        const int q = qe >> 1, e = qe & 1;
                        const float re = fmaf(__uint_as_float(v0[4 * q + 2 + e]), lo_w, __uint_as_float(v0[4 * q + e]));
                        const float im = fmaf(__uint_as_float(v1[4 * q + 2 + e]), lo_w, __uint_as_float(v1[4 * q + e]));
                        const float2 z = ffma2(im, make_float2(-t[qe].y, t[qe].x), make_float2(re * t[qe].x, re * t[qe].y));
                        const long long m = (long long)(32 * cb + 8 * q + 2 * j + e) * R;
                        const bool ok = (p.debug_mode & 0x800) ? (z.x == 1.2345f) : (m < n_left);   // debug_mode bit 11: (almost) no stores
                        st_cs_v2_if(o + m, z.x, z.y, ok);
                    }
                }
            }
            acc ^= 1;
            if (acc == 0) fpar ^= 1u;
            cs += gs;
            cc += gc;
            if (cc >= cps) { cc -= cps; ++cs; }
        }
    } else if (warp >= 4 + 4 * S::NEPI) {
        // ------------------------------------------------------------------ unpack warps: all of them share every tile
        const int u = warp - (4 + 4 * S::NEPI);
        // lane -> 16-sample group.  A quarter warp's two STS.128 must hit eight distinct 16-byte bank groups: four even
        // sub-streams of one row and the same four of the next row (sub-stream pitch = odd number of units), so lane bit 2
        // selects the row (group bit LOG_NS - 1); lane bits 0, 1, 3, 4 are group bits 0 .. 3; a group bit between them (rows of
        // 32 groups) comes from the low bit of the warp index, the rest from its high bits and the batch.
        constexpr int LG = S::LOG_NS - 1;   // log2(groups per row)
        constexpr int XB = LG - 4;          // group bits 4 .. LG - 1 taken from the warp index
        static_assert(LG >= 4 && XB <= 1, "rows of 16 or 32 groups");
        const int low = lane & 3, rsel = (lane >> 2) & 1, rest = lane >> 3;
        // (rows of 32 groups: bit 4 is flipped in the lanes of the second row, so that the 32 lanes' groups stay distinct modulo
        // 32 and their 20-byte-strided raw loads conflict-free; the stores only need the row parity to differ)
        const int gfirst = XB == 0 ? (low | (rest << 2) | (rsel << 4) | (u << 5))
                                   : (low | (rest << 2) | (((u & 1) ^ rsel) << 4) | (rsel << 5) | ((u >> 1) << 6));
        constexpr int UB = S::UNP_BATCH;    // groups a lane has in flight: all loads first, then the integer work, then the stores
        // a lane's groups are GSTEP apart, so both its raw address (20 bytes per group) and its destination (2 * GSTEP units on =
        // the same sub-stream, 2 * GSTEP / NS rows down) advance by compile-time constants
        constexpr int GSTEP = 32 * NU, LD_STEP = 20 * GSTEP, ST_STEP = 2 * GSTEP / NS * 16;
        static_assert((2 * GSTEP) % NS == 0, "a lane must stay on one sub-stream pair");
        const uint32_t ld_off = 20u * (uint32_t)gfirst;
        const uint32_t st_off = (uint32_t)((2 * gfirst) & (NS - 1)) * (uint32_t)tc.a_pitch + (uint32_t)((2 * gfirst) >> S::LOG_NS) * 16u;
        const bool on = !(p.debug_mode & 0x200);   // debug_mode bit 9: no unpack (tuning ceiling)
        bool valid[UB];
#pragma unroll
        for (int b = 0; b < UB; ++b) valid[b] = on && gfirst + b * GSTEP < tc.n_groups;
        uint32_t rw[UB][5];
        auto load_raw = [&](int slot) {
            const unsigned char* rp = rsm + (size_t)slot * tc.raw_slot_bytes + ld_off;
#pragma unroll
            for (int b = 0; b < UB; ++b) {
                const uint32_t* src = reinterpret_cast<const uint32_t*>(rp + b * LD_STEP);
#pragma unroll
                for (int ii = 0; ii < 5; ++ii) rw[b][ii] = valid[b] ? src[ii] : 0u;
            }
        };
        int rs = 0, as = 0;
        uint32_t rpar = 0, epar = 1;
        if (n_k > 0) {
            mbar_wait_uni(&raw_full[0], 0);
            load_raw(0);
        }
        for (int k = 0; k < n_k; ++k) {
            const long long t0 = p.dbg ? clock64() : 0;
            mbar_wait_uni(&a_empty[as], epar);
            const long long t2 = p.dbg ? clock64() : 0;
            if (p.dbg) tw0 += t2 - t0;
#pragma unroll
            for (int hc = 0; hc < S::NCHUNK; ++hc) {
                // chunk hc of the tile: its groups start hc * CHUNK_ST bytes down every sub-stream (the 15-odd halo groups of a
                // chunk are the first groups of the next one: written twice with the same values)
                unsigned char* sp = asm_ + (size_t)as * tc.a_stage_bytes + st_off + hc * S::CHUNK_ST;
#pragma unroll
                for (int b = 0; b < UB; ++b) {
                    uint32_t h[8];
                    tc_unpack16(rw[b], h, tc);
                    unsigned char* dst = sp + b * ST_STEP;   // units 2g (even sub-stream) and 2g + 1 (the next sub-stream, same row)
                    if (valid[b]) {
                        *reinterpret_cast<uint4*>(dst) = make_uint4(h[0], h[1], h[2], h[3]);
                        *reinterpret_cast<uint4*>(dst + tc.a_pitch) = make_uint4(h[4], h[5], h[6], h[7]);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&raw_empty[rs]);   // every lane's raw words have been consumed
                if (++rs == tc.n_raw) { rs = 0; rpar ^= 1u; }
                // next chunk's raw words: now if they have landed (the usual case), else after the hand-over of this stage
                const bool more = hc + 1 < S::NCHUNK || k + 1 < n_k;
                const bool early = more && mbar_test_uni(&raw_full[rs], rpar);
                if (early) load_raw(rs);
                if (hc == S::NCHUNK - 1) {
                    fence_proxy_async();   // my stores before the tensor core's reads of this stage
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&a_full[as]);
                }
                if (more && !early) {
                    const long long t3 = p.dbg ? clock64() : 0;
                    mbar_wait_uni(&raw_full[rs], rpar);
                    if (p.dbg) tw0 += clock64() - t3;
                    load_raw(rs);
                }
            }
            if (p.dbg) tw1 += clock64() - t2;
            if (++as == tc.n_a) { as = 0; epar ^= 1u; }
        }
    }

    if (p.dbg && lane == 0 && (warp == 0 || warp == 1 || warp == 4 || warp == 4 + 4 * S::NEPI)) {
        const int role = warp == 0 ? 0 : (warp == 1 ? 1 : (warp == 4 ? 2 : 3));
        atomicAdd(p.dbg + 2 + 3 * role, (unsigned long long)tw0);
        atomicAdd(p.dbg + 3 + 3 * role, (unsigned long long)tw1);
        atomicAdd(p.dbg + 4 + 3 * role, (unsigned long long)(clock64() - t_begin));
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    }
}

}  // namespace ddck
