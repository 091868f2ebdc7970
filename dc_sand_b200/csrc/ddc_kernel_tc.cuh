// Tensor-core engine for PACKED 10-bit input ("kernel TC"): the default for packed input (option "packed_engine" = 1; 0 keeps
// the CUDA-core kernels of ddc_kernel_w10.cuh / _p.cuh; option "variant" = 13 forces this engine).
//
// Why it exists: packed input is FP32-bound by 3.2x on the CUDA cores (1.75 B against 64 flop per sample at T = 256, D = 16), so
// the fused-unpack fast-FIR kernel sits at 26 % of the HBM roofline however well it is scheduled.  10-bit samples are EXACT in
// fp16, and a decimating FIR over a tile of outputs is a Toeplitz product, so for this input type -- and only for it -- the
// filter can run on tcgen05 without giving up the reference's precision: taps are split hi + lo into two fp16 values
// (22 significant bits, the float32 taps of the CUDA-core kernels have 24), products are exact, accumulation is FP32 in TMEM.
// Measured against the float64 oracle it is as close as the CUDA-core kernels (max error 7e-7 of full scale at T = 256, 2e-6
// at T = 1024; tests/test_gpu_tensor_engine.py).  The float32-input kernels stay on the CUDA cores (north star): a float32
// sample is not exact in any tensor-core input type.
//
// Formulation.  One MMA row covers ROW_S = 8 NS consecutive input samples (NS = 16: 128 samples) = R = ROW_S / D outputs:
//      Y[row, (r, c)] = sum_k  X[row, k] * B[(r, c), k],     X[row, k] = x[ROW_S row + k],   k = 0 .. K-1,  K = ROW_S - D + T (padded to 16)
//      B[(r, c), k]   = part c of  tap(k - D r) e^{-j 2 pi step k}     (c: re_hi, re_lo, im_hi, im_lo;  0 outside the tap window)
// i.e. M = 128 rows x N = 4 R columns x K per tile; the structured zeros of B cost (R - 1) D / T extra MACs and make N a legal
// tcgen05 shape.  The NCO rotation inside B goes by the sample's position in the ROW, which leaves every output of a row
// with the same residual rotation for the epilogue.  X is never materialised: the unpack warps write the fp16 sample stream
// into shared memory as NS SUB-STREAMS of 16-byte units (unit u = samples 8u .. 8u+7 goes to sub-stream u % NS, row u / NS),
// and because consecutive rows of X start exactly one row further on in every sub-stream, the canonical no-swizzle K-major
// operand layout of tcgen05 (8-row core matrices of 16-byte rows, SBO = 128 bytes between row groups, LBO = the distance
// between two sub-streams for the two K-units of one MMA) describes the overlapping windows directly: K-unit q of row m is
// unit NS m + q = sub-stream q % NS, row m + q / NS.  An MMA for K-units (2i, 2i+1) just starts (2i) / NS rows down
// sub-stream (2i) % NS.
//
// Pipeline per CTA (one per SM, persistent, 8 + NUNP warps):
//      warp 0      TMA producer: one bulk copy of packed bytes per tile -> raw ring (up to 8 slots: HBM latency)
//      warps 8..   unpack, in NTEAM teams that take the tiles in turn: 20 bytes -> 16 fp16 per lane and group, bit-exact (PRMT,
//                  LOP3, IMAD.WIDE, LEA.HI, HFMA2 per sample pair), raw words of the team's NEXT tile already in registers,
//                  two STS.128 per group into the sub-streams
//      warp 1      one elected thread issues K / 16 tcgen05.mma per tile into one of two TMEM accumulators
//      warps 4-7   epilogue: tcgen05.ld in the 16x256b fragment layout (a quad of lanes holds one row's outputs, so the warp's
//                  stores cover whole sectors), hi + lo recombination and NCO rotation in four FFMA2 per output, streaming stores
// Shared-memory bandwidth bounds it (operand re-reads of the MMA: (128 + N) * 32 B per K-step, + unpack traffic), not the
// tensor pipe and not the CUDA cores: DESIGN.md section 4.7.  SASS evidence: UBLKCP, UTCHMMA, LDTM (profiles/r2_sass_mnemonics.txt).
#pragma once
#include <cuda_fp16.h>

#include "ddc_common.cuh"

#ifndef DDCB200_TC_NUNP
#define DDCB200_TC_NUNP 12
#endif
#ifndef DDCB200_TC_UB
#define DDCB200_TC_UB 3
#endif
// The unpack warps work in NTEAM teams that take the tiles in turn (template parameter of the kernel): three teams of four
// warps where the pipeline has three or more sample stages, two teams of six where it has two -- a team must never wait for
// a stage two phases ahead of the barrier (parity waits alias), so NTEAM <= stages.

// Per-tile event trace of CTA 0 (build with -DDDCB200_TC_TRACE, option dbg_counters = 3 prints it): clock64 of ten events
// of the first 96 tiles, at p.dbg[64 + 96 * event + tile] -- how the hand-shakes of the five warp roles line up in time
// (profiles/r2_tensor_engine_trace.md).
#ifdef DDCB200_TC_TRACE
#define TC_TRACE(ev, k) \
    do { if (p.dbg && blockIdx.x == 0 && (k) < 96 && lane == 0) p.dbg[64 + 96 * (ev) + (k)] = (unsigned long long)clock64(); } while (0)
#else
#define TC_TRACE(ev, k) do { } while (0)
#endif

namespace ddck {

struct TcParams {
    const void* b_mat;        // fp16 B operand in its shared-memory image: [K / 8][N][8] halves
    int k16;                  // K / 16: MMAs per tile
    int b_bytes;              // N * K * 2
    int a_rows;               // rows of a sub-stream
    int a_pitch;              // L: bytes between sub-streams (16 * odd)
    int a_stage_bytes;        // 8 * L rounded to 128
    int raw_bytes;            // packed bytes copied per tile (multiple of 16)
    int raw_slot_bytes;       // raw_bytes rounded to 128
    int n_groups;             // 16-sample groups unpacked per tile
    int n_a;                  // A stages
    int n_raw;                // raw slots
    float inv_scale;          // 512 / S  (the unpack delivers v / 512)
    float lo_scale;           // 512 * 2^-11 / S
    uint32_t unp_mul[2];      // 2^10, 2^14: multipliers of the unpack (kept in the parameter block on purpose)
};

template <int NS_>
struct TcShape {
    static constexpr int NS = NS_;                        // sub-streams = 16-byte units per MMA row
    static constexpr int TILE_ROWS = 128;
    static constexpr int ROW_S = 8 * NS;                  // samples per MMA row
    static constexpr int TILE_S = TILE_ROWS * ROW_S;      // samples per tile
    static constexpr int TILE_PACKED = TILE_S / 4 * 5;    // packed bytes per tile
    static constexpr int NUNP = DDCB200_TC_NUNP;          // unpack warps
    static constexpr int UNP_BATCH = DDCB200_TC_UB;       // 16-sample groups a lane unpacks per batch
    static constexpr int UNP_CAP = DDCB200_TC_UB * 32 * DDCB200_TC_NUNP;   // groups per tile the unpack warps cover, with any number of teams
    static constexpr int NTHREADS = (8 + NUNP) * 32;
    static constexpr int HDR = 1024;
    static constexpr int LOG_NS = NS == 8 ? 3 : (NS == 16 ? 4 : 5);
    static_assert(NS == 8 || NS == 16 || NS == 32, "8, 16 or 32 sub-streams");
    __host__ __device__ static constexpr int n_cols(int D) { return 4 * ROW_S / D < 16 ? 16 : 4 * ROW_S / D; }
    __host__ __device__ static constexpr int tmem_cols(int D) {
        return 2 * n_cols(D) <= 32 ? 32 : (2 * n_cols(D) <= 64 ? 64 : (2 * n_cols(D) <= 128 ? 128 : (2 * n_cols(D) <= 256 ? 256 : 512)));
    }
};

// descriptor words: lo = start address | leading (K-unit) byte offset, hi = stride (8-row group) byte offset | version 1,
// all in 16-byte units; no swizzle, K-major
__device__ __forceinline__ uint32_t tc_desc_lo(uint32_t addr, uint32_t lbo_bytes) { return ((addr >> 4) & 0x3FFFu) | (((lbo_bytes >> 4) & 0x3FFFu) << 16); }
__device__ __forceinline__ constexpr uint32_t tc_desc_hi(uint32_t sbo_bytes) { return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14); }

// issued by ONE elected lane (predicate `issue`); the other operands are warp-uniform
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                           bool accumulate, bool issue) {
    asm volatile(
        "{\n"
        ".reg .pred p, q;\n"
        ".reg .b64 da, db;\n"
        "setp.ne.b32 p, %6, 0;\n"
        "setp.ne.b32 q, %7, 0;\n"
        "mov.b64 da, {%1, %2};\n"
        "mov.b64 db, {%3, %4};\n"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n"
        "}\n" ::"r"(tmem_d),
        "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"((uint32_t)accumulate), "r"((uint32_t)issue)
        : "memory");
}
__device__ __forceinline__ void tc_commit_if(uint64_t* bar, bool issue) {
    asm volatile(
        "{\n"
        ".reg .pred q;\n"
        "setp.ne.b32 q, %1, 0;\n"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"((uint32_t)issue)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
// 16 lanes x 256 bits per block of eight columns, like an MMA accumulator fragment: lane t of the warp receives, for block q,
// (row t / 4, columns 8 q + 2 (t % 4) + {0, 1}) in v[4 q + {0, 1}] and (row t / 4 + 8, same columns) in v[4 q + {2, 3}]
__device__ __forceinline__ void tc_ld_16x256_x4(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_ld_16x256_x2(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}
// Accumulator column of part c4 (0 re_hi, 1 re_lo, 2 im_hi, 3 im_lo) of output r of a row with R outputs.  With R >= 4 the
// columns are ordered for the 16x256b fragment load: lane t % 4 = j then owns ALL FOUR parts of outputs 2 j and 2 j + 1 of a
// row (of output j when R = 4), so a quad of lanes writes 64 (32) contiguous bytes and the eight quads of a warp store eight
// consecutive rows -- whole sectors, where the row-per-lane 32x32b load gives 16-byte pieces 64 bytes apart.
__host__ __device__ constexpr int tc_col(int R, int r, int c4) {
    return R >= 8 ? 32 * (r / 8) + 8 * (2 * (r % 2) + c4 / 2) + 2 * ((r % 8) / 2) + (c4 % 2)
                  : (R == 4 ? 8 * (c4 / 2) + 2 * r + (c4 % 2) : 4 * r + c4);
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 16 packed samples (five little-endian words of the big-endian bit stream: sample s = bits [10 s, 10 s + 10) from the top) ->
// eight fp16 pairs holding v / 512, bit-exact.  Per pair q (samples 2q, 2q+1 = 20 bits from bit 20 q, i.e. three bytes from byte
// 5q / 2): one PRMT gathers the three bytes, most significant first, from the little-endian words; one LOP3 flips the two sign
// bits (offset binary u = v + 512) and isolates the pair; one 32 x 32 -> 64 bit multiply by a power of two (IMAD.WIDE on the FMA
// pipe; the multipliers come from the parameter block so that ptxas keeps the multiply instead of ALU-pipe shifts) leaves u_a in
// the high word and u_b in the top ten bits of the low word; one shift-add (LEA.HI) drops u_b into the high half.  Each half
// now holds u as an fp16 DENORMAL (u 2^-24), and one HFMA2 finishes: u 2^-24 * 2^15 - 1 = (u - 512) / 512 = v / 512, exact.
// 4 instructions per pair, 2 on the ALU pipe and 2 on the FMA pipe; the factor 512 goes into the epilogue's scales.
__device__ __forceinline__ void tc_unpack16(const uint32_t (&rw)[5], uint32_t (&h)[8], const TcParams& tc) {
    const __half2 k15 = __floats2half2_rn(32768.f, 32768.f), m1 = __floats2half2_rn(-1.f, -1.f);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int bo = (20 * q) >> 3, wi = bo >> 2, j0 = bo & 3, s2 = (20 * q) & 7;   // s2 = 0 or 4
        const int j2 = j0 + 2 > 7 ? 7 : j0 + 2;
        const uint32_t sel = (uint32_t)((j0 << 12) | ((j0 + 1) << 8) | (j2 << 4) | j2);
        const uint32_t src = __byte_perm(rw[wi], rw[wi + 1 > 4 ? 4 : wi + 1], sel);
        uint32_t fm;
        asm("lop3.b32 %0, %1, %2, %3, 0x28;" : "=r"(fm) : "r"(src), "r"(0x80200000u >> s2), "r"(0xFFFFF000u >> s2));   // (src ^ x) & m
        uint32_t plo, phi;
        asm("{\n.reg .b64 t;\nmul.wide.u32 t, %2, %3;\nmov.b64 {%0, %1}, t;\n}" : "=r"(plo), "=r"(phi) : "r"(fm), "r"(tc.unp_mul[s2 >> 2]));
        const uint32_t w = phi + (plo >> 6);
        const __half2 hv = __hfma2(*reinterpret_cast<const __half2*>(&w), k15, m1);
        h[q] = *reinterpret_cast<const uint32_t*>(&hv);
    }
}

template <int D, int NS, int NTEAM>
__global__ void __launch_bounds__(TcShape<NS>::NTHREADS, 1) ddc_tc10_kernel(const __grid_constant__ RunParams p,
                                                                            const __grid_constant__ TcParams tc) {
    using S = TcShape<NS>;
    constexpr int R = S::ROW_S / D;           // outputs per MMA row
    constexpr int N = S::n_cols(D);           // accumulator columns: 4 per output, at least 16
    constexpr int TILE_OUT = S::TILE_ROWS * R;
    constexpr int NU = S::NUNP;
    constexpr int H = NS / 2;                 // MMAs per row shift of the window (one per pair of sub-streams)
    static_assert(D == 4 || D == 8 || D == 16 || D == 32 || D == 64, "decimation must divide 64");
    static_assert(N <= 256, "accumulator too wide");

    extern __shared__ __align__(1024) unsigned char smem[];
    uint64_t* raw_full = reinterpret_cast<uint64_t*>(smem);     // [8]
    uint64_t* raw_empty = raw_full + 8;                         // [8]
    uint64_t* a_full = raw_full + 16;                           // [8]
    uint64_t* a_empty = raw_full + 24;                          // [8]
    uint64_t* acc_full = raw_full + 32;                         // [2]
    uint64_t* acc_empty = raw_full + 34;                        // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 512);
    unsigned char* bsm = smem + S::HDR;
    unsigned char* asm_ = bsm + ((tc.b_bytes + 127) & ~127);
    unsigned char* rsm = asm_ + (size_t)tc.n_a * tc.a_stage_bytes;

    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int lane = tid & 31;

    if (tid == 0) {
#pragma unroll 1
        for (int s = 0; s < 8; ++s) {
            mbar_init(&raw_full[s], 1);
            mbar_init(&raw_empty[s], NU / NTEAM);
            mbar_init(&a_full[s], NU / NTEAM);
            mbar_init(&a_empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&acc_full[s], 1);
            mbar_init(&acc_empty[s], 4);
        }
        mbar_fence_init();
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"((uint32_t)S::tmem_cols(D))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // B operand: the same image for every tile, copied once per CTA (L2 hits after the first CTA)
    {
        const uint4* src = reinterpret_cast<const uint4*>(tc.b_mat);
        uint4* dst = reinterpret_cast<uint4*>(bsm);
        for (int i = tid; i < tc.b_bytes / 16; i += S::NTHREADS) dst[i] = src[i];
    }
    fence_proxy_async();   // generic-proxy writes of B before the tensor core (async proxy) reads them
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int cps = (int)p.tiles_per_stream;
    const int n_k = (int)((p.total_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);
    const int gs = (int)((long long)gridDim.x / cps), gc = (int)((long long)gridDim.x % cps);
    int cs = (int)(blockIdx.x / cps), cc = (int)(blockIdx.x % cps);
    long long tw0 = 0, tw1 = 0;                    // diagnostic: cycles this warp spent in its two waits (option dbg_counters)
    const long long t_begin = p.dbg ? clock64() : 0;

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer: one bulk copy per tile
        int slot = 0;
        uint32_t par = 1;   // first pass: the slots are free
        for (int k = 0; k < n_k; ++k) {
            const long long t0 = p.dbg ? clock64() : 0;
            mbar_wait_uni(&raw_empty[slot], par);
            if (p.dbg) tw0 += clock64() - t0;
            TC_TRACE(0, k);
            const unsigned char* src = reinterpret_cast<const unsigned char*>(p.in) + (long long)cs * p.in_stride + (long long)cc * S::TILE_PACKED;
            unsigned char* dst = rsm + (size_t)slot * tc.raw_slot_bytes;
            const long long valid_s = p.n_samples - (long long)cc * S::TILE_S;   // samples of this stream from the tile start
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
            cs += gs;
            cc += gc;
            if (cc >= cps) { cc -= cps; ++cs; }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer: the whole warp runs the (warp-uniform)
        // descriptor arithmetic, which the compiler keeps on the uniform datapath; one elected lane issues
        const bool leader = elect_one();
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(S::TILE_ROWS >> 4) << 24);   // f16 x f16 -> f32, K-major A and B
        const uint32_t a_hi = tc_desc_hi(128u), b_hi = tc_desc_hi(128u);
        const uint32_t b_lo0 = tc_desc_lo(smem_u32(bsm), (uint32_t)(N * 16));
        const uint32_t a_lo0 = tc_desc_lo(smem_u32(asm_), (uint32_t)tc.a_pitch);
        const uint32_t pair16 = (uint32_t)(2 * tc.a_pitch) >> 4;      // two sub-streams on, in 16-byte units
        const uint32_t stage16 = (uint32_t)tc.a_stage_bytes >> 4;
        constexpr uint32_t BSTEP = (uint32_t)(2 * N * 16) >> 4;       // two K-units of B
        int as = 0, acc = 0;
        uint32_t apar = 0, cpar = 1;
        for (int k = 0; k < n_k; ++k) {
            const long long t0 = p.dbg ? clock64() : 0;
            mbar_wait_uni(&a_full[as], apar);
            const long long t1 = p.dbg ? clock64() : 0;
            mbar_wait_uni(&acc_empty[acc], cpar);
            if (p.dbg) { tw0 += t1 - t0; tw1 += clock64() - t1; }
            TC_TRACE(4, k);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(acc * N);
            uint32_t a_lo = a_lo0 + (uint32_t)as * stage16;   // + one 16-byte row per H MMAs
            uint32_t b_lo = b_lo0;
#pragma unroll 1
            for (int i0 = 0; i0 < ((p.debug_mode & 255) == 1 ? 0 : tc.k16); i0 += H) {   // debug_mode 1: no MMAs (tuning ceiling)
#pragma unroll
                for (int ii = 0; ii < H; ++ii)
                    tc_mma_f16(d_tmem, a_lo + (uint32_t)ii * pair16, a_hi, b_lo + (uint32_t)ii * BSTEP, b_hi, idesc, (i0 + ii) > 0,
                               leader && (i0 + ii) < tc.k16);
                a_lo += 1u;
                b_lo += (uint32_t)H * BSTEP;
            }
            tc_commit_if(&a_empty[as], leader);      // the A stage may be overwritten once these MMAs have read it
            tc_commit_if(&acc_full[acc], leader);    // and the accumulator is complete
            TC_TRACE(5, k);
            __syncwarp();
            if (++as == tc.n_a) { as = 0; apar ^= 1u; }
            acc ^= 1;
            if (acc == 0) cpar ^= 1u;
        }
    } else if (warp >= 4 && warp < 8 && R >= 4) {
        // ------------------------------------------------------------------ epilogue warps (four or more outputs per row):
        // fragment loads, lane (i, j) = (lane / 4, lane % 4) handles rows quad * 32 + i + 8 h, h = 0 .. 3, outputs 2 j, 2 j + 1
        // of every 32-column block (output j when R = 4).
        // NCO: the tap matrix carries the rotation by the sample's position INSIDE the row (k_tc.cu: build_b), so every output
        // of a row takes the same residual rotation e^{-j 2 pi step (first sample of the row)}: one polynomial rotation (64-bit
        // fixed-point phase) per thread and tile, three angle-addition steps for its other rows, then four FFMA2 per output
        // that also recombine the hi and lo tap parts:  z = (re_hi + j im_hi) a rot + (re_lo + j im_lo) b rot
        const int quad = warp & 3, i = lane >> 2, j = lane & 3;
        const unsigned long long row_dph = (unsigned long long)S::ROW_S * p.step_fx;
        const unsigned long long tile_dph = (unsigned long long)S::TILE_S * p.step_fx;
        const int row0 = quad * 32 + i;
        const unsigned long long row_ph = p.phase0_fx + (unsigned long long)row0 * row_dph;
        float2 cstep[3];
#pragma unroll
        for (int h = 1; h < 4; ++h) cstep[h - 1] = nco_rot((unsigned long long)(8 * h) * row_dph);
        constexpr int NB = R >= 8 ? R / 8 : 1;      // 32-column blocks (one 16-column block when R = 4)
        constexpr int OB = R >= 8 ? 8 : 4;          // outputs of a row per block
        int acc = 0;
        uint32_t fpar = 0;
        for (int k = 0; k < n_k; ++k) {
            const long long t0 = p.dbg ? clock64() : 0;
            mbar_wait_uni(&acc_full[acc], fpar);
            if (p.dbg) tw0 += clock64() - t0;
            if (warp == 4) TC_TRACE(6, k);
            tc_fence_after();
            const float2 rot0 = nco_rot_bf(row_ph + (unsigned long long)cc * tile_dph);
            float2* o = p.out + (long long)cs * p.out_stride + (long long)cc * TILE_OUT;
            const long long n_left = p.n_out - (long long)cc * TILE_OUT;   // outputs of this stream from the tile start on
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
#pragma unroll
                for (int blk = 0; blk < NB; ++blk) {
                    uint32_t v[16];
                    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32 + 16 * hf) << 16) + (uint32_t)(acc * N + 32 * blk);
                    if (R >= 8) tc_ld_16x256_x4(taddr, v);
                    else tc_ld_16x256_x2(taddr, v);
                    tc_wait_ld();
                    if (hf == 1 && blk == NB - 1) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&acc_empty[acc]);   // accumulator back to the MMA warp
                        if (warp == 4) TC_TRACE(7, k);
                    }
                    if ((p.debug_mode & 255) == 3) continue;   // debug_mode 3: no epilogue arithmetic or stores (tuning ceiling)
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        const int h = 2 * hf + hh;
                        const float2 rot = h == 0 ? rot0 : cmul(rot0, cstep[h > 0 ? h - 1 : 0]);
                        const float2 t_hi_re = make_float2(rot.x * tc.inv_scale, rot.y * tc.inv_scale);     // times a real part
                        const float2 t_hi_im = make_float2(-t_hi_re.y, t_hi_re.x);                          // times an imaginary part
                        const float2 t_lo_re = make_float2(rot.x * tc.lo_scale, rot.y * tc.lo_scale);
                        const float2 t_lo_im = make_float2(-t_lo_re.y, t_lo_re.x);
                        constexpr int NO = R >= 8 ? 2 : 1;   // outputs of this lane in the block
                        float2 z[NO];
#pragma unroll
                        for (int oo = 0; oo < NO; ++oo) {
                            const float re_hi = __uint_as_float(v[8 * oo + 2 * hh]), re_lo = __uint_as_float(v[8 * oo + 2 * hh + 1]);
                            const float im_hi = __uint_as_float(v[8 * oo + 4 + 2 * hh]), im_lo = __uint_as_float(v[8 * oo + 4 + 2 * hh + 1]);
                            float2 zz = make_float2(re_lo * t_lo_re.x, re_lo * t_lo_re.y);
                            zz = ffma2(im_lo, t_lo_im, zz);
                            zz = ffma2(re_hi, t_hi_re, zz);
                            z[oo] = ffma2(im_hi, t_hi_im, zz);
                        }
                        const long long m = (long long)(row0 + 8 * h) * R + OB * blk + NO * j;   // first output of this lane
                        const long long l = (p.debug_mode & 255) == 4 ? (long long)(z[0].x == 1.2345f) : n_left - m;   // debug_mode 4: (almost) no stores
                        if (NO == 2 && p.vec_store) {
                            st_cs_v4_if(o + m, z[0].x, z[0].y, z[NO - 1].x, z[NO - 1].y, l >= 2);
                            st_cs_v2_if(o + m, z[0].x, z[0].y, l == 1);
                        } else {
#pragma unroll
                            for (int oo = 0; oo < NO; ++oo) st_cs_v2_if(o + m + oo, z[oo].x, z[oo].y, l >= oo + 1);
                        }
                    }
                }
            }
            if (warp == 4) TC_TRACE(8, k);
            acc ^= 1;
            if (acc == 0) fpar ^= 1u;
            cs += gs;
            cc += gc;
            if (cc >= cps) { cc -= cps; ++cs; }
        }
    } else if (warp >= 4 && warp < 8) {
        // ------------------------------------------------------------------ epilogue warps (fewer than four outputs per row):
        // one row per lane, TMEM lanes 32 (warp % 4) ..
        const int row = (warp & 3) * 32 + lane;
        const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
        const unsigned long long row_dph = (unsigned long long)S::ROW_S * p.step_fx;
        const unsigned long long tile_dph = (unsigned long long)S::TILE_S * p.step_fx;
        const unsigned long long row_ph = p.phase0_fx + (unsigned long long)row * row_dph;
        constexpr int RC = R < 4 ? R : 4;           // outputs per 16-column piece
        // NCO: the tap matrix carries the rotation by the sample's position INSIDE the row (k_tc.cu: build_b), so every output
        // of a row takes the same residual rotation e^{-j 2 pi step (first sample of the row)}: one polynomial rotation (64-bit
        // fixed-point phase) per row and tile, then four FFMA2 per output that also recombine the hi and lo tap parts:
        //      z = (re_hi + j im_hi) a rot + (re_lo + j im_lo) b rot,      a = 512 / S, b = a 2^-11
        int acc = 0;
        uint32_t fpar = 0;
        for (int k = 0; k < n_k; ++k) {
            const long long t0 = p.dbg ? clock64() : 0;
            mbar_wait_uni(&acc_full[acc], fpar);
            if (p.dbg) tw0 += clock64() - t0;
            tc_fence_after();
            const long long m_row = (long long)cc * TILE_OUT + (long long)row * R;
            float2* o = p.out + (long long)cs * p.out_stride + m_row;
            const long long left = p.n_out - m_row;     // outputs of this row that exist (ragged stream tail)
            const float2 rot = nco_rot_bf(row_ph + (unsigned long long)cc * tile_dph);
            const float2 t_hi_re = make_float2(rot.x * tc.inv_scale, rot.y * tc.inv_scale);     // times a real part
            const float2 t_hi_im = make_float2(-t_hi_re.y, t_hi_re.x);                          // times an imaginary part
            const float2 t_lo_re = make_float2(rot.x * tc.lo_scale, rot.y * tc.lo_scale);
            const float2 t_lo_im = make_float2(-t_lo_re.y, t_lo_re.x);
            constexpr int NCH = (4 * R + 15) / 16;      // 16-column pieces that hold outputs
#pragma unroll
            for (int ch = 0; ch < NCH; ++ch) {
                uint32_t v[16];
                tc_ld16(tmem_base + lane_addr + (uint32_t)(acc * N + ch * 16), v);
                tc_wait_ld();
                if (ch == NCH - 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_empty[acc]);   // accumulator back to the MMA warp
                }
                if ((p.debug_mode & 255) == 3) continue;   // debug_mode 3: no epilogue arithmetic or stores (tuning ceiling)
                float2 z[RC];
#pragma unroll
                for (int r = 0; r < RC; ++r) {
                    float2 zz = make_float2(__uint_as_float(v[4 * r + 1]) * t_lo_re.x, __uint_as_float(v[4 * r + 1]) * t_lo_re.y);
                    zz = ffma2(__uint_as_float(v[4 * r + 3]), t_lo_im, zz);
                    zz = ffma2(__uint_as_float(v[4 * r + 0]), t_hi_re, zz);
                    z[r] = ffma2(__uint_as_float(v[4 * r + 2]), t_hi_im, zz);
                }
                const long long l = (p.debug_mode & 255) == 4 ? (long long)(z[0].x == 1.2345f) : left - ch * 4;   // debug_mode 4: (almost) no stores
                float2* oc = o + ch * 4;
                if (RC >= 2 && p.vec_store) {
#pragma unroll
                    for (int r = 0; r + 1 < RC; r += 2) {
                        st_cs_v4_if(oc + r, z[r].x, z[r].y, z[r + 1].x, z[r + 1].y, l >= r + 2);
                        st_cs_v2_if(oc + r, z[r].x, z[r].y, l == r + 1);
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < RC; ++r) st_cs_v2_if(oc + r, z[r].x, z[r].y, l >= r + 1);
                }
            }
            acc ^= 1;
            if (acc == 0) fpar ^= 1u;
            cs += gs;
            cc += gc;
            if (cc >= cps) { cc -= cps; ++cs; }
        }
    } else if (warp >= 8) {
        // ------------------------------------------------------------------ unpack warps: NTEAM teams, team t takes the tiles
        // k = t (mod NTEAM).  With all twelve warps on every tile the stage is handed over when the slowest of them has drained its
        // stores behind the tensor core's operand reads, and the tensor pipe idled a third of every tile waiting for that
        // (profiles/r2_tensor_engine_trace.md); a team has NTEAM tile times per tile (configs[2]: 0.419 -> 0.405 -> 0.387 ms with one,
        // two, three teams)
        constexpr int NUT = NU / NTEAM;   // warps per team
        static_assert(NU % NTEAM == 0, "teams of equal size");
        const int team = (warp - 8) % NTEAM, u = (warp - 8) / NTEAM;   // my team takes the tiles k = team (mod NTEAM)
        // lane -> 16-sample group inside a run of 32 groups.  A quarter warp's two STS.128 must hit eight distinct 16-byte bank
        // groups: four even sub-streams of one row and the same four of the next (sub-stream pitch = odd number of units), so
        // lane bit 2 selects the row (group bit LOG_NS - 1) and the other lane bits fill the remaining group bits in order.
        constexpr int HB = S::LOG_NS - 1;   // log2(groups per row)
        const int low = lane & 3, rsel = (lane >> 2) & 1, rest = lane >> 3;   // 2 + 1 + 2 bits
        const int gl = low | ((rest & ((1 << (HB - 2)) - 1)) << 2) | (rsel << HB) | ((rest >> (HB - 2)) << (HB + 1));
        constexpr int UB = S::UNP_BATCH * NTEAM;    // groups a lane has in flight: all loads first, then the integer work, then the stores
        // a lane's groups are GSTEP apart, so both its raw address (20 bytes per group) and its destination (2 * GSTEP units on =
        // the same sub-stream, 2 * GSTEP / NS rows down) advance by compile-time constants
        constexpr int GSTEP = 32 * NUT, LD_STEP = 20 * GSTEP, ST_STEP = 2 * GSTEP / NS * 16;
        static_assert((2 * GSTEP) % NS == 0, "a lane must stay on one sub-stream pair");
        const int gfirst = u * 32 + gl;
        const uint32_t ld_off = 20u * (uint32_t)gfirst;
        const uint32_t st_off = (uint32_t)((2 * gfirst) & (NS - 1)) * (uint32_t)tc.a_pitch + (uint32_t)((2 * gfirst) >> S::LOG_NS) * 16u;
        // One batch covers a lane's share of a tile (the launcher guarantees n_groups <= UB * GSTEP).  The raw words of the
        // team's next tile are loaded BEFORE the fence / arrive of this one, so the shared-memory latency of the loads overlaps
        // the drain of the stores and the integer work of a tile starts from registers.
        const bool on = (p.debug_mode & 255) != 2;   // debug_mode 2: no unpack (tuning ceiling)
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
                for (int i = 0; i < 5; ++i) rw[b][i] = valid[b] ? src[i] : 0u;
            }
        };
        // tile k lives in sample stage k % n_a (free once the MMAs of tile k - n_a are done) and raw slot k % n_raw; the cursors
        // step by NTEAM without divisions (the launcher guarantees n_a >= NTEAM; the raw ring may be shorter)
        int as = team, rs = team, rn = team;              // stage / raw slot of my tile, raw slot of my next tile
        uint32_t epar = 1, rpar = 0;                      // parities: stage free (first pass: free), next raw slot full
        auto step = [&](int& idx, uint32_t& par, int n) {
            idx += NTEAM;
            while (idx >= n) { idx -= n; par ^= 1u; }
        };
        if (team < n_k) {
            mbar_wait_uni(&raw_full[rs], 0);
            load_raw(rs);
        }
        for (int k = team; k < n_k; k += NTEAM) {
            const long long t0 = p.dbg ? clock64() : 0;
            mbar_wait_uni(&a_empty[as], epar);
            const long long t2 = p.dbg ? clock64() : 0;
            if (p.dbg) tw0 += t2 - t0;
            if (u == 0) TC_TRACE(1, k);
            unsigned char* sp = asm_ + (size_t)as * tc.a_stage_bytes + st_off;
#pragma unroll
            for (int b = 0; b < UB; ++b) {
                uint32_t h[8];
                tc_unpack16(rw[b], h, tc);
                // units 2g (even sub-stream) and 2g + 1 (the next sub-stream, same row)
                unsigned char* dst = sp + b * ST_STEP;
                if (valid[b]) {
                    *reinterpret_cast<uint4*>(dst) = make_uint4(h[0], h[1], h[2], h[3]);
                    *reinterpret_cast<uint4*>(dst + tc.a_pitch) = make_uint4(h[4], h[5], h[6], h[7]);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&raw_empty[rs]);   // every lane's raw words have been consumed
            if (p.dbg) tw1 += clock64() - t2;
            // the team's next tile: its raw words now if they have landed (the usual case: the producer runs slots ahead), else
            // after the hand-over of this stage, so that a late copy never delays the MMA of the tile just unpacked
            const bool more = k + NTEAM < n_k;
            step(rn, rpar, tc.n_raw);
            const bool early = more && mbar_test_uni(&raw_full[rn], rpar);
            if (early) load_raw(rn);
            fence_proxy_async();   // my stores before the tensor core's reads of this stage
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_full[as]);
            if (u == 0) TC_TRACE(2, k);
            if (u == NUT - 1) TC_TRACE(3, k);
            if (u == 0 && early) TC_TRACE(9, k);
            if (more && !early) {
                const long long t3 = p.dbg ? clock64() : 0;
                mbar_wait_uni(&raw_full[rn], rpar);
                if (p.dbg) tw0 += clock64() - t3;
                load_raw(rn);
            }
            step(as, epar, tc.n_a);
            rs = rn;
        }
    }

    if (p.dbg && lane == 0 && (warp == 0 || warp == 1 || warp == 4 || warp == 8)) {
        // counters 2 .. 13: (wait 0, wait 1, total) of the producer, MMA, first epilogue and first unpack warp
        const int role = warp == 0 ? 0 : (warp == 1 ? 1 : (warp == 4 ? 2 : 3));
        atomicAdd(p.dbg + 2 + 3 * role, (unsigned long long)tw0);
        atomicAdd(p.dbg + 3 + 3 * role, (unsigned long long)tw1);
        atomicAdd(p.dbg + 4 + 3 * role, (unsigned long long)(clock64() - t_begin));
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)S::tmem_cols(D)) : "memory");
    }
}

}  // namespace ddck
