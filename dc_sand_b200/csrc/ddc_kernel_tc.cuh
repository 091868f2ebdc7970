// Tensor-core engine for PACKED 10-bit input ("kernel TC", opt-in: option "variant" = 13 / engine="tensor").
//
// Why it exists: packed input is FP32-bound by 3.2x on the CUDA cores (1.75 B against 64 flop per sample at T = 256, D = 16), so
// the fused-unpack fast-FIR kernel sits at 26 % of the HBM roofline however well it is scheduled.  10-bit samples are EXACT in
// fp16, and a decimating FIR over a tile of outputs is a Toeplitz product, so for this input type -- and only for it -- the
// filter can run on tcgen05 without giving up the reference's precision: taps are split hi + lo into two fp16 values
// (22 significant bits, the float32 taps of the CUDA-core kernels have 24), products are exact, accumulation is FP32 in TMEM.
// The float32-input kernels stay on the CUDA cores (north star): a float32 sample is not exact in any tensor-core input type.
//
// Formulation.  One MMA row covers 64 consecutive input samples = R = 64 / D outputs:
//      Y[row, (r, c)] = sum_k  X[row, k] * B[(r, c), k],     X[row, k] = x[64 row + k],   k = 0 .. K-1,  K = 64 - D + T (padded to 16)
//      B[(r, c), k]   = part c of tap  k - D r               (c: re_hi, re_lo, im_hi, im_lo;  0 outside [0, T))
// i.e. M = 128 rows (8192 samples) x N = 4 R columns x K per tile; the structured zeros of B cost (R - 1) D / T extra MACs
// (19 % at T = 256, D = 16) and make N a legal tcgen05 shape.  X is never materialised: the unpack warps write the fp16 sample
// stream into shared memory as EIGHT SUB-STREAMS of 16-byte units (unit u = samples 8u .. 8u+7 goes to sub-stream u % 8, row
// u / 8), and because consecutive rows of X start exactly one row further on in every sub-stream, the canonical no-swizzle
// K-major operand layout of tcgen05 (8-row core matrices of 16-byte rows, SBO = 128 bytes between row groups, LBO = the
// distance between two sub-streams for the two K-units of one MMA) describes the overlapping windows directly: K-unit q of
// row m is unit 8 m + q = sub-stream q % 8, row m + q / 8.  An MMA for K-units (2i, 2i+1) just starts (2i) / 8 rows down
// sub-stream (2i) % 8.
//
// Pipeline per CTA (one per SM, persistent, 16 warps):
//      warp 0      TMA producer: one bulk copy of packed bytes per tile -> raw ring
//      warps 8-15  unpack: 20 bytes -> 16 fp16 per lane and step (integer work, bit-exact), two STS.128 into the sub-streams
//      warp 1      one elected thread issues K / 16 tcgen05.mma per tile into one of two TMEM accumulators
//      warps 4-7   epilogue: tcgen05.ld, hi + lo recombination, NCO rotation (64-bit fixed-point phase), streaming stores
// SASS evidence: UBLKCP, UTCHMMA, LDTM (profiles/r2_sass_mnemonics.txt).
#pragma once
#include <cuda_fp16.h>

#include "ddc_common.cuh"

#ifndef DDCB200_TC_NUNP
#define DDCB200_TC_NUNP 12
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
    float inv_scale;          // 1 / S
    float lo_scale;           // 2^-11 / S
};

template <int NS_>
struct TcShape {
    static constexpr int NS = NS_;                        // sub-streams = 16-byte units per MMA row
    static constexpr int TILE_ROWS = 128;
    static constexpr int ROW_S = 8 * NS;                  // samples per MMA row
    static constexpr int TILE_S = TILE_ROWS * ROW_S;      // samples per tile
    static constexpr int TILE_PACKED = TILE_S / 4 * 5;    // packed bytes per tile
    static constexpr int NUNP = DDCB200_TC_NUNP;          // unpack warps
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
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <int D, int NS>
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
            mbar_init(&raw_empty[s], NU);
            mbar_init(&a_full[s], NU);
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
            __syncwarp();
            if (++as == tc.n_a) { as = 0; apar ^= 1u; }
            acc ^= 1;
            if (acc == 0) cpar ^= 1u;
        }
    } else if (warp >= 4 && warp < 8) {
        // ------------------------------------------------------------------ epilogue warps: TMEM lanes 32 (warp % 4) ..
        const int row = (warp & 3) * 32 + lane;
        const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
        const unsigned long long row_dph = (unsigned long long)S::ROW_S * p.step_fx;
        const unsigned long long out_dph = (unsigned long long)D * p.step_fx;
        const unsigned long long tile_dph = (unsigned long long)S::TILE_S * p.step_fx;
        const unsigned long long row_ph = p.phase0_fx + (unsigned long long)row * row_dph;
        constexpr int RC = R < 4 ? R : 4;           // outputs per 16-column piece
        // NCO: one polynomial rotation per piece (64-bit fixed-point phase of its first output), outputs 1 .. 3 of the piece by
        // the angle-addition step e^{-j 2 pi r D step} (three constants per thread): one extra rounding, 14 instead of 45
        // instructions per output
        float2 rstep[RC > 1 ? RC - 1 : 1];
#pragma unroll
        for (int r = 1; r < RC; ++r) rstep[r - 1] = nco_rot((unsigned long long)r * out_dph);
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
            unsigned long long ph = row_ph + (unsigned long long)cc * tile_dph;
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
                const float2 rot0 = nco_rot_bf(ph);
                ph += (unsigned long long)RC * out_dph;
                float2 z[RC];
#pragma unroll
                for (int r = 0; r < RC; ++r) {
                    const float yre = fmaf(__uint_as_float(v[4 * r + 1]), tc.lo_scale, __uint_as_float(v[4 * r + 0]) * tc.inv_scale);
                    const float yim = fmaf(__uint_as_float(v[4 * r + 3]), tc.lo_scale, __uint_as_float(v[4 * r + 2]) * tc.inv_scale);
                    const float2 y = r == 0 ? make_float2(yre, yim) : cmul(make_float2(yre, yim), rstep[r > 0 ? r - 1 : 0]);
                    z[r] = cmul(y, rot0);
                }
                const long long l = left - ch * 4;
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
        // ------------------------------------------------------------------ unpack warps: all of them share every tile
        const int u = warp - 8;
        const __half2 bias = __floats2half2_rn(1536.f, 1536.f);
        uint32_t mulk[4] = {1u << 10, 1u << 14, 1u << 18, 1u << 22};
        unsigned long long addend = 0x6400640000000000ull;
        asm volatile("" : "+r"(mulk[0]), "+r"(mulk[1]), "+r"(mulk[2]), "+r"(mulk[3]), "+l"(addend));
        // lane -> 16-sample group inside a run of 32 groups.  A quarter warp's two STS.128 must hit eight distinct 16-byte bank
        // groups: four even sub-streams of one row and the same four of the next (sub-stream pitch = odd number of units), so
        // lane bit 2 selects the row (group bit LOG_NS - 1) and the other lane bits fill the remaining group bits in order.
        constexpr int HB = S::LOG_NS - 1;   // log2(groups per row)
        const int low = lane & 3, rsel = (lane >> 2) & 1, rest = lane >> 3;   // 2 + 1 + 2 bits
        const int gl = low | ((rest & ((1 << (HB - 2)) - 1)) << 2) | (rsel << HB) | ((rest >> (HB - 2)) << (HB + 1));
        int rs = 0, as = 0;
        uint32_t rpar = 0, epar = 1;
        for (int k = 0; k < n_k; ++k) {
            const long long t0 = p.dbg ? clock64() : 0;
            mbar_wait_uni(&raw_full[rs], rpar);
            const long long t1 = p.dbg ? clock64() : 0;
            mbar_wait_uni(&a_empty[as], epar);
            const long long t2 = p.dbg ? clock64() : 0;
            if (p.dbg) tw0 += t2 - t0;
            const unsigned char* raw = rsm + (size_t)rs * tc.raw_slot_bytes;
            unsigned char* ast = asm_ + (size_t)as * tc.a_stage_bytes;
            const int g_end = (p.debug_mode & 255) == 2 ? 0 : tc.n_groups;   // debug_mode 2: no unpack (tuning ceiling)
#pragma unroll 2
            for (int g = u * 32 + gl; g < g_end; g += 32 * NU) {
                // 16 samples = 160 bits, big-endian bit stream: sample s = bits [10 s, 10 s + 10) from the top
                const uint32_t* rw = reinterpret_cast<const uint32_t*>(raw + 20 * g);
                uint32_t be[5];
#pragma unroll
                for (int i = 0; i < 5; ++i) be[i] = __byte_perm((p.debug_mode & 0x20) ? (uint32_t)g * 2654435761u + i : rw[i], 0, 0x0123);
                uint32_t h[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    // pair q = samples 2q, 2q+1 = 20 bits from bit 20 q of the group.  One LOP3 flips the two sign bits (offset
                    // binary u = v + 512) and isolates the pair; one 32 x 32 -> 64 bit multiply-add by a power of two then leaves
                    // u_a (+ the exponent pattern of 1024.0 for both halves) in the high word and u_b in the top ten bits of the
                    // low word; one shift-add drops u_b into the high half: 0x6400 | u in each half = 1024 + u, minus 1536 = v.
                    // (4 instructions per pair, 2 on the ALU pipe and 2 on the FMA pipe; pairs that straddle two words take a
                    // funnel shift first.)
                    const int bit = 20 * q, wi = bit >> 5, sh = bit & 31;
                    const bool straddle = sh + 20 > 32;
                    const uint32_t src = straddle ? __funnelshift_l(be[wi + 1 > 4 ? 4 : wi + 1], be[wi], sh) : be[wi];
                    const int s2 = straddle ? 0 : sh;
                    uint32_t fm;
                    asm("lop3.b32 %0, %1, %2, %3, 0x28;" : "=r"(fm) : "r"(src), "r"(0x80200000u >> s2), "r"(0xFFFFF000u >> s2));   // (src ^ x) & m
                    uint32_t plo, phi;   // IMAD.WIDE with the 64-bit addend; the multipliers sit in registers so that ptxas keeps the multiply
                    asm("{\n.reg .b64 t;\nmad.wide.u32 t, %2, %3, %4;\nmov.b64 {%0, %1}, t;\n}" : "=r"(plo), "=r"(phi) : "r"(fm), "r"(mulk[s2 >> 2]), "l"(addend));
                    const uint32_t w = phi + (plo >> 6);
                    const __half2 hv = __hsub2(*reinterpret_cast<const __half2*>(&w), bias);   // exact
                    h[q] = *reinterpret_cast<const uint32_t*>(&hv);
                }
                // units 2g (even sub-stream) and 2g + 1 (the next sub-stream, same row)
                unsigned char* dst = ast + (size_t)((2 * g) & (NS - 1)) * tc.a_pitch + (size_t)((2 * g) >> S::LOG_NS) * 16;
                if (!(p.debug_mode & 0x10) || (h[0] ^ h[5]) == 0x12345678u) {   // debug bit 0x10: no stores (tuning)
                    *reinterpret_cast<uint4*>(dst) = make_uint4(h[0], h[1], h[2], h[3]);
                    *reinterpret_cast<uint4*>(dst + tc.a_pitch) = make_uint4(h[4], h[5], h[6], h[7]);
                }
            }
            if (p.dbg) tw1 += clock64() - t2;
            fence_proxy_async();   // my stores before the tensor core's reads of this stage
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&a_full[as]);
                mbar_arrive(&raw_empty[rs]);
            }
            if (++rs == tc.n_raw) { rs = 0; rpar ^= 1u; }
            if (++as == tc.n_a) { as = 0; epar ^= 1u; }
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
