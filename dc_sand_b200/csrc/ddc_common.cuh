// Shared device helpers for the fused DDC kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ddck {

// Complex taps travel as kernel parameters (constant bank 0): they are warp-uniform operands, so the compiler
// feeds them to FFMA2 through uniform registers (LDCU -> UR) with no vector-register or shared-memory traffic.
// Two complex taps per float4: (re[k], im[k], re[k+1], im[k+1]).
template <int MAXT>
struct TapsParam {
    float4 c2[MAXT / 2];
};

struct RunParams {
    const void* in;              // float32 samples, or packed 10-bit bytes
    float2* out;                 // complex64
    long long n_samples;         // N per stream
    long long in_stride;         // elements (float32) or bytes (packed) between streams
    long long out_stride;        // complex elements between streams
    long long n_out;             // M per stream
    long long m_begin;           // first output this launch is responsible for (per stream)
    long long tiles_per_stream;  // full tiles handled by the fused kernel
    long long total_tiles;       // tiles_per_stream * n_streams
    unsigned long long step_fx;  // frac(step) * 2^64           (NCO cycles per sample, fixed point)
    unsigned long long phase0_fx;// frac(sample_offset * step) * 2^64
    int n_taps;                  // T (padded to a whole number of tap blocks for the fused kernel)
    int n_tap_blocks;            // J = n_taps / D
    int halo_rows;               // extra rows staged behind each tile
    int n_streams;
    int vec_store;               // 1 if out pointer / stride allow 16-byte stores
    unsigned long long* dbg;     // optional diagnostic counters (wait cycles / total cycles of compute warps), or NULL
    int debug_mode;              // 0 normal; 1 compute only (no TMA, no waits); 2 memory only (no FIR) -- ceilings for tuning
    unsigned long long cps_magic;// kernel WS: ceil(2^64 / tiles_per_stream), 0 when tiles_per_stream == 1 (see chunk_of)
};

// (stream, chunk inside the stream) of global chunk gk without a hardware division: floor(gk / cps) = mulhi(gk, ceil(2^64 / cps)),
// exact while gk * cps < 2^64 (the host checks total_tiles * tiles_per_stream)
__device__ __forceinline__ void chunk_of(const RunParams& p, long long gk, int& cs, int& cc) {
    const unsigned long long q = p.cps_magic ? __umul64hi((unsigned long long)gk, p.cps_magic) : (unsigned long long)gk;
    cs = (int)q;
    cc = (int)(gk - (long long)q * p.tiles_per_stream);
}

// ---- NCO: exp(-j 2 pi ph / 2^64) from a 64-bit fixed-point phase ------------------------------------------
// Quadrant reduction in integers keeps the float conversion exact to 2^-28 cycle (2.3e-8 rad); the remaining
// angle |a| <= 1/4 (in units of pi) goes through sincospif.  No table.
__device__ __forceinline__ float2 nco_rot(unsigned long long ph) {
    const uint32_t up = (uint32_t)(ph >> 32);
    const uint32_t q = (up + 0x20000000u) >> 30;            // nearest quarter cycle (mod 4)
    const int32_t r = (int32_t)(up - (q << 30));            // residual in [-2^29, 2^29)
    float s, c;
    sincospif((float)r * (1.0f / 2147483648.0f), &s, &c);   // angle = pi * r / 2^31
    float re = c, im = -s;                                   // exp(-j a)
    float2 o;
    switch (q & 3u) {                                        // times (-j)^q
        case 0: o = make_float2(re, im); break;
        case 1: o = make_float2(im, -re); break;
        case 2: o = make_float2(-re, -im); break;
        default: o = make_float2(-im, re); break;
    }
    return o;
}

// Branch-free variant for code that must stay one basic block (deferred epilogues): after the integer quadrant reduction
// |x| <= 1/4, so sin(pi x) = x P(x^2) and cos(pi x) = Q(x^2) with degree-3 / degree-4 minimax polynomials (absolute error
// 8.5e-8 / 5.8e-8, i.e. float32 rounding level); the quadrant is applied with selects.  sincospif carries branches for
// special values that split the block.
__device__ __forceinline__ float2 nco_rot_bf(unsigned long long ph) {
    const uint32_t up = (uint32_t)(ph >> 32);
    const uint32_t q = (up + 0x20000000u) >> 30;
    const int32_t r = (int32_t)(up - (q << 30));
    const float x = (float)r * (1.0f / 2147483648.0f);
    const float s2 = x * x;
    float sp = fmaf(s2, -5.890768766e-01f, 2.549767017e+00f);
    sp = fmaf(sp, s2, -5.167707920e+00f);
    sp = fmaf(sp, s2, 3.141592741e+00f);
    float cs = fmaf(s2, 2.313292474e-01f, -1.335044503e+00f);
    cs = fmaf(cs, s2, 4.058707237e+00f);
    cs = fmaf(cs, s2, -4.934802055e+00f);
    cs = fmaf(cs, s2, 1.0f);
    const float re = cs, im = -(sp * x);                     // exp(-j pi x)
    const bool sw = (q & 1u) != 0, ng = (q & 2u) != 0;       // times (-j)^q
    const float a = sw ? im : re, b = sw ? -re : im;
    return make_float2(ng ? -a : a, ng ? -b : b);
}

// predicated streaming stores (never a branch)
__device__ __forceinline__ void st_cs_v4_if(float2* ptr, float a, float b, float c, float d, bool pr) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.u32 p, %5, 0;\n"
        "@p st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};\n"
        "}\n" ::"l"(ptr),
        "f"(a), "f"(b), "f"(c), "f"(d), "r"((uint32_t)pr));
}
__device__ __forceinline__ void st_cs_v2_if(float2* ptr, float a, float b, bool pr) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.u32 p, %3, 0;\n"
        "@p st.global.cs.v2.f32 [%0], {%1, %2};\n"
        "}\n" ::"l"(ptr),
        "f"(a), "f"(b), "r"((uint32_t)pr));
}

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}

// ---- packed 10-bit samples: sample k = bits [10k, 10k+10) of a big-endian bit stream, two's complement ----
__device__ __forceinline__ int unpack10_at(const uint8_t* __restrict__ p, long long k) {
    const long long bit = k * 10;
    const long long byte = bit >> 3;
    const int sh = (int)(bit & 7);                           // 0, 2, 4, 6
    const uint32_t w = ((uint32_t)p[byte] << 8) | (uint32_t)p[byte + 1];
    const int v = (int)((w >> (6 - sh)) & 0x3FFu);
    return (v ^ 0x200) - 0x200;                              // sign-extend 10 bits
}

// 4 samples from 5 bytes given as a 40-bit big-endian word in (hi8, lo32)
__device__ __forceinline__ void unpack10_word(uint32_t b0, uint32_t lo, int* v) {
    // word = b0<<32 | lo ; sample0 = bits 39..30, sample1 = 29..20, sample2 = 19..10, sample3 = 9..0
    const int s0 = (int)(((b0 << 2) | (lo >> 30)) & 0x3FFu);
    const int s1 = (int)((lo >> 20) & 0x3FFu);
    const int s2 = (int)((lo >> 10) & 0x3FFu);
    const int s3 = (int)(lo & 0x3FFu);
    v[0] = (s0 ^ 0x200) - 0x200;
    v[1] = (s1 ^ 0x200) - 0x200;
    v[2] = (s2 ^ 0x200) - 0x200;
    v[3] = (s3 ^ 0x200) - 0x200;
}

// ---- mbarrier / bulk-copy (TMA 1-D) primitives --------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_if(uint64_t* bar, bool pr) {   // predicated, never a branch
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.u32 p, %1, 0;\n"
        "@p mbarrier.arrive.shared::cta.b64 _, [%0];\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"((uint32_t)pr)
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// Waits for CONVERGED warps in which every lane polls the same location: the loop branch is taken on a warp vote and marked
// .uni, so the compiler knows the warp does not diverge here.  With a per-thread predicate on the back edge (mbar_wait,
// a C++ spin loop) every value carried across the wait stops being warp-uniform for ptxas -- tap pointers then leave the
// uniform registers and the taps are fetched with per-thread LDC instead of LDCU (measured 2.2x slower in kernel WS).
__device__ __forceinline__ void spin_until_eq_uni(const volatile int* addr, int want) {
    asm volatile(
        "{\n"
        ".reg .pred p, q;\n"
        ".reg .b32 v;\n"
        "SPIN_%=:\n"
        "ld.volatile.shared.b32 v, [%0];\n"
        "setp.eq.s32 p, v, %1;\n"
        "vote.sync.all.pred q, p, 0xffffffff;\n"
        "@!q bra.uni SPIN_%=;\n"
        "}\n" ::"r"(smem_u32(const_cast<const int*>(addr))),
        "r"(want)
        : "memory");
}
__device__ __forceinline__ void mbar_wait_uni(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p, q;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "vote.sync.all.pred q, p, 0xffffffff;\n"
        "@!q bra.uni WAIT_%=;\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// one poll, warp-uniform result: has the phase with this parity completed?
__device__ __forceinline__ bool mbar_test_uni(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p, q;\n"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "vote.sync.all.pred q, p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, q;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// global -> shared bulk copy (SASS: UBLKCP), completion counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// shared -> global bulk copy (SASS: UBLKCP / bulk store), completion tracked by the bulk async-group
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
// one elected lane of a converged warp (ptxas then predicates uniform-datapath instructions such as UBLKCP on it instead
// of serialising "divergent" lanes in a BRA.U.ANY loop)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred P;\n"
        "elect.sync _|P, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, P;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

#ifndef DDCB200_PIN_FFMA2_ORDER
#define DDCB200_PIN_FFMA2_ORDER 1
#endif
__device__ __forceinline__ float2 ffma2(float x, float2 t, float2 acc) {
    // acc(re,im) += x * t(re,im); SASS: FFMA2 Racc, Rx.F32 (broadcast), URt.F32x2, Racc
#if DDCB200_PIN_FFMA2_ORDER
    // volatile asm keeps the source order of the FMAs (tap-major, R independent accumulator chains round-robin):
    // left alone, the scheduler interleaves only two chains and the 2-cycle FFMA2 issue cadence exposes its latency.
    asm volatile(
        "{\n"
        ".reg .b64 xx, tt, aa;\n"
        "mov.b64 xx, {%2, %2};\n"
        "mov.b64 tt, {%3, %4};\n"
        "mov.b64 aa, {%0, %1};\n"
        "fma.rn.f32x2 aa, xx, tt, aa;\n"
        "mov.b64 {%0, %1}, aa;\n"
        "}\n"
        : "+f"(acc.x), "+f"(acc.y)
        : "f"(x), "f"(t.x), "f"(t.y));
    return acc;
#else
    return __ffma2_rn(make_float2(x, x), t, acc);
#endif
}

}  // namespace ddck
