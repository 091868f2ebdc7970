// Rotating-window tile kernel ("ddc_fused_kernel"): the fused DDC for any tap count up to 2048 at D = 4 .. 64 (see the
// description of the kernel families at the top of ddc_kernels.cuh).
#pragma once
#include "ddc_common.cuh"

namespace ddck {


// =============================================================================================================
// Fused persistent kernel
// =============================================================================================================
// Shared-memory layout of one pipeline stage.  A "thread-row" is the ROW = R*D float32 samples that produce R
// consecutive outputs (256 B for every supported D).  S consecutive thread-rows form a "super-row" that is staged by
// ONE bulk copy (S*ROW*4 bytes, 8 KB for S = 32) and is followed by a 16-byte pad, i.e. super-row pitch = S*ROW + 4
// floats.  A tile is NROWS = 256 thread-rows (+ halo).  Within a warp, lane bits [0,3) = i select one of eight
// consecutive super-rows, so the eight lanes of a quarter warp issue LDS.128 at addresses that differ by
// (S*ROW + 4) floats = 16 B mod 128 B: eight distinct bank groups, no bank conflicts, with dense TMA-friendly rows
// (a tile + halo is 9 bulk copies).
//
// KS = tap split: the T taps of one thread-row are shared by KS threads, lane l of warps w and w + 8 (the half must be
// warp-uniform so that taps stay uniform-register operands); each accumulates J/KS tap blocks for the same R outputs,
// the halves are exchanged through a small double-buffered shared-memory area under a 64-thread named barrier, and
// each thread then rotates and stores R/KS outputs.  KS = 2 doubles the resident compute warps (16 per SM) for the
// same input staging footprint.
template <int D, int R, int S, int KS>
struct FusedCfg {
    static constexpr int ROW = R * D;            // samples per thread-row
    static constexpr int SRP = S * ROW + 4;      // super-row pitch in floats
    static constexpr int V = D / 4;              // float4 per tap block
    static constexpr int NROWS = 256;            // thread-rows per tile
    static constexpr int QN = 4;                 // thread-rows per warp at the same super-row
    static constexpr int WPG = S / QN;           // warps per group of 8 super-rows
    static constexpr int SENDN = (R >= 2) ? R / 2 : 1;  // float2 exchanged per thread when KS == 2
    static constexpr size_t XBUF_BYTES = (KS == 2) ? (size_t)2 * 2 * NROWS * SENDN * sizeof(float2) : 0;
    static constexpr int NT = NROWS * KS;        // compute threads
    static constexpr int TILE_OUT = NROWS * R;   // outputs per tile
    static_assert(D % 4 == 0, "D must be a multiple of 4");
    static_assert(KS == 1 || KS == 2, "tap split 1 or 2");
    static_assert(S == 4 || S == 8 || S == 16 || S == 32, "S must be 4, 8, 16 or 32");
    static_assert((S * ROW / 4) % 8 == 0, "super-row must be a whole number of 128-byte lines");
    static_assert(NROWS % (8 * S) == 0, "tile must be a whole number of 8-super-row groups");
    __host__ __device__ static constexpr int row_offset(int row) { return (row / S) * SRP + (row % S) * ROW; }
    __host__ __device__ static constexpr size_t stage_floats(int halo_rows) {
        // whole super-rows of the tile and of the halo, plus a last partial super-row (16-byte pad kept)
        return (size_t)(NROWS / S + halo_rows / S) * SRP + (size_t)((halo_rows % S) ? (halo_rows % S) * ROW + 4 : 0);
    }
};

template <int D, int R, int S, int KS, int STAGES, int MAXT, bool PACKED>
__global__ void __launch_bounds__(FusedCfg<D, R, S, KS>::NT + 32, 1)
ddc_fused_kernel(const __grid_constant__ RunParams p, const __grid_constant__ TapsParam<MAXT> taps) {
    using C = FusedCfg<D, R, S, KS>;
    constexpr int ROW = C::ROW, SRP = C::SRP, V = C::V, NT = C::NT, NROWS = C::NROWS;
    constexpr int TILE_OUT = C::TILE_OUT;
    constexpr int TILE_S = NROWS * ROW;  // samples per tile

    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw);
    uint64_t* empty_bar = full_bar + STAGES;
    float2* xbuf = reinterpret_cast<float2*>(smem_raw + 128);  // [parity][half][NROWS][SENDN], KS == 2 only
    float* buf = reinterpret_cast<float*>(smem_raw + 128 + C::XBUF_BYTES);
    const int stage_floats = (int)C::stage_floats(p.halo_rows);

    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);  // warp-uniform by construction: lets ptxas keep loop state in uniform registers
    const int lane = tid & 31;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], NT / 32);
        }
        mbar_fence_init();
    }
    __syncthreads();

    // Tiles are numbered stream-major; a CTA walks tile = blockIdx.x, += gridDim.x.  (stream, tile-in-stream) are
    // advanced incrementally (no 64-bit division on the per-tile path).
    const int tps = (int)p.tiles_per_stream;
    const int gstep_s = (int)(gridDim.x / tps), gstep_t = (int)(gridDim.x % tps);
    int cur_s = (int)(blockIdx.x / tps), cur_t = (int)(blockIdx.x % tps);
    const int n_iter = (int)((p.total_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);  // blockIdx.x < total_tiles

    if (warp == NT / 32) {
        // ------------------------------------------------ producer warp
        const int n_sr = NROWS / S + (p.halo_rows + S - 1) / S;  // super-rows per stage
        const long long want = (long long)(NROWS + p.halo_rows) * ROW;
        for (int it = 0; it < n_iter && p.debug_mode != 1 && p.debug_mode != 3; ++it) {
            const int stage = it % STAGES;
            const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
            mbar_wait(&empty_bar[stage], ph ^ 1u);
            const float* src = reinterpret_cast<const float*>(p.in) + (long long)cur_s * p.in_stride + (long long)cur_t * TILE_S;
            float* dst = buf + (size_t)stage * stage_floats;
            const long long valid = p.n_samples - (long long)cur_t * TILE_S;  // samples of this stream from the tile start
            if (valid >= want) {
                // full tile: one elected lane issues the TMA bulk copies (8 KB each)
                if (lane == 0) {
                    mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)want * 4u);
                    int left = NROWS + p.halo_rows;
#pragma unroll 1
                    for (int sr = 0; left > 0; ++sr, left -= S) {
                        const int nrow = left < S ? left : S;
                        bulk_g2s(dst + sr * SRP, src + (size_t)sr * S * ROW, (uint32_t)nrow * ROW * 4u, &full_bar[stage]);
                    }
                }
            } else {
                // ragged last tile of a stream: bulk-copy what is whole 16-byte groups, hand-copy the last 1-3 samples,
                // zero-fill the rest (zero taps of a padded tap set must not meet stale shared memory)
                uint32_t tx = 0;
                for (int sr = 0; sr < n_sr; ++sr) {
                    const int left = NROWS + p.halo_rows - sr * S;
                    const int cap = (left < S ? left : S) * ROW;          // floats this super-row holds
                    const long long s0 = (long long)sr * S * ROW;
                    long long cnt = valid - s0;
                    cnt = cnt < 0 ? 0 : (cnt > cap ? cap : cnt);
                    const int bulk = (int)cnt & ~3;
                    for (int k = bulk + lane; k < cap; k += 32) dst[sr * SRP + k] = (k < (int)cnt) ? src[s0 + k] : 0.f;
                    tx += (uint32_t)bulk * 4u;
                }
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive_expect_tx(&full_bar[stage], tx);  // release: the plain stores above become visible
                    for (int sr = 0; sr < n_sr; ++sr) {
                        const int left = NROWS + p.halo_rows - sr * S;
                        const int cap = (left < S ? left : S) * ROW;
                        const long long s0 = (long long)sr * S * ROW;
                        long long cnt = valid - s0;
                        cnt = cnt < 0 ? 0 : (cnt > cap ? cap : cnt);
                        const int bulk = (int)cnt & ~3;
                        if (bulk > 0) bulk_g2s(dst + sr * SRP, src + s0, (uint32_t)bulk * 4u, &full_bar[stage]);
                    }
                }
            }
            __syncwarp();
            cur_s += gstep_s;
            cur_t += gstep_t;
            if (cur_t >= tps) { cur_t -= tps; ++cur_s; }
        }
    } else {
        // ------------------------------------------------ compute warps
        const int JH = p.n_tap_blocks / KS;                  // tap blocks per thread
        const int li = lane & 7;
        const int lq = lane >> 3;
        const int half = (KS == 2) ? (warp / (NROWS / 32)) : 0;   // warp-uniform
        const int wr = warp % (NROWS / 32);
        const int g = (wr / C::WPG) * (8 * S) + li * S + (wr % C::WPG) * C::QN + lq;  // thread-row within the tile
        const int jbeg = half * JH;                          // first tap block of this thread (multiple of R)
        const int row0 = g + jbeg / R;
        constexpr int RO = (KS == 2 && R >= 2) ? R / 2 : R;  // outputs this thread finishes
        const int rbeg = (KS == 2 && R >= 2) ? half * RO : 0;
        // NCO rotation of output m = tile*TILE_OUT + g*R + rbeg + r is rot_tile(tile) * rot_thr[r]: the per-thread factor
        // is computed once per kernel, the per-tile factor once per tile (one sincospif instead of R per tile).
        float2 rot_thr[RO];
#pragma unroll
        for (int r = 0; r < RO; ++r)
            rot_thr[r] = nco_rot((unsigned long long)((long long)(g * R + rbeg + r) * D) * p.step_fx);
        const unsigned long long tile_dph = (unsigned long long)((long long)TILE_OUT * D) * p.step_fx;

        for (int it = 0; it < n_iter; ++it) {
            const int stage = it % STAGES;
            const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
            if (p.debug_mode != 1 && p.debug_mode != 3) mbar_wait(&full_bar[stage], ph);
            const float* sbuf = buf + (size_t)stage * stage_floats;

#ifndef DDCB200_SPLIT_ACC
#define DDCB200_SPLIT_ACC 0
#endif
            constexpr int NA = DDCB200_SPLIT_ACC ? 2 : 1;  // partial sums per output (more independent FMA chains)
            float2 accp[R][NA];
#pragma unroll
            for (int r = 0; r < R; ++r)
#pragma unroll
                for (int a = 0; a < NA; ++a) accp[r][a] = make_float2(0.f, 0.f);
            float4 xw[R][V];  // rotating window: slot (j + r) % R holds tap-block-sized sample block j + r
            // Row pointers advance by a thread-local recurrence (not a function of j0) so that the compiler keeps the
            // tap-block counter j0 in a uniform register and fetches taps with LDCU -> UR operands of FFMA2.
            const float* p0 = sbuf + C::row_offset(row0);
            int sub = row0 % S;
#pragma unroll
            for (int s = 0; s < R - 1; ++s)
#pragma unroll
                for (int v = 0; v < V; ++v) xw[s][v] = *reinterpret_cast<const float4*>(p0 + s * D + 4 * v);
            const float4* tbase = &taps.c2[(size_t)jbeg * (D / 2)];
            const int jend = (p.debug_mode == 2) ? 0 : JH;
            for (int j0 = 0; j0 < jend; j0 += R) {
                const bool wrap = (sub == S - 1);
                const float* p1 = p0 + ROW + (wrap ? 4 : 0);
                sub = wrap ? 0 : sub + 1;
#pragma unroll
                for (int jj = 0; jj < R; ++jj) {
                    const int srel = jj + R - 1;  // newest block of this step, relative to block j0 (row p0)
                    const float* src = (srel / R) ? p1 : p0;
                    if (p.debug_mode != 3) {  // 3: compute only AND no shared-memory loads in the loop (FMA ceiling)
#pragma unroll
                        for (int v = 0; v < V; ++v)
                            xw[srel % R][v] = *reinterpret_cast<const float4*>(src + (srel % R) * D + 4 * v);
                    }
                    const float4* tp = tbase + (size_t)(j0 + jj) * (D / 2);
#pragma unroll
                    for (int v = 0; v < V; ++v) {
                        const float4 ta = tp[2 * v], tb = tp[2 * v + 1];
#pragma unroll
                        for (int r = 0; r < R; ++r) accp[r][0] = ffma2(xw[(jj + r) % R][v].x, make_float2(ta.x, ta.y), accp[r][0]);
#pragma unroll
                        for (int r = 0; r < R; ++r) accp[r][NA - 1] = ffma2(xw[(jj + r) % R][v].y, make_float2(ta.z, ta.w), accp[r][NA - 1]);
#pragma unroll
                        for (int r = 0; r < R; ++r) accp[r][0] = ffma2(xw[(jj + r) % R][v].z, make_float2(tb.x, tb.y), accp[r][0]);
#pragma unroll
                        for (int r = 0; r < R; ++r) accp[r][NA - 1] = ffma2(xw[(jj + r) % R][v].w, make_float2(tb.z, tb.w), accp[r][NA - 1]);
                    }
                }
                p0 = p1;
            }
            float2 acc[R];
#pragma unroll
            for (int r = 0; r < R; ++r)
                acc[r] = (NA == 2) ? make_float2(accp[r][0].x + accp[r][NA - 1].x, accp[r][0].y + accp[r][NA - 1].y) : accp[r][0];
            // all shared-memory reads of this stage are done -> hand the slot back to the producer
            __syncwarp();
            if (lane == 0 && p.debug_mode != 1 && p.debug_mode != 3) mbar_arrive(&empty_bar[stage]);

            // epilogue: combine tap halves, rotate each output by the NCO phase of its first input sample, store
            float2 y[RO];
            bool writer = true;
            if (KS == 2) {
                constexpr int SN = C::SENDN;
                const int xt = wr * 32 + lane;
                float2* xs = xbuf + ((size_t)((it & 1) * 2 + half) * NROWS + xt) * SN;         // what I send
                const float2* xr = xbuf + ((size_t)((it & 1) * 2 + (half ^ 1)) * NROWS + xt) * SN;  // what my partner sent
                if (R >= 2) {
                    // this thread keeps outputs [half*RO, half*RO + RO) and receives the partner's partial sums for them
                    if (SN % 2 == 0) {
#pragma unroll
                        for (int r = 0; r < SN; r += 2) {
                            const float2 a = half ? acc[r] : acc[r + RO], b = half ? acc[r + 1] : acc[r + 1 + RO];
                            *reinterpret_cast<float4*>(xs + r) = make_float4(a.x, a.y, b.x, b.y);
                        }
                    } else {
#pragma unroll
                        for (int r = 0; r < SN; ++r) xs[r] = half ? acc[r] : acc[r + RO];
                    }
                    asm volatile("bar.sync %0, 64;" ::"r"(1 + wr) : "memory");
                    if (SN % 2 == 0) {
#pragma unroll
                        for (int r = 0; r < SN; r += 2) {
                            const float4 v = *reinterpret_cast<const float4*>(xr + r);
                            const float2 k0 = half ? acc[r + RO] : acc[r], k1 = half ? acc[r + 1 + RO] : acc[r + 1];
                            y[r] = make_float2(k0.x + v.x, k0.y + v.y);
                            y[r + 1] = make_float2(k1.x + v.z, k1.y + v.w);
                        }
                    } else {
#pragma unroll
                        for (int r = 0; r < SN; ++r) {
                            const float2 v = xr[r];
                            const float2 k0 = half ? acc[r + RO] : acc[r];
                            y[r] = make_float2(k0.x + v.x, k0.y + v.y);
                        }
                    }
                } else {
                    if (half) xs[0] = acc[0];
                    asm volatile("bar.sync %0, 64;" ::"r"(1 + wr) : "memory");
                    const float2 v = half ? make_float2(0.f, 0.f) : xr[0];
                    y[0] = make_float2(acc[0].x + v.x, acc[0].y + v.y);
                    writer = (half == 0);
                }
            } else {
#pragma unroll
                for (int r = 0; r < RO; ++r) y[r] = acc[r];
            }
            const float2 rot_tile = nco_rot(p.phase0_fx + (unsigned long long)cur_t * tile_dph);
            const long long m0 = (long long)cur_t * TILE_OUT + (g * R + rbeg);
            float2* o = p.out + (long long)cur_s * p.out_stride + m0;
#pragma unroll
            for (int r = 0; r < RO; ++r) y[r] = cmul(cmul(y[r], rot_thr[r]), rot_tile);
            if (writer) {
                if (m0 + RO <= p.n_out) {
                    if (p.vec_store && (RO % 2 == 0)) {
#pragma unroll
                        for (int r = 0; r < RO; r += 2)
                            __stcs(reinterpret_cast<float4*>(o + r), make_float4(y[r].x, y[r].y, y[r + 1].x, y[r + 1].y));
                    } else {
#pragma unroll
                        for (int r = 0; r < RO; ++r) __stcs(o + r, y[r]);
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < RO; ++r)
                        if (m0 + r < p.n_out) __stcs(o + r, y[r]);
                }
            }
            cur_s += gstep_s;
            cur_t += gstep_t;
            if (cur_t >= tps) { cur_t -= tps; ++cur_s; }
        }
    }
}

}  // namespace ddck
