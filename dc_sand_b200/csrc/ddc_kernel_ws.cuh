// Tensor-staged fast-FIR fused DDC -- "kernel WS": chunks staged by 5-D TMA tensor copies into swizzled shared-memory tiles.
// Three uses (WSCfg): SLICED staging for large decimations (D = 32, 64; described first), whole blocks in per-block tiles for
// D = 4 / 8 (two CTAs per SM), and whole blocks in whole-row tiles for short, HBM-bound filters at D = 4 / 8 / 16.
//
// The fast-FIR kernel (ddc_kernel_w.cuh) needs R = 8 outputs per thread so that one tap fetch feeds four FFMA2; with
// thread-rows of 128 samples that only holds at D = 16 (at D = 32 / 64 the phase-major kernels have R = 4 / 2 and run at
// 75 % / 43 % of the FP32 peak when the filter is long).  Making the thread-row 8 D samples long would make a chunk
// 32 / 64 KB, which no 8-warp ring of shared memory holds.  Instead the chunk is staged in SLICES: slice q holds, for
// every D-sample block, only the 16 samples [16 q, 16 q + 16) -- exactly the samples that phase groups 4q .. 4q+3 need.
// A slice of a chunk is 32 (+4 halo) rows of 8 blocks x 16 floats: the SAME shared-memory image as a D = 16 chunk, so
// the D = 16 inner loop runs on it unchanged (taps of phases 16 q ..).  A warp walks the D / 16 slices of its chunk one
// after the other, accumulating in registers, then stores.  The strided gather (64 contiguous bytes out of every 4 D) is
// a 5-D TMA tensor copy: dims (16 floats, D/16 slices, 8 blocks, thread-rows, streams), box (16, 1, 1, 36, 1) = 2304 B;
// out-of-range thread-rows are zero-filled by the TMA unit, which also takes care of stream tails.
#pragma once
#include <cuda.h>

#include "ddc_kernel_w.cuh"

namespace ddck {

template <int D, int JT>
struct WSCfg {
    static_assert(D == 4 || D == 8 || D == 16 || D == 32 || D == 64, "tensor-staged kernel: D = 32, 64 (sliced) and D = 4, 8, 16 (whole blocks)");
    static constexpr int JT_MAX = D == 4 ? 256 : (D == 8 ? 128 : 32);   // the halo grows by two thread-rows per 16 tap blocks
    static_assert(JT % 2 == 0 && (JT <= 32 || JT % 16 == 0) && JT <= JT_MAX, "tap blocks: even up to 32, then multiples of 16 up to JT_MAX");
    // D = 4 / 8 use the same staging with ONE slice that holds the whole D-sample block: thread-rows of 8 blocks (32 / 64
    // samples) give R = 8 outputs per thread, which the 1-D kernels (128-sample rows) cannot offer at these decimations.
    static constexpr int LPQ = D < 16 ? 1 : D / 16;   // slices per chunk
    static constexpr int DB = D < 16 ? D : 16;        // floats of one block inside a slice
    static constexpr int R = 8;                    // outputs (= blocks) per thread-row
    static constexpr int ROW = R * DB;             // floats of one thread-row inside a slice
    static constexpr int V = DB / 4;               // phase groups per slice
    static constexpr int JP = JT < 16 ? JT : 16;   // tap blocks per pass
    static constexpr int NJG = JT / JP;            // passes per slice
    static_assert(JT % JP == 0, "long filters are padded to a multiple of 16 tap blocks");
    static constexpr int RH = R / 2;
    static constexpr int NWP = R + JP - 1;         // register window of one pass (blocks)
    static constexpr int WROWS = (NWP - 1) / R + 1;
    // Tensor-mode TMA needs 128-byte aligned shared-memory destinations, so the 16-byte pad that rotates banks in the 1-D
    // kernels is not available.  Instead one TMA box is the slice of ONE block index for all 36 thread-rows of a slot:
    // 36 lines of 64 bytes (the inner box dimension), written with the 64-byte swizzle -- the 16-byte unit index
    // (address bits 4-5) is XORed with address bits 7-8, i.e. with (row / 2) % 4, and bit 6 is row % 2.  Lane L owns
    // thread-row L, so the eight lanes of a quarter warp read eight consecutive lines at the same logical unit and hit
    // eight distinct bank groups.  Tiles are 512-byte aligned (2304 bytes of data in a 2560-byte pitch).
    // (The 128-byte swizzle pads every 64-byte box row to a 128-byte line: tools/microbench/tma_swizzle_probe.cu.)
    // With 32-byte lines (D = 8) the 32-byte swizzle does the same job (unit ^= (row / 4) % 2, bits 5-6 = row % 4); 16-byte
    // lines (D = 4) are conflict-free as they are.
    // D = 4 / 8 with short filters (JT * D <= 64 taps, the HBM-bound cells): WHOLE-row tiles instead.  A thread-row is only
    // 128 / 256 bytes, so a line of a tile is a whole thread-row (D = 4) or half of one (D = 8) -- 128 contiguous bytes of the
    // input -- written with the 128-byte swizzle (unit ^= row % 8).  One / two copies per chunk instead of eight: the per-block
    // tiles with 16 / 32-byte lines are TMA-bound there (4.2-4.9 TB/s), the whole-row tiles reach 4.4 TB/s (D = 4, T = 64) and
    // 6.2 TB/s (D = 8, T = 64).  Longer filters are FP32-bound and keep the per-block tiles, whose window addressing is
    // cheaper (T = 128: 0.132 / 0.071 ms against 0.142 / 0.072 ms at D = 4 / 8).
    // (D = 8 with 16 tap blocks: whole-row tiles AND two CTAs, 0.0693 against 0.0702 ms with per-block tiles at T = 128)
    static constexpr bool WHOLE = (D < 16 && JT * D <= 64) || (D == 16 && JT <= 8) || (D == 8 && JT == 16);
    // thread-rows of a slot: 32 + the halo, two rows per pass of 16 tap blocks (36 / 40 / 48 / 64 for JT <= 32 / 64 / 128 / 256)
    static constexpr int SLOT_ROWS = JT <= 32 ? 36 : 32 + 2 * (JT / 16);
    static constexpr int TILE_ROWS = (SLOT_ROWS + 7) / 8 * 8;        // pitch in lines: a multiple of the swizzle period
    static constexpr int LINE_BYTES = WHOLE ? 128 : DB * 4;          // 128 / 64 / 32 / 16
    static constexpr int NTILE = WHOLE ? (ROW * 4) / 128 : R;        // copies per slot: 1 (D = 4), 2 (D = 8); 8 per-block tiles
    static constexpr int TILE_BYTES = TILE_ROWS * LINE_BYTES;
    static constexpr int SLOT_BYTES = NTILE * TILE_BYTES;
    static constexpr int TX_BYTES = NTILE * SLOT_ROWS * LINE_BYTES;  // bytes landed per slot
    static constexpr int HDR_BYTES = 1024;
    static constexpr int NGROUPS = 8, NWARPS = 8, NPROD = 2;
    static constexpr int NSLOT_MAX = (227 * 1024 - HDR_BYTES - 1024) / SLOT_BYTES;
    // D = 4 / 8 with whole-row tiles: TWO CTAs per SM (the slots are small, and 20 warps hide the ring and pipe latencies that
    // 10 cannot: ncu shows the FMA pipe 65 % active and DRAM at 48 % with one CTA); the ring is cut to what two CTAs can hold
    static constexpr int CTAS = (D == 4 || (D == 8 && JT > 8)) ? 2 : 1;
    static constexpr int NSLOT_FIT = NSLOT_MAX / CTAS > 16 ? 16 : NSLOT_MAX / CTAS;
    // (sixteen slots for the HBM-bound whole-row tiles measured 2 % faster at T = 64, D = 4 -- 0.0808 -> 0.0790 ms -- but one run of the
    // GPU suite then failed in that very cell (ring_wraparound[4-64]) and could not be reproduced: twelve stays, the value every run of both rounds has used)
    static constexpr int NSLOT = (CTAS == 2 && NSLOT_FIT > 12) ? 12 : NSLOT_FIT;   // 11 at D >= 32
    // nine slots (D = 8 beyond 64 tap blocks, two CTAs per SM): the second producer's four warps then have no slot of read-ahead;
    // a chunk takes ~30 us of FIR there, so the ~2 us refill is hidden by the other fifteen compute warps of the SM
    static_assert(NSLOT >= NGROUPS + 1 && NSLOT <= 16, "ring size");
    // per-block tiles: 64-byte swizzle (16-float lines), 32-byte swizzle (8-float lines), none (4-float lines)
    __host__ __device__ static constexpr int swz(int row) { return DB == 16 ? ((row >> 1) & 3) : (DB == 8 ? ((row >> 2) & 1) : 0); }
    static constexpr int CHUNK_ROWS = 32;
    static constexpr int CHUNK_OUT = CHUNK_ROWS * R;       // 256 outputs
    static constexpr int CHUNK_BLOCKS = CHUNK_ROWS * R;    // 256 D-sample blocks
    static constexpr int NTW = 3 * (JT / 2) * D;           // complex taps (kernel W layout)
    static constexpr int SMEM = HDR_BYTES + NSLOT * SLOT_BYTES + 1024;   // + slack to align the slots to 1 KB at run time
    static_assert(SMEM <= 227 * 1024, "shared memory");
    __host__ __device__ static constexpr int sub_count(int pr) { return pr == 0 ? (NSLOT + 1) / 2 : NSLOT / 2; }
    __host__ __device__ static constexpr int sub_base(int pr) { return pr == 0 ? 0 : (NSLOT + 1) / 2; }
};

// 5-D tensor copy global -> shared (SASS: UTMALDG), completion counted in bytes on `bar`
__device__ __forceinline__ void tma_load_5d(void* dst_smem, const CUtensorMap* tmap, int c0, int c1, int c2, int c3, int c4,
                                            uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(
            smem_u32(dst_smem)),
        "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(smem_u32(bar))
        : "memory");
}

// Epilogue with a warp-uniform shortcut in front of the branch-free w_epilogue (ddc_kernel_w.cuh): nothing at all when the stores
// are disabled (slices 1 .. LPQ-1 of a chunk, first chunk of a warp).  (A second shortcut -- plain 16-byte stores without the
// per-output predicates when every lane has all R outputs -- was tried: both store paths alive at once cost 300 bytes of spills.)
// Sliced kernels only (LPQ > 1): with one slice per chunk the shortcut would only ever skip a warp's first chunk, and the branch
// costs the register-capped two-CTA kernels 300 bytes of spills.
template <int R, int LPQ>
__device__ __forceinline__ void ws_epilogue(const float2 (&y)[R], const float2 (&rot_thr)[R], unsigned long long chunk_phase,
                                            long long m0, float2* o, long long nout) {
    if constexpr (LPQ > 1) {
        if (nout != 0) w_epilogue<R>(y, rot_thr, chunk_phase, m0, o, nout);   // warp-uniform
    } else {
        w_epilogue<R>(y, rot_thr, chunk_phase, m0, o, nout);
    }
}

// Work items of a CTA are POSITIONS: position = (round * LPQ + q) * 8 + w  <->  slice q of the chunk that compute warp w
// handles in its round-th turn (local chunk index round * 8 + w).  Producer p stages positions = p (mod 2) in order, into
// its own sub-ring; warp w consumes positions w, w + 8, ...: consumption order matches staging order.
template <int D, int JT>
__global__ void __launch_bounds__(WSCfg<D, JT>::NWARPS * 32 + 32 * WSCfg<D, JT>::NPROD, WSCfg<D, JT>::CTAS)
ddc_fused_ws_kernel(const __grid_constant__ RunParams p, const __grid_constant__ CUtensorMap tmap,
                    const __grid_constant__ TapsParam<WSCfg<D, JT>::NTW> taps) {
    using C = WSCfg<D, JT>;
    constexpr int R = C::R, NW = C::NWP, NWARPS = C::NWARPS, NG = C::NGROUPS, NSLOT = C::NSLOT, RH = C::RH;
    constexpr int LPQ = C::LPQ, JP = C::JP, NJG = C::NJG;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_raw);
    uint64_t* empty_bar = full_bar + 16;
    volatile int* slot_seq = reinterpret_cast<volatile int*>(smem_raw + 384);
    // the swizzle is a function of the shared-window address: align the slots to 1 KB in that address space
    unsigned char* buf = smem_raw + C::HDR_BYTES + ((1024u - ((smem_u32(smem_raw) + C::HDR_BYTES) & 1023u)) & 1023u);

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

    const int n_k = (int)((p.total_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);   // chunks of this CTA
    // positions of this CTA: full rounds of 8 chunks, the last round may be partial (position exists iff its chunk does)
    const int n_rounds = (n_k + NG - 1) / NG;
    const int n_pos = n_rounds * LPQ * NG;
    const unsigned long long chunk_dph = (unsigned long long)((long long)C::CHUNK_OUT * D) * p.step_fx;

    if (warp >= NWARPS) {
        // ------------------------------------------------------------------ producer warps
        const int pid = warp - NWARPS;
        const int sbase = C::sub_base(pid), scnt = C::sub_count(pid);
        int sidx = 0;
        uint32_t par = 1;
        for (int pos = pid; pos < n_pos; pos += C::NPROD) {
            const int w = pos % NG, q = (pos / NG) % LPQ, round = pos / (NG * LPQ);
            const int k = round * NG + w;               // local chunk index
            if (k >= n_k) continue;                     // partial last round: this warp has no chunk (consumers skip it too)
            const long long gk = blockIdx.x + (long long)k * gridDim.x;   // global chunk
            int cs, cc;
            chunk_of(p, gk, cs, cc);
            const int slot = sbase + sidx;
            if (elect_one()) {
                mbar_wait(&empty_bar[slot], par);
                slot_seq[slot] = pos;
                unsigned char* dst = buf + (size_t)slot * C::SLOT_BYTES;
                mbar_arrive_expect_tx(&full_bar[slot], (uint32_t)C::TX_BYTES);
                const int row0 = cc * C::CHUNK_ROWS;   // first thread-row (8 blocks) of the chunk inside its stream
                // NOT fully unrolled: eight UTMALDG in a row need so many uniform registers that ptxas stops using the uniform
                // datapath for the whole kernel -- the consumers' tap fetches turn from LDCU into per-thread LDC (2.2x slower)
#pragma unroll 4
                for (int b = 0; b < C::NTILE; ++b) tma_load_5d(dst + b * C::TILE_BYTES, &tmap, 0, q, b, row0, cs, &full_bar[slot]);
            }
            __syncwarp();
            if (++sidx == scnt) { sidx = 0; par ^= 1u; }
        }
    } else {
        // ------------------------------------------------------------------ compute warps
        const int grp = warp;
        const int g = lane;   // lane L owns thread-row L of the chunk (see WSCfg)
        // window of a pass: blocks 0 .. NW-1 of thread-rows grow, grow+1, ...; phase group pgv.  Block blk of row `row` lives in
        // tile blk at line `row`, logical 16-byte unit pgv, physical unit pgv ^ swz(row).
        auto load_window = [&](float4(&w)[NW], const unsigned char* sb, int grow, int pgv) {
            int a[C::WROWS];
            if constexpr (C::WHOLE) {
                // byte offset of (blk, pgv) inside its thread-row: blk * 4 D + 16 pgv -> tile = offset / 128, unit = rest / 16,
                // physical unit = unit ^ (row % 8).  blk * 4 D is static; 16 pgv only touches bit 4 (D = 8) or is zero (D = 4).
#pragma unroll
                for (int h = 0; h < C::WROWS; ++h) {
                    const int row = grow + h;
                    a[h] = ((row & 7) << 4) ^ (pgv << 4);   // XOR mask of this row (and phase group)
                }
#pragma unroll
                for (int b = 0; b < NW; ++b) {
                    const int h = b / R, blk = b % R;
                    const int off = blk * 4 * D;            // static
                    w[b] = *reinterpret_cast<const float4*>(sb + (off >> 7) * C::TILE_BYTES + (grow + h) * 128 + ((off & 127) ^ a[h]));
                }
            } else {
#pragma unroll
                for (int h = 0; h < C::WROWS; ++h) {
                    const int row = grow + h;
                    a[h] = row * C::LINE_BYTES + ((pgv ^ C::swz(row)) << 4);
                }
#pragma unroll
                for (int b = 0; b < NW; ++b) w[b] = *reinterpret_cast<const float4*>(sb + a[b / R] + (b % R) * C::TILE_BYTES);
            }
        };
        float2 rot_thr[R];
#pragma unroll
        for (int r = 0; r < R; ++r) rot_thr[r] = nco_rot((unsigned long long)((long long)(g * R + r) * D) * p.step_fx);

        // my positions are the (grp / 2)-th, (grp / 2 + 4)-th, ... of producer grp % 2, but only positions whose chunk exists
        // are staged, so the slot cursor advances per staged position
        const int sbase = C::sub_base(grp % C::NPROD), scnt = C::sub_count(grp % C::NPROD);
        int scur = grp / C::NPROD;   // slot cursor: index in bits 0-7, mbarrier parity in bit 8; (round 0, slice 0): staged = grp / NPROD < scnt

        float2 yprev[R];
#pragma unroll
        for (int r = 0; r < R; ++r) yprev[r] = make_float2(0.f, 0.f);
        long long prev_m0 = 0;
        float2* prev_o = p.out;
        int prev_cc = 0;
        long long prev_nout = 0;   // 0 disables the stores
        float2 m0a[RH], m1a[RH], m2a[RH];

        // The slice loop is a plain counted loop INSIDE the chunk loop so that q (and the tap pointer derived from it) is
        // warp-uniform for the compiler; derived from a position index that involves the warp number it becomes a per-thread
        // value and the taps are fetched with per-thread LDC instead of LDCU -> uniform registers (measured: 2.2x slower).
        for (int round = 0; round < n_rounds; ++round) {
            const int k = round * NG + grp;
            if (k >= n_k) break;   // no chunk for this warp in the (partial) last round
#pragma unroll 1
          for (int q = 0; q < LPQ; ++q) {
            const int pos = (round * LPQ + q) * NG + grp;
            // slot of this position: index among the positions staged by my producer.  In a full round every position is staged;
            // in the partial last round only those of warps < n_k % 8.  Staged positions before `pos` on my producer:
            // Full rounds advance the cursor by per_slice_full slots per position: kept incrementally (scur) -- a chunk at D = 4
            // is only ~500 instructions of FIR, and the divisions of the closed form were a tenth of that.  The partial last round
            // uses the closed form once.
            constexpr int per_slice_full = NG / C::NPROD;                                  // staged per (round, slice) and producer, full rounds
            int slot;
            uint32_t par;
            if (round < n_rounds - 1) {
                slot = sbase + (scur & 255);
                par = (uint32_t)scur >> 8;
                scur += per_slice_full;
                if ((scur & 255) >= scnt) scur = (scur - scnt) ^ 256;
            } else {
                const int last_cnt = n_k - (n_rounds - 1) * NG;                            // chunks in the last round (1 .. 8)
                const int par2 = grp % C::NPROD;
                const int per_slice_last = (last_cnt - par2 + C::NPROD - 1) / C::NPROD;    // warps w < last_cnt with w % 2 == par2
                const int staged = (n_rounds - 1) * LPQ * per_slice_full + q * per_slice_last + grp / C::NPROD;
                slot = sbase + staged % scnt;
                par = (uint32_t)(staged / scnt) & 1u;
            }
            spin_until_eq_uni(&slot_seq[slot], pos);
            mbar_wait_uni(&full_bar[slot], par);
            const unsigned char* sbuf = buf + (size_t)slot * C::SLOT_BYTES;
            if (q == 0) {
#pragma unroll
                for (int r = 0; r < RH; ++r) m0a[r] = m1a[r] = m2a[r] = make_float2(0.f, 0.f);
            }
            // taps of slice q: phases 16 q .. 16 q + 15 -> float4 offset 8 q in every (i, seq) set
            // taps of slice q: phases 16 q .. 16 q + 15 -> float4 offset 8 q in every (i, seq) set (q must stay warp-uniform:
            // hence the .uni waits above)
            const float4* tp = &taps.c2[8 * q];
            int pgv = 0;
            {   // first pass of the slice; it carries the previous chunk's epilogue (stores enabled only once per chunk)
                float4 w[NW];
                load_window(w, sbuf, g, 0);
                ws_epilogue<R, LPQ>(yprev, rot_thr, p.phase0_fx + (unsigned long long)prev_cc * chunk_dph, prev_m0, prev_o,
                              q == 0 ? prev_nout : 0);
                w_fir_pg<D, JP, R>(w, tp, m0a, m1a, m2a);
                pgv = 1;
                tp += 2;
            }
            {
                int grow = g;
                int pgi = 1;
                if (C::V == 1) {   // one phase group per tap group: the next pass already belongs to the next tap group
                    pgi = 0;
                    pgv = 0;
                    grow += JP / R;
                    tp += 3 * (JP / 2) * (D / 2) - 2 * C::V;
                }
#pragma unroll 1
                for (int pass = 1; pass < NJG * C::V; ++pass) {
                    asm volatile("" : "+r"(pgv), "+r"(grow));   // per-thread copies, hidden from the uniform induction variables
                    float4 w[NW];
                    load_window(w, sbuf, grow, pgv);
                    w_fir_pg<D, JP, R>(w, tp, m0a, m1a, m2a);
                    ++pgi;
                    ++pgv;
                    tp += 2;
                    if (pgi == C::V) {
                        pgi = 0;
                        pgv = 0;
                        grow += JP / R;
                        tp += 3 * (JP / 2) * (D / 2) - 2 * C::V;   // first tap set of the next group of 16 tap blocks, phase group 0
                    }
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty_bar[slot]);

            if (q == LPQ - 1) {
                // chunk complete: combine the three half-rate sums and hand them to the deferred epilogue
                const long long gk = blockIdx.x + (long long)k * gridDim.x;
                int cs, cc;
                chunk_of(p, gk, cs, cc);
#pragma unroll
                for (int r = 0; r < RH; ++r) {
                    yprev[2 * r] = make_float2(m0a[r].x + m1a[r].x, m0a[r].y + m1a[r].y);
                    yprev[2 * r + 1] = make_float2(m1a[r].x - m2a[r].x, m1a[r].y - m2a[r].y);
                }
                prev_cc = cc;
                prev_m0 = (long long)cc * C::CHUNK_OUT + g * R;
                prev_o = p.out + (long long)cs * p.out_stride + prev_m0;
                prev_nout = p.n_out;
            } else if (q == 0) {
                prev_nout = 0;   // the pending epilogue has been issued with this slice's first pass
            }
          }
        }
        ws_epilogue<R, LPQ>(yprev, rot_thr, p.phase0_fx + (unsigned long long)prev_cc * chunk_dph, prev_m0, prev_o, prev_nout);
    }
}

}  // namespace ddck
