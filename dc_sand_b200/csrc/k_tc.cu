// Launcher of the tensor-core engine for packed 10-bit input (ddc_kernel_tc.cuh): builds the fp16 hi / lo B operand (the
// banded tap matrix) in the shared-memory image the kernel copies, keeps it in a small device-side ring keyed by
// (step, decimation, tap version), sizes the shared-memory pipeline and launches.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>

#include <cuda_fp16.h>

#include "ddc_host.h"
#include "ddc_kernel_tc.cuh"

using namespace ddck;

namespace ddch {
namespace {

struct TcGeom {
    int D, T, NS, row_s, tile_s, R, N, K, k16, n_groups, a_rows, a_pitch, a_stage, raw_bytes, raw_slot, b_bytes, n_a, n_raw;
    size_t smem;
    bool ok;
};

// geometry of one (T, D, NS): NS sub-streams = MMA rows of 8 NS samples = R = 8 NS / D outputs, N = 4 R accumulator columns
TcGeom tc_geometry(int T, int D, int NS, int force_na = 0, int force_nraw = 0) {
    TcGeom g{};
    g.D = D;
    g.T = T;
    g.NS = NS;
    g.row_s = 8 * NS;
    g.tile_s = 128 * g.row_s;
    g.R = g.row_s / D;
    g.N = 4 * g.R < 16 ? 16 : 4 * g.R;
    g.K = (g.row_s - D + T + 15) / 16 * 16;
    g.k16 = g.K / 16;
    g.n_groups = (g.tile_s + g.K - g.row_s) / 16;             // 16-sample groups a tile's windows touch
    g.a_rows = (2 * g.n_groups + NS - 1) / NS;
    g.a_pitch = 16 * (g.a_rows | 1);                          // odd number of 16-byte units: conflict-free unit stores
    g.a_stage = (NS * g.a_pitch + 127) & ~127;
    g.raw_bytes = (20 * g.n_groups + 15) & ~15;
    g.raw_slot = (g.raw_bytes + 127) & ~127;
    g.b_bytes = g.N * g.K * 2;
    g.ok = false;
    if (g.R < 1 || g.N > 256) return g;
    if (g.n_groups > TcShape<8>::UNP_CAP) return g;   // what the unpack warps of a team cover per tile
    const size_t cap = 227 * 1024;
    const size_t fixed = 1024 + (size_t)((g.b_bytes + 127) & ~127);
    // two A stages are enough for the MMA of one tile to overlap the unpack of the next (three when memory allows); everything
    // else goes to the raw ring: packed bytes in flight are what hides the HBM latency (up to 8 slots, the barrier arrays' size)
    for (int na = (force_na ? force_na : 3); na >= 2 && !g.ok; --na) {
        if (fixed + (size_t)na * g.a_stage + 2 * (size_t)g.raw_slot > cap) { if (force_na) break; continue; }
        int nr = (int)((cap - fixed - (size_t)na * g.a_stage) / (size_t)g.raw_slot);
        nr = std::min(nr, 8);
        if (force_nraw) nr = std::min(nr, force_nraw);
        // three stages want three raw slots: then three teams of unpack warps run (T = 256, D = 8: 0.0433 -> 0.0417 ms with three
        // stages and three slots instead of two and five); with fewer slots the deeper raw ring of two stages is the better trade
        if (na == 3 && nr < 3 && !force_na) continue;
        g.n_a = na;
        g.n_raw = nr;
        g.smem = fixed + (size_t)na * g.a_stage + (size_t)nr * g.raw_slot;
        g.ok = nr >= 2;
    }
    // descriptor fields are 14 bits of 16-byte units (the sums the MMA warp forms must not carry out of the address field)
    if (g.smem / 16 >= (1u << 14) || g.a_pitch / 16 >= (1 << 13)) g.ok = false;
    return g;
}

// sub-stream count for a (T, D): option "tc_ns" forces one; otherwise the widest rows that leave three pipeline stages
TcGeom tc_pick(const ddcb200* h, int T, int D) {
    if (h && (h->tc_ns || h->tc_na || h->tc_nraw)) return tc_geometry(T, D, h->tc_ns ? h->tc_ns : 16, h->tc_na, h->tc_nraw);
    TcGeom best{};
    for (int ns : {16, 8}) {   // 128-sample rows halve the operand re-reads of 64-sample rows; the latter fit longer filters
        if (ns == 16 && D < 8) continue;
        const TcGeom g = tc_geometry(T, D, ns);
        if (g.ok && !best.ok) best = g;
    }
    return best;
}

// B[(r, c), k] = part c of S * tap(k - D r) * e^{-j 2 pi step k}: c = 0 re_hi, 1 re_lo * 2^11, 2 im_hi, 3 im_lo * 2^11; image
// [K / 8][N][8] halves.  The rotation goes by the sample's position k inside the ROW, not inside the tap window, so that all
// the outputs of a row are left with one common rotation for the epilogue.
void build_b(const ddcb200* h, double step, const TcGeom& g, __half* out, float* inv_scale, float* lo_scale) {
    const int T = g.T;
    std::vector<double> rc(g.K), rs(g.K);
    const double fstep = step - std::floor(step);
    for (int k = 0; k < g.K; ++k) {
        double ph = fstep * (double)k;
        ph -= std::floor(ph);
        rc[k] = std::cos(-2.0 * M_PI * ph);
        rs[k] = std::sin(-2.0 * M_PI * ph);
    }
    double hmax = 0.0;
    for (int t = 0; t < T; ++t) hmax = std::max(hmax, std::fabs(h->taps[t] / h->taps_sum));
    int e = 0;
    if (hmax > 0.0) e = (int)std::floor(std::log2(32768.0 / hmax));
    if (std::ldexp(hmax, e) >= 32768.0) --e;
    e = std::max(-14, std::min(e, 40));
    const double S = std::ldexp(1.0, e);
    *inv_scale = (float)(512.0 / S);             // the kernel's unpack delivers v / 512
    *lo_scale = (float)(512.0 / (2048.0 * S));
    std::memset(out, 0, (size_t)g.b_bytes);
    auto split = [&](double v, __half& hi, __half& lo) {
        hi = __float2half_rn((float)(v * S));
        const double res = v * S - (double)__half2float(hi);
        lo = __float2half_rn((float)(res * 2048.0));
    };
    for (int r = 0; r < g.R; ++r)
        for (int t = 0; t < T; ++t) {
            const int k = t + g.D * r;
            const double hk = h->taps[T - 1 - t] / h->taps_sum;
            __half* col = out + ((size_t)(k / 8) * g.N) * 8 + (k % 8);
            __half hi, lo;
            split(hk * rc[k], hi, lo);
            col[(size_t)tc_col(g.R, r, 0) * 8] = hi;
            col[(size_t)tc_col(g.R, r, 1) * 8] = lo;
            split(hk * rs[k], hi, lo);
            col[(size_t)tc_col(g.R, r, 2) * 8] = hi;
            col[(size_t)tc_col(g.R, r, 3) * 8] = lo;
        }
}

template <int D, int NS, int NTEAM>
int launch_tc10_n(ddcb200* h, RunParams& p, const TcParams& tc, size_t smem, cudaStream_t st) {
    auto kern = ddc_tc10_kernel<D, NS, NTEAM>;
    static bool attr_set[64] = {};
    if (h->device < 64 && !attr_set[h->device]) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set[h->device] = true;
    }
    const long long grid = std::min<long long>(p.total_tiles, h->sm_count);
    kern<<<(unsigned)grid, TcShape<NS>::NTHREADS, smem, st>>>(p, tc);
    CUDA_TRY(cudaGetLastError());
    return DDCB200_OK;
}

// three teams of unpack warps where there are three sample stages for them, two otherwise (ddc_kernel_tc.cuh)
template <int D, int NS>
int launch_tc10_t(ddcb200* h, RunParams& p, const TcParams& tc, size_t smem, cudaStream_t st) {
    // (a team per stage AND per raw slot at most: a parity wait two phases ahead of its barrier would alias)
    return (tc.n_a >= 3 && tc.n_raw >= 3) ? launch_tc10_n<D, NS, 3>(h, p, tc, smem, st) : launch_tc10_n<D, NS, 2>(h, p, tc, smem, st);
}

}  // namespace

template <int D>
int launch_tc10_d(ddcb200* h, RunParams& p, const TcParams& tc, const TcGeom& g, cudaStream_t st) {
    switch (g.NS) {
        case 8: return launch_tc10_t<D, 8>(h, p, tc, g.smem, st);
        case 16:
            if constexpr (D >= 8) return launch_tc10_t<D, 16>(h, p, tc, g.smem, st);
            break;
    }
    return fail(DDCB200_EINVAL, "tensor engine: %d sub-streams at decimation %d are not built", g.NS, D);
}

bool tc10_supported(const ddcb200* h, int T, int D) {
    if (!(D == 4 || D == 8 || D == 16 || D == 32 || D == 64) || T < 1) return false;
    return tc_pick(h, T, D).ok;
}

// the pipeline the engine would run for (n_taps, D) with default options: see ddcb200_tensor_engine_geometry (include/ddcb200.h)
int tc10_describe(int T, int D, int32_t out[12]) {
    if (!(D == 4 || D == 8 || D == 16 || D == 32 || D == 64) || T < 1) return 0;
    const TcGeom g = tc_pick(nullptr, T, D);
    if (!g.ok) return 0;
    const int teams = (g.n_a >= 3 && g.n_raw >= 3) ? 3 : 2;   // launch_tc10_t
    const int32_t v[12] = {g.row_s, g.N, g.K, g.n_a, g.n_raw, teams, (int32_t)g.smem, g.n_groups,
                           TcShape<8>::UNP_CAP, g.a_pitch, g.raw_bytes, g.b_bytes};
    std::memcpy(out, v, sizeof(v));
    return 1;
}

int launch_tc10(ddcb200* h, RunParams& p, cudaStream_t st, double step, int D) {
    const int T = (int)h->taps.size();
    const TcGeom g = tc_pick(h, T, D);
    if (!g.ok) return fail(DDCB200_EINVAL, "tensor engine: %d taps at decimation %d do not fit shared memory", T, D);

    // ---- B operand: cached while (step, D, taps) repeat; a ring of device images so that a new key never overwrites one that
    // a queued kernel still reads
    const bool hit = h->tc_slot >= 0 && h->tc_step == step && h->tc_d == D && h->tc_ns_built == g.NS && h->tc_version == h->taps_version;
    if (!hit) {
        const int slot = (h->tc_slot + 1) % ddcb200::kTcRing;
        if (h->tc_ev[slot]) CUDA_TRY(cudaEventSynchronize(h->tc_ev[slot]));   // the last kernel that read this image is done
        else CUDA_TRY(cudaEventCreateWithFlags(&h->tc_ev[slot], cudaEventDisableTiming));
        if (!h->tc_up_ev) CUDA_TRY(cudaEventCreateWithFlags(&h->tc_up_ev, cudaEventDisableTiming));
        if ((size_t)g.b_bytes > h->tc_cap[slot]) {
            if (h->d_tc_b[slot]) cudaFree(h->d_tc_b[slot]);
            if (h->h_tc_b[slot]) cudaFreeHost(h->h_tc_b[slot]);
            h->d_tc_b[slot] = h->h_tc_b[slot] = nullptr;
            h->tc_cap[slot] = 0;
            const size_t cap = std::max<size_t>((size_t)g.b_bytes, 32768);
            CUDA_TRY(cudaMalloc(&h->d_tc_b[slot], cap));
            CUDA_TRY(cudaMallocHost(&h->h_tc_b[slot], cap));
            h->tc_cap[slot] = cap;
        }
        build_b(h, step, g, static_cast<__half*>(h->h_tc_b[slot]), &h->tc_inv_scale, &h->tc_lo_scale);
        CUDA_TRY(cudaMemcpyAsync(h->d_tc_b[slot], h->h_tc_b[slot], (size_t)g.b_bytes, cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaEventRecord(h->tc_up_ev, st));
        h->tc_slot = slot;
        h->tc_step = step;
        h->tc_d = D;
        h->tc_ns_built = g.NS;
        h->tc_version = h->taps_version;
    } else {
        CUDA_TRY(cudaStreamWaitEvent(st, h->tc_up_ev, 0));   // the image may have been uploaded on another stream
    }

    TcParams tc{};
    tc.b_mat = h->d_tc_b[h->tc_slot];
    tc.k16 = g.k16;
    tc.b_bytes = g.b_bytes;
    tc.a_rows = g.a_rows;
    tc.a_pitch = g.a_pitch;
    tc.a_stage_bytes = g.a_stage;
    tc.raw_bytes = g.raw_bytes;
    tc.raw_slot_bytes = g.raw_slot;
    tc.n_groups = g.n_groups;
    tc.n_a = g.n_a;
    tc.n_raw = g.n_raw;
    tc.inv_scale = h->tc_inv_scale;
    tc.lo_scale = h->tc_lo_scale;
    tc.unp_mul[0] = 1u << 10;
    tc.unp_mul[1] = 1u << 14;

    const long long tile_out = 128LL * g.R;
    p.tiles_per_stream = (p.n_out + tile_out - 1) / tile_out;
    p.total_tiles = p.tiles_per_stream * p.n_streams;
    p.n_taps = T;
    p.n_tap_blocks = g.k16;
    p.m_begin = 0;

    int rc = DDCB200_OK;
    switch (D) {
        case 4: rc = launch_tc10_d<4>(h, p, tc, g, st); break;
        case 8: rc = launch_tc10_d<8>(h, p, tc, g, st); break;
        case 16: rc = launch_tc10_d<16>(h, p, tc, g, st); break;
        case 32: rc = launch_tc10_d<32>(h, p, tc, g, st); break;
        default: rc = launch_tc10_d<64>(h, p, tc, g, st); break;
    }
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(h->tc_ev[h->tc_slot], st));
    h->launches++;
    char name[128];
    snprintf(name, sizeof(name), "tensor_fir_packed10<D%d,ROW%d,M128,N%d,K%d,A%d,RAW%d>", D, g.row_s, g.N, g.K, g.n_a, g.n_raw);
    h->last_variant = name;
    return DDCB200_OK;
}

}  // namespace ddch
