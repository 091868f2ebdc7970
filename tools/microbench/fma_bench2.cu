// Second round of micro-benchmarks: FFMA2 operand forms with NON-reused, distinct x registers (as in the real FIR loop).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)
__constant__ float2 ctaps[2048];
struct Res { unsigned long long cyc; };
#define ITERS 1024
#define NX 64

// v1: acc_r += bcast(x[k]) * tapUR[k]   (design A)  -- 4 chains, 64 distinct x registers, static taps
__global__ void k_bcast_ur(float* out, Res* res, const float* in) {
    float x[NX]; float2 acc[4];
#pragma unroll
    for (int i = 0; i < NX; ++i) x[i] = in[threadIdx.x + 32 * i];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] = make_float2(0.f, 0.f);
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            float2 t = ctaps[k];
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[r] = __ffma2_rn(make_float2(x[16 * r + k], x[16 * r + k]), t, acc[r]);
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc[0].x + acc[1].y + acc[2].x + acc[3].y;
    if (threadIdx.x == 0) res[blockIdx.x].cyc = t1 - t0;
}
// v1d: as v1 but taps indexed dynamically (LDCU in loop, 1 per 4 FFMA2)
__global__ void k_bcast_ur_dyn(float* out, Res* res, const float* in) {
    float x[NX]; float2 acc[4];
#pragma unroll
    for (int i = 0; i < NX; ++i) x[i] = in[threadIdx.x + 32 * i];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] = make_float2(0.f, 0.f);
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        const int base = (it & 15) * 16;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            float2 t = ctaps[base + k];
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[r] = __ffma2_rn(make_float2(x[16 * r + k], x[16 * r + k]), t, acc[r]);
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc[0].x + acc[1].y + acc[2].x + acc[3].y;
    if (threadIdx.x == 0) res[blockIdx.x].cyc = t1 - t0;
}
// v2: packed x pairs (no broadcast): acc_r(2 partial sums) += xpair[k] * tappairUR[k]   (design B)
__global__ void k_pair_ur(float* out, Res* res, const float* in) {
    float2 x[NX / 2]; float2 acc[8];
#pragma unroll
    for (int i = 0; i < NX / 2; ++i) x[i] = make_float2(in[threadIdx.x + 32 * i], in[threadIdx.x + 32 * i + 7]);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = make_float2(0.f, 0.f);
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float2 tre = ctaps[2 * k], tim = ctaps[2 * k + 1];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                acc[2 * r] = __ffma2_rn(x[8 * r + k], tre, acc[2 * r]);
                acc[2 * r + 1] = __ffma2_rn(x[8 * r + k], tim, acc[2 * r + 1]);
            }
        }
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) res[blockIdx.x].cyc = t1 - t0;
}
// v3: broadcast x, taps in vector registers (no UR)
__global__ void k_bcast_rr(float* out, Res* res, const float* in, const float2* tg) {
    float x[NX]; float2 acc[4]; float2 t[16];
#pragma unroll
    for (int i = 0; i < NX; ++i) x[i] = in[threadIdx.x + 32 * i];
#pragma unroll
    for (int i = 0; i < 16; ++i) t[i] = tg[i + (threadIdx.x & 1)];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] = make_float2(0.f, 0.f);
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[r] = __ffma2_rn(make_float2(x[16 * r + k], x[16 * r + k]), t[k], acc[r]);
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc[0].x + acc[1].y + acc[2].x + acc[3].y;
    if (threadIdx.x == 0) res[blockIdx.x].cyc = t1 - t0;
}
// v4: scalar FFMA, distinct x, UR taps: 8 chains (re, im per output)
__global__ void k_scalar_ur(float* out, Res* res, const float* in) {
    float x[NX]; float ar[4], ai[4];
#pragma unroll
    for (int i = 0; i < NX; ++i) x[i] = in[threadIdx.x + 32 * i];
#pragma unroll
    for (int i = 0; i < 4; ++i) { ar[i] = 0.f; ai[i] = 0.f; }
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            float2 t = ctaps[k];
#pragma unroll
            for (int r = 0; r < 4; ++r) { ar[r] = fmaf(x[16 * r + k], t.x, ar[r]); ai[r] = fmaf(x[16 * r + k], t.y, ai[r]); }
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = ar[0] + ai[1] + ar[2] + ai[3] + ar[1] + ai[0] + ar[3] + ai[2];
    if (threadIdx.x == 0) res[blockIdx.x].cyc = t1 - t0;
}
// v5: input-stationary order: the same x register used with 4 different taps for 4 outputs (x.reuse), static taps
__global__ void k_bcast_ur_xstat(float* out, Res* res, const float* in) {
    float x[NX]; float2 acc[4];
#pragma unroll
    for (int i = 0; i < NX; ++i) x[i] = in[threadIdx.x + 32 * i];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] = make_float2(0.f, 0.f);
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[r] = __ffma2_rn(make_float2(x[k + 16 * (it & 3)], x[k + 16 * (it & 3)]), ctaps[16 * r + k], acc[r]);
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc[0].x + acc[1].y + acc[2].x + acc[3].y;
    if (threadIdx.x == 0) res[blockIdx.x].cyc = t1 - t0;
}

template <typename F>
static void run(const char* name, F launch, int grid, int block, double fma_per_thread, Res* d_res) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    std::vector<Res> h(grid); CK(cudaMemcpy(h.data(), d_res, grid * sizeof(Res), cudaMemcpyDeviceToHost));
    double cyc = 0; for (auto& r : h) cyc += (double)r.cyc; cyc /= grid;
    double fma_per_clk_sm = fma_per_thread * block / cyc;
    printf("%-28s block=%4d  cyc=%10.0f  FMA/clk/SM=%7.2f (%5.1f%%)  wall=%8.3f ms  eff_clk=%5.0f MHz\n", name, block, cyc,
           fma_per_clk_sm, fma_per_clk_sm / 1.28, ms, cyc / (ms * 1e-3) / 1e6);
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int nsm = p.multiProcessorCount;
    float *d_out, *d_in; Res* d_res; float2* d_t;
    CK(cudaMalloc(&d_out, sizeof(float) * nsm * 1024)); CK(cudaMalloc(&d_res, sizeof(Res) * nsm));
    CK(cudaMalloc(&d_in, sizeof(float) * 32 * 128)); CK(cudaMemset(d_in, 0, sizeof(float) * 32 * 128));
    CK(cudaMalloc(&d_t, sizeof(float2) * 64)); CK(cudaMemset(d_t, 0, sizeof(float2) * 64));
    std::vector<float2> h(2048); for (int i = 0; i < 2048; ++i) h[i] = make_float2(1e-3f * i, -1e-3f * i);
    CK(cudaMemcpyToSymbol(ctaps, h.data(), sizeof(float2) * 2048));
    const double fma = (double)ITERS * 128;
    for (int b : {128, 256, 512}) {
        printf("--- %d warps/SM ---\n", b / 32);
        run("bcast_ur (design A)", [&] { k_bcast_ur<<<nsm, b>>>(d_out, d_res, d_in); }, nsm, b, fma, d_res);
        run("bcast_ur dyn taps (LDCU)", [&] { k_bcast_ur_dyn<<<nsm, b>>>(d_out, d_res, d_in); }, nsm, b, fma, d_res);
        run("pair_ur (design B)", [&] { k_pair_ur<<<nsm, b>>>(d_out, d_res, d_in); }, nsm, b, fma, d_res);
        run("bcast_rr (taps in regs)", [&] { k_bcast_rr<<<nsm, b>>>(d_out, d_res, d_in, d_t); }, nsm, b, fma, d_res);
        run("scalar_ur (FFMA)", [&] { k_scalar_ur<<<nsm, b>>>(d_out, d_res, d_in); }, nsm, b, fma, d_res);
        run("bcast_ur x-stationary", [&] { k_bcast_ur_xstat<<<nsm, b>>>(d_out, d_res, d_in); }, nsm, b, fma, d_res);
    }
    return 0;
}
