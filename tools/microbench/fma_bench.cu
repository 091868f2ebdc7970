// Micro-benchmarks that decide the inner-loop shape of the fused DDC kernel on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o fma_bench fma_bench.cu
// Each kernel reports FMA lane-ops per SM clock (peak = 128) from clock64 deltas, plus wall-time rates.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <string>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__constant__ float2 ctaps[2048];

struct Res { unsigned long long cyc; };

#define ITERS 2048

// (a) scalar FFMA, three register operands
__global__ void k_ffma_rrr(float* out, Res* res, float a, float b) {
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x * 0.001f + i;
    float m0 = a, m1 = b;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = fmaf(acc[i], (u & 1) ? m0 : m1, (i&1)? m1 : m0);
        }
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < 8; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) res[blockIdx.x].cyc = t1 - t0;
}
// (b) scalar FFMA with uniform-register tap operand: acc_i += x_i * tap
__global__ void k_ffma_ur(float* out, Res* res, float a) {
    float acc[8], x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc[i] = 0.f; x[i] = threadIdx.x * 0.001f + i * a; }
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            float2 t = ctaps[u];
#pragma unroll
            for (int i = 0; i < 4; ++i) { acc[2*i] = fmaf(x[i], t.x, acc[2*i]); acc[2*i+1] = fmaf(x[i], t.y, acc[2*i+1]); }
        }
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < 8; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) res[blockIdx.x].cyc = t1 - t0;
}
// (c) FFMA2, register pairs
__global__ void k_ffma2_rr(float* out, Res* res, float a, float b) {
    float2 acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = make_float2(threadIdx.x * 0.001f + i, i);
    float2 m0 = make_float2(a, b), m1 = make_float2(b, a);
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = __ffma2_rn(acc[i], (u & 1) ? m0 : m1, (i&1) ? m1 : m0);
        }
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) res[blockIdx.x].cyc = t1 - t0;
}
// (d) FFMA2 design A: acc(re,im)_r += bcast(x_r) * tap(UR pair); NACC accumulators, TAPS_PER_IT taps (static const offsets)
template <int NACC>
__global__ void k_ffma2_bcast_ur(float* out, Res* res, float a) {
    float2 acc[NACC]; float x[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { acc[i] = make_float2(0.f, 0.f); x[i] = threadIdx.x * 0.001f + i * a; }
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int u = 0; u < 64 / NACC; ++u) {
            float2 t = ctaps[u];
#pragma unroll
            for (int i = 0; i < NACC; ++i) acc[i] = __ffma2_rn(make_float2(x[i], x[i]), t, acc[i]);
        }
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < NACC; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) res[blockIdx.x].cyc = t1 - t0;
}
// (e) as (d) but taps change every iteration (dynamic uniform index -> LDCU in the loop). NACC FFMA2 per tap.
template <int NACC>
__global__ void k_ffma2_ldcu(float* out, Res* res, float a) {
    float2 acc[NACC]; float x[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { acc[i] = make_float2(0.f, 0.f); x[i] = threadIdx.x * 0.001f + i * a; }
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        const int base = (it & 15) * (64 / NACC);
#pragma unroll
        for (int u = 0; u < 64 / NACC; ++u) {
            float2 t = ctaps[base + u];
#pragma unroll
            for (int i = 0; i < NACC; ++i) acc[i] = __ffma2_rn(make_float2(x[i], x[i]), t, acc[i]);
        }
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < NACC; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) res[blockIdx.x].cyc = t1 - t0;
}
// (e2) as (e) but two taps per LDCU.128
template <int NACC>
__global__ void k_ffma2_ldcu128(float* out, Res* res, float a) {
    float2 acc[NACC]; float x[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) { acc[i] = make_float2(0.f, 0.f); x[i] = threadIdx.x * 0.001f + i * a; }
    const float4* t4 = reinterpret_cast<const float4*>(ctaps);
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        const int base = (it & 15) * (32 / NACC);
#pragma unroll
        for (int u = 0; u < 32 / NACC; ++u) {
            float4 t = t4[base + u];
#pragma unroll
            for (int i = 0; i < NACC; ++i) acc[i] = __ffma2_rn(make_float2(x[i], x[i]), make_float2(t.x, t.y), acc[i]);
#pragma unroll
            for (int i = 0; i < NACC; ++i) acc[i] = __ffma2_rn(make_float2(x[i], x[i]), make_float2(t.z, t.w), acc[i]);
        }
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < NACC; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) res[blockIdx.x].cyc = t1 - t0;
}
// (f) as (e, NACC=4) plus one LDS.128 of fresh samples per LDSPER taps (x rotates) — smem issue + bandwidth interplay
template <int NACC, int TAPS_PER_LDS>
__global__ void k_ffma2_ldcu_lds(float* out, Res* res, float a) {
    extern __shared__ float4 sm[];
    for (int i = threadIdx.x; i < 17 * (int)blockDim.x + 16; i += blockDim.x) sm[i] = make_float4(a * i, a, 1.f, 2.f);
    __syncthreads();
    float2 acc[NACC]; float4 xv = sm[threadIdx.x];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = make_float2(0.f, 0.f);
    const float4* p = sm + threadIdx.x * 17;   // pitch 17 chunks: conflict-free across a quarter warp
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        const int base = (it & 15) * 16;
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            if (u % TAPS_PER_LDS == 0) xv = p[(it + u) & 15];
            float2 t = ctaps[base + u];
            acc[0] = __ffma2_rn(make_float2(xv.x, xv.x), t, acc[0]);
            acc[1] = __ffma2_rn(make_float2(xv.y, xv.y), t, acc[1]);
            acc[2] = __ffma2_rn(make_float2(xv.z, xv.z), t, acc[2]);
            acc[3] = __ffma2_rn(make_float2(xv.w, xv.w), t, acc[3]);
        }
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < NACC; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) res[blockIdx.x].cyc = t1 - t0;
}

template <typename F>
static void run(const char* name, F launch, int grid, int block, double fma_per_thread, float* d_out, Res* d_res) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    std::vector<Res> h(grid); CK(cudaMemcpy(h.data(), d_res, grid * sizeof(Res), cudaMemcpyDeviceToHost));
    double cyc = 0; for (auto& r : h) cyc += (double)r.cyc; cyc /= grid;
    int dev; cudaGetDevice(&dev); cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
    double ctas_per_sm = (double)grid / p.multiProcessorCount;
    double fma_per_clk_sm = fma_per_thread * block * ctas_per_sm / cyc;   // valid when all CTAs co-resident
    double tfma = fma_per_thread * block * grid / (ms * 1e-3) / 1e12;
    printf("%-34s grid=%5d block=%4d  cyc=%10.0f  FMA/clk/SM=%7.2f  wall=%8.3f ms  TFMA/s=%6.2f (=%6.2f TFLOP/s)  eff_clk=%5.0f MHz\n",
           name, grid, block, cyc, fma_per_clk_sm, ms, tfma, 2 * tfma, cyc / (ms * 1e-3) / 1e6);
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    printf("device %s SMs=%d clock=%d kHz\n", p.name, p.multiProcessorCount, p.clockRate);
    int nsm = p.multiProcessorCount;
    float* d_out; Res* d_res;
    CK(cudaMalloc(&d_out, sizeof(float) * nsm * 8 * 1024)); CK(cudaMalloc(&d_res, sizeof(Res) * nsm * 8));
    std::vector<float2> h(2048); for (int i = 0; i < 2048; ++i) h[i] = make_float2(1e-3f * i, -1e-3f * i);
    CK(cudaMemcpyToSymbol(ctaps, h.data(), sizeof(float2) * 2048));
    CK(cudaFuncSetAttribute(k_ffma2_ldcu_lds<4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200000));
    CK(cudaFuncSetAttribute(k_ffma2_ldcu_lds<4, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200000));
    CK(cudaFuncSetAttribute(k_ffma2_ldcu_lds<4, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200000));
    const int blocks[] = {128, 256, 512, 1024};
    for (int b : blocks) {
        int grid = nsm;
        printf("--- %d warps/SM ---\n", b / 32);
        run("ffma_rrr", [&] { k_ffma_rrr<<<grid, b>>>(d_out, d_res, 1.0001f, 0.9999f); }, grid, b, (double)ITERS * 64, d_out, d_res);
        run("ffma_ur", [&] { k_ffma_ur<<<grid, b>>>(d_out, d_res, 0.5f); }, grid, b, (double)ITERS * 64, d_out, d_res);
        run("ffma2_rr", [&] { k_ffma2_rr<<<grid, b>>>(d_out, d_res, 1.0001f, 0.9999f); }, grid, b, (double)ITERS * 128, d_out, d_res);
        run("ffma2_bcast_ur<4>", [&] { k_ffma2_bcast_ur<4><<<grid, b>>>(d_out, d_res, 0.5f); }, grid, b, (double)ITERS * 128, d_out, d_res);
        run("ffma2_bcast_ur<8>", [&] { k_ffma2_bcast_ur<8><<<grid, b>>>(d_out, d_res, 0.5f); }, grid, b, (double)ITERS * 128, d_out, d_res);
        run("ffma2_ldcu<1> (1 tap/FFMA2)", [&] { k_ffma2_ldcu<1><<<grid, b>>>(d_out, d_res, 0.5f); }, grid, b, (double)ITERS * 128, d_out, d_res);
        run("ffma2_ldcu<2>", [&] { k_ffma2_ldcu<2><<<grid, b>>>(d_out, d_res, 0.5f); }, grid, b, (double)ITERS * 128, d_out, d_res);
        run("ffma2_ldcu<4>", [&] { k_ffma2_ldcu<4><<<grid, b>>>(d_out, d_res, 0.5f); }, grid, b, (double)ITERS * 128, d_out, d_res);
        run("ffma2_ldcu128<1> (2 taps/LDCU.128, 2 FFMA2)", [&] { k_ffma2_ldcu128<1><<<grid, b>>>(d_out, d_res, 0.5f); }, grid, b, (double)ITERS * 128, d_out, d_res);
        run("ffma2_ldcu128<2> (4 FFMA2 per LDCU.128)", [&] { k_ffma2_ldcu128<2><<<grid, b>>>(d_out, d_res, 0.5f); }, grid, b, (double)ITERS * 128, d_out, d_res);
        run("ffma2_ldcu128<4> (8 FFMA2 per LDCU.128)", [&] { k_ffma2_ldcu128<4><<<grid, b>>>(d_out, d_res, 0.5f); }, grid, b, (double)ITERS * 128, d_out, d_res);
        run("ffma2_ldcu<8>", [&] { k_ffma2_ldcu<8><<<grid, b>>>(d_out, d_res, 0.5f); }, grid, b, (double)ITERS * 128, d_out, d_res);
        if (b <= 512) {  // 17*b+16 float4 of shared memory
            run("ffma2_ldcu_lds<4, lds/16taps>", [&] { k_ffma2_ldcu_lds<4, 16><<<grid, b, (17 * b + 16) * 16>>>(d_out, d_res, 0.5f); }, grid, b, (double)ITERS * 128, d_out, d_res);
            run("ffma2_ldcu_lds<4, lds/4taps>", [&] { k_ffma2_ldcu_lds<4, 4><<<grid, b, (17 * b + 16) * 16>>>(d_out, d_res, 0.5f); }, grid, b, (double)ITERS * 128, d_out, d_res);
            run("ffma2_ldcu_lds<4, lds/1tap>", [&] { k_ffma2_ldcu_lds<4, 1><<<grid, b, (17 * b + 16) * 16>>>(d_out, d_res, 0.5f); }, grid, b, (double)ITERS * 128, d_out, d_res);
        }
    }
    return 0;
}
