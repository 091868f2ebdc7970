// Third round: does a non-FFMA2 vector instruction (ALU op, LDS.128) steal FMA-pipe time when interleaved with FFMA2?
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)
__constant__ float2 ctaps[2048];
struct Res { unsigned long long cyc; };
#define ITERS 1024

// EXTRA: 0 none, 1 = one IADD3-like per PER FFMA2, 2 = one LDS.128 per PER FFMA2 (result consumed one iteration later), 3 = LOP3
template <int EXTRA, int PER>
__global__ void k_mix(float* out, Res* res, const float* in, int salt) {
    extern __shared__ float4 sm[];
    for (int i = threadIdx.x; i < 17 * (int)blockDim.x + 64; i += blockDim.x) sm[i] = make_float4(i, 1.f, 2.f, 3.f);
    __syncthreads();
    float x[64]; float2 acc[4];
#pragma unroll
    for (int i = 0; i < 64; ++i) x[i] = in[threadIdx.x + 32 * i];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] = make_float2(0.f, 0.f);
    int ia[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) ia[i] = threadIdx.x + i * salt;
    float4 ld[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) ld[i] = make_float4(0, 0, 0, 0);
    const float4* sp = sm + threadIdx.x * 17;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        int e = 0;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            float2 t = ctaps[k];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                acc[r] = __ffma2_rn(make_float2(x[16 * r + k], x[16 * r + k]), t, acc[r]);
                if (EXTRA && ((k * 4 + r) % PER == 0)) {
                    if (EXTRA == 1) { ia[e & 7] = ia[e & 7] + ia[(e + 3) & 7] + salt; }
                    if (EXTRA == 3) { ia[e & 7] = (ia[e & 7] ^ ia[(e + 3) & 7]) & salt; }
                    if (EXTRA == 2) { float4 v = sp[(e + it) & 15]; ia[e & 7] ^= __float_as_int(v.x) ^ __float_as_int(v.w); }
                    if (EXTRA == 4) { float v = reinterpret_cast<const float*>(sp)[(e + it) & 63]; ia[e & 7] ^= __float_as_int(v); }
                    if (EXTRA == 5) { float2 v = reinterpret_cast<const float2*>(sp)[(e + it) & 31]; ia[e & 7] ^= __float_as_int(v.x) ^ __float_as_int(v.y); }
                    if (EXTRA == 6) { reinterpret_cast<float*>(sm)[threadIdx.x + 32 * ((e + it) & 63)] = x[e & 63]; }
                    if (EXTRA == 8) { ia[e & 7] ^= ia[(e + 1) & 7]; ia[(e + 2) & 7] ^= ia[(e + 3) & 7]; }
                    ++e;
                }
            }
        }
        if (EXTRA == 2) { x[it & 63 ? 0 : 1] += ld[0].x * 0.f; }
    }
    long long t1 = clock64();
    int s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += ia[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc[0].x + acc[1].y + acc[2].x + acc[3].y + s + ld[0].x + ld[1].y + ld[2].z + ld[3].w;
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
    double f = fma_per_thread * block / cyc;
    printf("%-34s block=%4d  cyc=%10.0f  FMA/clk/SM=%7.2f (%5.1f%%)\n", name, block, cyc, f, f / 1.28);
}
#define RUN(E, P, label) { auto kf = k_mix<E, P>; CK(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, 200000)); \
    run(label, [&] { kf<<<nsm, b, (17 * b + 64) * 16>>>(d_out, d_res, d_in, 3); }, nsm, b, fma, d_res); }
int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int nsm = p.multiProcessorCount;
    float *d_out, *d_in; Res* d_res;
    CK(cudaMalloc(&d_out, sizeof(float) * nsm * 1024)); CK(cudaMalloc(&d_res, sizeof(Res) * nsm));
    CK(cudaMalloc(&d_in, sizeof(float) * 32 * 128)); CK(cudaMemset(d_in, 0, sizeof(float) * 32 * 128));
    std::vector<float2> h(2048); for (int i = 0; i < 2048; ++i) h[i] = make_float2(1e-3f * i, -1e-3f * i);
    CK(cudaMemcpyToSymbol(ctaps, h.data(), sizeof(float2) * 2048));
    const double fma = (double)ITERS * 128;
    for (int b : {256, 512}) {
        printf("--- %d warps/SM ---\n", b / 32);
        RUN(0, 1, "FFMA2 only");
        RUN(1, 8, "+1 IADD3 per 8 FFMA2");
        RUN(1, 4, "+1 IADD3 per 4 FFMA2");
        RUN(1, 2, "+1 IADD3 per 2 FFMA2");
        RUN(1, 1, "+1 IADD3 per 1 FFMA2");
        RUN(3, 4, "+1 LOP3 per 4 FFMA2");
        RUN(3, 1, "+1 LOP3 per 1 FFMA2");
        RUN(4, 16, "+1 (LDS.32+LOP3) per 16 FFMA2");
        RUN(4, 4, "+1 (LDS.32+LOP3) per 4 FFMA2");
        RUN(5, 16, "+1 (LDS.64+2LOP3) per 16 FFMA2");
        RUN(5, 4, "+1 (LDS.64+2LOP3) per 4 FFMA2");
        RUN(6, 4, "+1 STS.32 per 4 FFMA2");
        RUN(6, 1, "+1 STS.32 per 1 FFMA2");
        RUN(8, 4, "+2 LOP3 per 4 FFMA2");
        RUN(2, 32, "+1 (LDS.128+2LOP3) per 32 FFMA2");
        RUN(2, 16, "+1 (LDS.128+2LOP3) per 16 FFMA2");
        RUN(2, 8, "+1 LDS.128 per 8 FFMA2");
        RUN(2, 4, "+1 LDS.128 per 4 FFMA2");
    }
    return 0;
}
