// Fourth round: FFMA2 stream whose sample registers are refreshed from shared memory with LDS.32 / LDS.64 / LDS.128
// (same bytes per FFMA2), as in the FIR inner loop: 64 FFMA2 (16 taps x 4 outputs) consume NL loaded floats.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)
__constant__ float2 ctaps[2048];
struct Res { unsigned long long cyc; };
#define ITERS 1024
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// W: load width in floats (1, 2, 4); NL: floats refreshed per 64 FFMA2 (0, 8, 16)
template <int W, int NL>
__global__ void k_ld(float* out, Res* res, const float* in) {
    extern __shared__ float4 sm[];
    for (int i = threadIdx.x; i < 17 * (int)blockDim.x + 64; i += blockDim.x) sm[i] = make_float4(i * 1e-6f, 1.f, 2.f, 3.f);
    __syncthreads();
    float x[64]; float2 acc[4];
#pragma unroll
    for (int i = 0; i < 64; ++i) x[i] = in[threadIdx.x + 32 * i];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] = make_float2(0.f, 0.f);
    const uint32_t base = s32(sm + threadIdx.x * 17);
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        const uint32_t a = base + ((it & 3) << 6);
        // refresh NL of the 64 sample registers (those of "block" it & 3 would be the real pattern; any 16 do)
        if (NL > 0) {
#pragma unroll
            for (int e = 0; e < NL; e += W) {
                if (W == 4) asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x[e]), "=f"(x[e+1]), "=f"(x[e+2]), "=f"(x[e+3]) : "r"(a + e * 4));
                if (W == 2) asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(x[e]), "=f"(x[e+1]) : "r"(a + e * 4 + ((e & 2) ? 136 : 0)));
                if (W == 1) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x[e]) : "r"(a + e * 4 + (e & 3) * 132));
            }
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            float2 t = ctaps[k];
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[r] = __ffma2_rn(make_float2(x[16 * ((r + 1) & 3) + k], x[16 * ((r + 1) & 3) + k]), t, acc[r]);
        }
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc[0].x + acc[1].y + acc[2].x + acc[3].y;
    if (threadIdx.x == 0) res[blockIdx.x].cyc = t1 - t0;
}
template <typename F>
static void run(const char* name, F launch, int grid, int block, double fma_per_thread, Res* d_res) {
    launch(); CK(cudaDeviceSynchronize()); launch(); CK(cudaDeviceSynchronize());
    std::vector<Res> h(grid); CK(cudaMemcpy(h.data(), d_res, grid * sizeof(Res), cudaMemcpyDeviceToHost));
    double cyc = 0; for (auto& r : h) cyc += (double)r.cyc; cyc /= grid;
    double f = fma_per_thread * block / cyc;
    printf("%-40s block=%4d  cyc=%10.0f  FMA/clk/SM=%7.2f (%5.1f%%)\n", name, block, cyc, f, f / 1.28);
}
#define RUN(W, NL, label) { auto kf = k_ld<W, NL>; CK(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, 200000)); \
    run(label, [&] { kf<<<nsm, b, (17 * b + 64) * 16>>>(d_out, d_res, d_in); }, nsm, b, fma, d_res); }
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
        RUN(4, 0, "no loads");
        RUN(4, 16, "16 floats / 64 FFMA2 via 4 x LDS.128");
        RUN(2, 16, "16 floats / 64 FFMA2 via 8 x LDS.64");
        RUN(1, 16, "16 floats / 64 FFMA2 via 16 x LDS.32");
        RUN(4, 8, " 8 floats / 64 FFMA2 via 2 x LDS.128");
        RUN(2, 8, " 8 floats / 64 FFMA2 via 4 x LDS.64");
        RUN(1, 8, " 8 floats / 64 FFMA2 via 8 x LDS.32");
    }
    return 0;
}
