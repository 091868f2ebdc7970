// Throughput of the tensor engine's 10-bit -> fp16 unpack (ddc_kernel_tc.cuh: tc_unpack16) on an otherwise idle SM, and of its
// instruction classes one at a time: how many 16-sample groups per clock one SM sustains with W warps, registers only
// (mode 0), with the five LDS.32 per group (mode 1) and with LDS + the two STS.128 (mode 2); modes 3 .. 7 repeat one
// instruction class (PRMT, LOP3, IMAD.WIDE, LEA.HI, HFMA2 on a denormal input) 8 x per group.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)
struct P { uint32_t mul[2]; uint32_t sel; uint32_t msk; };
#define ITERS 2048
__device__ __forceinline__ void unpack16(const uint32_t (&rw)[5], uint32_t (&h)[8], const P& p) {
    const __half2 k15 = __floats2half2_rn(32768.f, 32768.f), m1 = __floats2half2_rn(-1.f, -1.f);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const int bo = (20 * q) >> 3, wi = bo >> 2, j0 = bo & 3, s2 = (20 * q) & 7;
        const int j2 = j0 + 2 > 7 ? 7 : j0 + 2;
        const uint32_t sel = (uint32_t)((j0 << 12) | ((j0 + 1) << 8) | (j2 << 4) | j2);
        const uint32_t src = __byte_perm(rw[wi], rw[wi + 1 > 4 ? 4 : wi + 1], sel);
        uint32_t fm;
        asm("lop3.b32 %0, %1, %2, %3, 0x28;" : "=r"(fm) : "r"(src), "r"(0x80200000u >> s2), "r"(0xFFFFF000u >> s2));
        uint32_t plo, phi;
        asm("{\n.reg .b64 t;\nmul.wide.u32 t, %2, %3;\nmov.b64 {%0, %1}, t;\n}" : "=r"(plo), "=r"(phi) : "r"(fm), "r"(p.mul[s2 >> 2]));
        const uint32_t w = phi + (plo >> 6);
        const __half2 hv = __hfma2(*reinterpret_cast<const __half2*>(&w), k15, m1);
        h[q] = *reinterpret_cast<const uint32_t*>(&hv);
    }
}
template <int MODE>
__global__ void k(uint32_t* out, unsigned long long* cyc, const __grid_constant__ P p) {
    extern __shared__ __align__(16) unsigned char sm[];
    const int tid = threadIdx.x;
    for (int i = tid; i < 40960 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = i * 2654435761u;
    __syncthreads();
    uint32_t rw[5], acc[8] = {};
#pragma unroll
    for (int i = 0; i < 5; ++i) rw[i] = tid * 7919u + i * 104729u;
    const unsigned char* raw = sm + 20 * tid;            // 20-byte stride: conflict-free LDS.32
    unsigned char* dst = sm + 65536 + 32 * tid;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
        uint32_t h[8];
        if (MODE == 1 || MODE == 2) {
#pragma unroll
            for (int i = 0; i < 5; ++i) rw[i] = reinterpret_cast<const volatile uint32_t*>(raw + ((it & 1) ? 20480 : 0))[i];
        }
        if (MODE <= 2) {
            unpack16(rw, h, p);
        } else {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                uint32_t v = rw[q % 5] + q;
                if (MODE == 3) v = __byte_perm(v, rw[(q + 1) % 5], p.sel);
                if (MODE == 4) asm("lop3.b32 %0, %1, %2, %3, 0x28;" : "=r"(v) : "r"(v), "r"(p.msk), "r"(0xFFFFF000u));
                if (MODE == 5) { uint32_t lo, hi; asm("{\n.reg .b64 t;\nmul.wide.u32 t, %2, %3;\nmov.b64 {%0, %1}, t;\n}" : "=r"(lo), "=r"(hi) : "r"(v), "r"(p.mul[q & 1])); v = lo ^ hi; }
                if (MODE == 6) asm("shf.r.wrap.b32 %0, %1, %2, 6;" : "=r"(v) : "r"(v), "r"(rw[(q + 2) % 5]));
                if (MODE == 7) { const __half2 hv = __hfma2(*reinterpret_cast<const __half2*>(&v), __floats2half2_rn(32768.f, 32768.f), __floats2half2_rn(-1.f, -1.f)); v = *reinterpret_cast<const uint32_t*>(&hv) & 0x03ff03ffu; }
                h[q] = v;
            }
        }
        if (MODE == 2) {
            const uint32_t a = (uint32_t)__cvta_generic_to_shared(dst);
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(a + 16), "r"(h[4]), "r"(h[5]), "r"(h[6]), "r"(h[7]) : "memory");
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] ^= h[q];
        if (MODE != 1 && MODE != 2) {
#pragma unroll
            for (int i = 0; i < 5; ++i) rw[i] += acc[i] & 1u;   // keep the chain data-dependent but cheap
        }
    }
    long long t1 = clock64();
    uint32_t r = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) r ^= acc[q];
    out[blockIdx.x * blockDim.x + tid] = r;
    if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int MODE>
void run(const char* name, int warps, uint32_t* d_out, unsigned long long* d_cyc, int nsm, P p) {
    auto kf = k<MODE>;
    CK(cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, 200000));
    for (int r = 0; r < 2; ++r) { kf<<<nsm, warps * 32, 65536 + 32 * warps * 32 + 64>>>(d_out, d_cyc, p); CK(cudaDeviceSynchronize()); }
    std::vector<unsigned long long> h(nsm);
    CK(cudaMemcpy(h.data(), d_cyc, nsm * 8, cudaMemcpyDeviceToHost));
    double c = 0; for (auto v : h) c += (double)v; c /= nsm;
    const double groups = (double)ITERS * warps * 32;
    printf("%-34s warps=%2d  cycles=%9.0f  samples/clk/SM=%7.2f  cycles per warp-group=%6.2f\n", name, warps, c, 16.0 * groups / c, c / ITERS / warps * 4);
}
int main() {
    int nsm; CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
    uint32_t* d_out; unsigned long long* d_cyc;
    CK(cudaMalloc(&d_out, nsm * 1024 * 4)); CK(cudaMalloc(&d_cyc, nsm * 8));
    P p{{1u << 10, 1u << 14}, 0x2344u, 0x80200000u};
    for (int w : {4, 8, 12, 16, 24}) {
        run<0>("unpack16 registers only", w, d_out, d_cyc, nsm, p);
        run<1>("unpack16 + 5 LDS.32", w, d_out, d_cyc, nsm, p);
        run<2>("unpack16 + 5 LDS.32 + 2 STS.128", w, d_out, d_cyc, nsm, p);
    }
    for (int w : {4, 12}) {
        run<3>("8 PRMT", w, d_out, d_cyc, nsm, p);
        run<4>("8 LOP3", w, d_out, d_cyc, nsm, p);
        run<5>("8 IMAD.WIDE.U32 (+8 LOP3)", w, d_out, d_cyc, nsm, p);
        run<6>("8 SHF", w, d_out, d_cyc, nsm, p);
        run<7>("8 HFMA2 denormal in (+8 LOP3)", w, d_out, d_cyc, nsm, p);
    }
    return 0;
}
