// Probe: where does a 5-D TMA tile copy with a swizzle mode put each 16-byte chunk when the inner box dimension is 64
// bytes?  Findings on B200: SWIZZLE_128B pads every 64-byte box row to its own 128-byte line (half the tile is unused)
// and XORs the unit index with the line index; SWIZZLE_64B keeps 64-byte lines dense and XORs unit bits 4-5 with address
// bits 7-8 (this file now probes the latter).  (Design input for ddc_kernel_ws.cuh.)  Build: nvcc -arch=sm_100a -o tma_swizzle_probe tma_swizzle_probe.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void probe(const __grid_constant__ CUtensorMap tmap, float* out, int q, int bp, int row0, int dst_off) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    unsigned base = (unsigned)__cvta_generic_to_shared(smem);
    unsigned pad = (1024u - (base & 1023u)) & 1023u;
    unsigned char* tile = smem + pad + dst_off;
    if (threadIdx.x == 0) {
        unsigned b = (unsigned)__cvta_generic_to_shared(&bar);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
        asm volatile("fence.mbarrier_init.release.cluster;");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(1024));
        asm volatile(
            "cp.async.bulk.tensor.5d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(
                (unsigned)__cvta_generic_to_shared(tile)),
            "l"(&tmap), "r"(0), "r"(q), "r"(bp), "r"(row0), "r"(0), "r"(b)
            : "memory");
        asm volatile(
            "{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(b)
            : "memory");
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += blockDim.x) out[i] = reinterpret_cast<float*>(tile)[i];
}

int main() {
    const int D = 64, rows = 64;
    const long long n = (long long)rows * 8 * D;
    std::vector<float> h(n);
    for (long long i = 0; i < n; ++i) h[i] = (float)i;   // value = sample index
    float *d, *o;
    cudaMalloc(&d, n * 4);
    cudaMalloc(&o, 1024);
    cudaMemcpy(d, h.data(), n * 4, cudaMemcpyHostToDevice);
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres);
    PFN_encodeTiled enc = (PFN_encodeTiled)sym;
    CUtensorMap tm;
    const cuuint64_t gdim[5] = {16, (cuuint64_t)(D / 16), 8, (cuuint64_t)rows, 1};
    const cuuint64_t gstr[4] = {64, (cuuint64_t)D * 4, (cuuint64_t)D * 32, (cuuint64_t)n * 4};
    const cuuint32_t box[5] = {16, 1, 1, 16, 1};   // 16 thread-rows of one block index: 16 lines of 64 B
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, d, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode: %d\n", (int)r);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192);
    for (int dst_off = 0; dst_off <= 1024; dst_off += 1024) {
        const int q = 1, bp = 2, row0 = 8;
        probe<<<1, 128, 8192>>>(tm, o, q, bp, row0, dst_off);
        cudaError_t e = cudaDeviceSynchronize();
        printf("dst_off %d: %s\n", dst_off, cudaGetErrorString(e));
        float res[256];
        cudaMemcpy(res, o, 1024, cudaMemcpyDeviceToHost);
        // expected logical: line i (row row0+i), block 2bp+beta, float f: sample = ((row0+i)*8 + 2bp+beta)*D + 16q + f
        for (int line = 0; line < 8; ++line) {
            printf("line %d:", line);
            for (int c = 0; c < 8; ++c) {
                const long long v = (long long)res[line * 32 + c * 4];
                const long long blk = v / D, within = v % D;            // block index, offset in block
                const long long row = blk / 8, b = blk % 8;
                printf("  [r%lld b%lld f%lld]", row, b, within - 16 * q);
            }
            printf("\n");
        }
    }
    return 0;
}
