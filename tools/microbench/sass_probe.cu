#include <cuda_runtime.h>
__constant__ float2 ctaps[512];
// A: (re,im) pairing with broadcast x
extern "C" __global__ void probeA(const float* __restrict__ x, float2* __restrict__ y) {
    float2 acc0 = make_float2(0.f, 0.f), acc1 = acc0;
    float x0 = x[threadIdx.x], x1 = x[threadIdx.x + 32];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        acc0 = __ffma2_rn(ctaps[k], make_float2(x0, x0), acc0);
        acc1 = __ffma2_rn(ctaps[k], make_float2(x1, x1), acc1);
        x0 += 1.f; x1 += 1.f;
    }
    y[threadIdx.x] = make_float2(acc0.x + acc1.x, acc0.y + acc1.y);
}
// B: natural pairs, taps from constant
extern "C" __global__ void probeB(const float2* __restrict__ x, float2* __restrict__ y) {
    float2 acc0 = make_float2(0.f, 0.f), acc1 = acc0;
    float2 x0 = x[threadIdx.x], x1 = x[threadIdx.x + 32];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        acc0 = __ffma2_rn(ctaps[k], x0, acc0);
        acc1 = __ffma2_rn(ctaps[k + 8], x1, acc1);
    }
    y[threadIdx.x] = make_float2(acc0.x + acc1.x, acc0.y + acc1.y);
}
// C: taps from smem broadcast
extern "C" __global__ void probeC(const float2* __restrict__ x, float2* __restrict__ y, const float4* __restrict__ t) {
    __shared__ float4 st[64];
    st[threadIdx.x & 63] = t[threadIdx.x & 63];
    __syncthreads();
    float2 acc0 = make_float2(0.f, 0.f), acc1 = acc0;
    float2 x0 = x[threadIdx.x], x1 = x[threadIdx.x + 32];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        float4 c = st[k];
        acc0 = __ffma2_rn(make_float2(c.x, c.y), x0, acc0);
        acc1 = __ffma2_rn(make_float2(c.z, c.w), x0, acc1);
        acc0 = __ffma2_rn(make_float2(c.x, c.y), x1, acc0);
        acc1 = __ffma2_rn(make_float2(c.z, c.w), x1, acc1);
    }
    y[threadIdx.x] = make_float2(acc0.x + acc1.x, acc0.y + acc1.y);
}
// D: scalar FFMA with const operand
extern "C" __global__ void probeD(const float* __restrict__ x, float2* __restrict__ y) {
    float ar = 0.f, ai = 0.f;
    float x0 = x[threadIdx.x];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        ar = fmaf(ctaps[k].x, x0, ar);
        ai = fmaf(ctaps[k].y, x0, ai);
        x0 += 1.f;
    }
    y[threadIdx.x] = make_float2(ar, ai);
}
