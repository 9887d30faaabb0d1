#include <cuda_runtime.h>
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__global__ void k(const float* __restrict__ t, const float2* __restrict__ v, float2* out) {
    float tr[8];
    for (int i = 0; i < 8; ++i) tr[i] = t[i + threadIdx.x];
    float2 acc[8];
    for (int i = 0; i < 8; ++i) acc[i] = v[i * 32 + threadIdx.x];
    float2 h = v[1000 + threadIdx.x];
#pragma unroll 1
    for (int it = 0; it < 100; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fma2(make_float2(tr[i], tr[i]), h, acc[i]);
        h.x += 1.0f;
    }
    for (int i = 0; i < 8; ++i) out[i * 32 + threadIdx.x] = acc[i];
}
