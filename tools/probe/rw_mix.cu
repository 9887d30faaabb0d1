// Development probe: achievable HBM bandwidth for the read:write mixes of the pyramid launches (1:1 copy, 1:2 blur+DoG, 1:0.25 reduce).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o rw_mix rw_mix.cu && ./rw_mix
#include <cstdio>
#include <cuda_runtime.h>
template <int NW>
__global__ void k(const float4* __restrict__ in, float4* __restrict__ o1, float4* __restrict__ o2, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float4 v = in[i];
        if (NW >= 1) o1[i] = v;
        if (NW >= 2) { v.x += 128.0f; o2[i] = v; }
        if (NW == 0 && v.x == 12345.678f) o1[0] = v;
    }
}
__global__ void wr(float4* o1, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) o1[i] = make_float4(1, 2, 3, 4);
}
int main() {
    const size_t n = (size_t)1 << 26;  // float4 elements: 1 GiB per buffer
    float4 *a, *b, *c;
    cudaMalloc(&a, n * 16); cudaMalloc(&b, n * 16); cudaMalloc(&c, n * 16);
    cudaMemset(a, 0, n * 16);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&](const char* name, int nw, double bytes) {
        float best = 1e9;
        for (int it = 0; it < 6; ++it) {
            cudaEventRecord(e0);
            if (nw == 0) k<0><<<148 * 16, 512>>>(a, b, c, n);
            else if (nw == 1) k<1><<<148 * 16, 512>>>(a, b, c, n);
            else if (nw == 2) k<2><<<148 * 16, 512>>>(a, b, c, n);
            else wr<<<148 * 16, 512>>>(b, n);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        printf("%-14s %.1f GB/s\n", name, bytes / best / 1e6);
    };
    run("read only", 0, n * 16.0); run("write only", 3, n * 16.0); run("1R:1W copy", 1, n * 32.0); run("1R:2W", 2, n * 48.0);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
