// Development probe: HBM bandwidth of the sliding blur's access pattern without any arithmetic.  Every warp walks down a strip of
// WC columns of a stack of 1920x1080 images (row lanes as in blur_slide.cu), reads each row once and writes it to NW planes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o strip_bw strip_bw.cu && ./strip_bw
#include <cstdio>
#include <cuda_runtime.h>
template <int V, int NW, int UNROLL>   // V floats per lane (2 -> 64-column strips, 4 -> 128)
__global__ void __launch_bounds__(64) k(const float* __restrict__ in, float* __restrict__ o1, float* __restrict__ o2, int w, int h, int frames, int strips, int lanes, int share) {
    const int gw = blockIdx.x * 2 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    const int rl = gw / strips;
    if (rl >= lanes) return;
    const int xs = (gw - rl * strips) * 32 * V + lane * V;
    long g = (long)rl * share, g_end = g + share;
    const long total = (long)frames * h;
    if (g_end > total) g_end = total;
    for (; g < g_end; g += UNROLL) {
        float v[UNROLL][V];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const long r = g + u < g_end ? g + u : g_end - 1;
            const float* p = in + r * w + xs;
            if (V == 2) { const float2 t = *reinterpret_cast<const float2*>(p); v[u][0] = t.x; v[u][1] = t.y; }
            else { const float4 t = *reinterpret_cast<const float4*>(p); v[u][0] = t.x; v[u][1] = t.y; v[u][2] = t.z; v[u][3] = t.w; }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            if (g + u >= g_end) break;
            float* q1 = o1 + (g + u) * w + xs;
            float* q2 = o2 + (g + u) * w + xs;
            if (V == 2) {
                *reinterpret_cast<float2*>(q1) = make_float2(v[u][0], v[u][1]);
                if (NW == 2) *reinterpret_cast<float2*>(q2) = make_float2(v[u][0] + 128.f, v[u][1] + 128.f);
            } else {
                *reinterpret_cast<float4*>(q1) = make_float4(v[u][0], v[u][1], v[u][2], v[u][3]);
                if (NW == 2) *reinterpret_cast<float4*>(q2) = make_float4(v[u][0] + 128.f, v[u][1] + 128.f, v[u][2] + 128.f, v[u][3] + 128.f);
            }
        }
    }
}
int main() {
    const int w = 1920, h = 1080, frames = 64;
    const size_t n = (size_t)w * h * frames;
    float *a, *b, *c;
    cudaMalloc(&a, n * 4); cudaMalloc(&b, n * 4); cudaMalloc(&c, n * 4);
    cudaMemset(a, 0, n * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&](const char* name, auto kern, int V, int nw, int cps) {
        const int strips = w / (32 * V), warps = 148 * cps * 2, lanes = warps / strips;
        const int share = (int)(((long)frames * h + lanes - 1) / lanes);
        float best = 1e9;
        for (int it = 0; it < 5; ++it) {
            cudaEventRecord(e0);
            kern<<<(lanes * strips + 1) / 2, 64>>>(a, b, c, w, h, frames, strips, lanes, share);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        printf("%-34s cps %2d: %.1f GB/s (%.1f us)\n", name, cps, (1.0 + nw) * n * 4 / best / 1e6, best * 1e3);
    };
    for (int cps : {9, 16, 32}) {
        run("64-col strips, 1R:2W, 8 rows/iter", k<2, 2, 8>, 2, 2, cps);
        run("128-col strips, 1R:2W, 8 rows/iter", k<4, 2, 8>, 4, 2, cps);
        run("64-col strips, 1R:1W, 8 rows/iter", k<2, 1, 8>, 2, 1, cps);
        run("128-col strips, 1R:1W, 8 rows/iter", k<4, 1, 8>, 4, 1, cps);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
