// Gaussian pyramid kernels: alg::convolveWithGauss (reference algorithms.cpp:10-22) with the DoG
// subtraction (algorithms.cpp:52-64) fused into the epilogue, nearest-neighbour resize with the
// Vigra index walk (algorithms.cpp:33,46) and the u8 -> f32 widening of importImage
// (main.cpp:52-54).  Borders are BORDER_TREATMENT_REFLECT (index -j -> j, w-1+j -> w-1-j); taps are
// applied to source indices x-r..x+r in ascending order; the row pass is rounded to fp32 before the
// column pass.  Compiled with -fmad=false: the exact path is mul-then-add like the
// reference's mulss/addss, the FMA path asks for fmaf explicitly.
#include "common.cuh"

namespace siftgpu {

__device__ __forceinline__ int reflect101(int v, int n) {
    if (v < 0) v = -v;
    if (v >= n) v = 2 * (n - 1) - v;
    v = v < 0 ? 0 : v;            // only reachable for tile padding that no valid output reads
    return v >= n ? n - 1 : v;
}

template <bool FMA>
__device__ __forceinline__ float tap_acc(float acc, float t, float v) {
    if (FMA) return fmaf(t, v, acc);
    return __fadd_rn(acc, __fmul_rn(t, v));
}

// ---------------------------------------------------------------------------------------------
// Generic tile kernel (any radius): one CTA computes a TW x TH output tile from a reflected
// (TW+2r) x (TH+2r) input tile in shared memory.
constexpr int kTW = 64, kTH = 32, kBlurThreads = 256;

template <bool FMA>
__global__ void __launch_bounds__(kBlurThreads) blur_tile_kernel(BlurArgs a) {
    extern __shared__ float smem[];
    const int r = a.r, w = a.w, h = a.h;
    const int iw = kTW + 2 * r, ih = kTH + 2 * r;
    float* s_taps = smem;                  // 2r+1 (padded to a multiple of 4)
    float* s_in = smem + ((2 * r + 1 + 3) & ~3);
    float* s_tmp = s_in + iw * ih;         // ih rows x kTW

    const int b = blockIdx.z;
    const float* src = a.src + (size_t)b * a.src_stride;
    const int x0 = blockIdx.x * kTW, y0 = blockIdx.y * kTH;

    for (int i = threadIdx.x; i < 2 * r + 1; i += kBlurThreads) s_taps[i] = a.taps[i];
    for (int i = threadIdx.x; i < iw * ih; i += kBlurThreads) {
        const int ty = i / iw, tx = i - ty * iw;
        const int gx = reflect101(x0 - r + tx, w), gy = reflect101(y0 - r + ty, h);
        s_in[i] = src[(size_t)gy * w + gx];
    }
    __syncthreads();

    // row pass: kernel walked from +r down while the source index ascends
    for (int i = threadIdx.x; i < kTW * ih; i += kBlurThreads) {
        const int ty = i / kTW, tx = i - ty * kTW;
        const float* p = s_in + ty * iw + tx;
        float sum = 0.0f;
        for (int j = 0; j <= 2 * r; ++j) sum = tap_acc<FMA>(sum, s_taps[2 * r - j], p[j]);
        s_tmp[i] = sum;
    }
    __syncthreads();

    // column pass + epilogue
    for (int i = threadIdx.x; i < kTW * kTH; i += kBlurThreads) {
        const int ty = i / kTW, tx = i - ty * kTW;
        const int gx = x0 + tx, gy = y0 + ty;
        if (gx >= w || gy >= h) continue;
        const float* p = s_tmp + ty * kTW + tx;
        float sum = 0.0f;
        for (int j = 0; j <= 2 * r; ++j) sum = tap_acc<FMA>(sum, s_taps[2 * r - j], p[j * kTW]);
        const size_t o = (size_t)gy * w + gx;
        if (a.dst) a.dst[(size_t)b * a.dst_stride + o] = sum;
        if (a.dog) {
            const float lower = s_in[(ty + r) * iw + tx + r];
            const float dif = __fsub_rn(sum, lower);       // higher - lower
            a.dog[(size_t)b * a.dog_stride + o] = __fadd_rn(128.0f, dif);
        }
    }
}

int launch_blur(const BlurArgs& a, int batch, bool fma, cudaStream_t s, uint64_t* launches) {
    const int r = a.r;
    const size_t smem = sizeof(float) * (size_t)(((2 * r + 1 + 3) & ~3) + (kTW + 2 * r) * (kTH + 2 * r) + kTW * (kTH + 2 * r));
    dim3 grid((a.w + kTW - 1) / kTW, (a.h + kTH - 1) / kTH, batch);
    static bool attr_set[2] = {false, false};
    if (!attr_set[fma ? 1 : 0]) {
        if (fma) SIFT_CUDA_TRY(cudaFuncSetAttribute(blur_tile_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        else SIFT_CUDA_TRY(cudaFuncSetAttribute(blur_tile_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set[fma ? 1 : 0] = true;
    }
    if (fma) blur_tile_kernel<true><<<grid, kBlurThreads, smem, s>>>(a);
    else blur_tile_kernel<false><<<grid, kBlurThreads, smem, s>>>(a);
    if (launches) ++*launches;
    SIFT_CUDA_TRY(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// resizeImageNoInterpolation: dst(x, y) = src(map_x[x], map_y[y]); the maps are produced on the host
// by the literal accumulated-double walk.
__global__ void resize_nn_kernel(const float* __restrict__ src, size_t src_stride, int sw, float* __restrict__ dst,
                                 size_t dst_stride, int dw, int dh, const int* __restrict__ map_x,
                                 const int* __restrict__ map_y) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (x >= dw || y >= dh) return;
    const int b = blockIdx.z;
    dst[(size_t)b * dst_stride + (size_t)y * dw + x] = src[(size_t)b * src_stride + (size_t)map_y[y] * sw + map_x[x]];
}

int launch_resize_nn(const float* src, size_t src_stride, int sw, int sh, float* dst, size_t dst_stride, int dw, int dh,
                     const int* map_x, const int* map_y, int batch, cudaStream_t s, uint64_t* launches) {
    (void)sh;
    dim3 grid((dw + 255) / 256, dh, batch);
    resize_nn_kernel<<<grid, 256, 0, s>>>(src, src_stride, sw, dst, dst_stride, dw, dh, map_x, map_y);
    if (launches) ++*launches;
    SIFT_CUDA_TRY(cudaGetLastError());
    return 0;
}

__global__ void u8_to_f32_kernel(const uint8_t* __restrict__ src, size_t src_stride, float* __restrict__ dst,
                                 size_t dst_stride, size_t n) {
    const int b = blockIdx.y;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[(size_t)b * dst_stride + i] = (float)src[(size_t)b * src_stride + i];
}

int launch_u8_to_f32(const uint8_t* src, size_t src_stride, float* dst, size_t dst_stride, size_t n, int batch,
                     cudaStream_t s, uint64_t* launches) {
    dim3 grid((unsigned)((n + 256 * 8 - 1) / (256 * 8)), batch);
    u8_to_f32_kernel<<<grid, 256, 0, s>>>(src, src_stride, dst, dst_stride, n);
    if (launches) ++*launches;
    SIFT_CUDA_TRY(cudaGetLastError());
    return 0;
}

}  // namespace siftgpu
