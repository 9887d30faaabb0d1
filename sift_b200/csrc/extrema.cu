// Scale-space extrema (reference sift.cpp:348-379): for the middle DoG layer a pixel
// (x, y), 1 <= x <= w-2, 1 <= y <= h-2, is a candidate iff over the 12 values
// {below, current, above} x {x-1, x} x {y-1, y} none is strictly greater OR none is strictly less
// (the reference's subarray((x-1,y-1),(x+1,y+1)) is half-open).  Candidates are emitted in the
// reference's order: octave, layer, x (outer), y (inner).  fp32 compares only => bit-exact.
//
// Four kernels: (A) predicate -> one bit per pixel, packed along y so a column's candidates are
// contiguous; (B) per-column popcounts; (C) exclusive scan over all columns of an image;
// (D) ordered emission.  (A) is the HBM-bound one: it reads each DoG layer once.
#include "common.cuh"

namespace siftgpu {

constexpr int kExThreads = 128;

// (A) one thread per column x, 32 rows per CTA row-block; rows y-1 are carried in registers.
__global__ void __launch_bounds__(kExThreads) extrema_mask_kernel(ScanLayer L, uint32_t* __restrict__ mask,
                                                                  uint32_t mask_words_per_image) {
    const int x = blockIdx.x * kExThreads + threadIdx.x;
    const int yw = blockIdx.y;
    const int b = blockIdx.z;
    if (x >= L.w) return;
    const float* d[3] = {L.d1 + (size_t)b * L.stride, L.d0 + (size_t)b * L.stride, L.d2 + (size_t)b * L.stride};
    uint32_t word = 0;
    if (x >= 1 && x <= L.w - 2) {
        const int ybeg = yw * 32;
        float pl[3], pc[3];  // previous row: (x-1, y-1), (x, y-1) per layer
        {
            const int yp = ybeg - 1 < 0 ? 0 : ybeg - 1;
#pragma unroll
            for (int l = 0; l < 3; ++l) {
                pl[l] = d[l][(size_t)yp * L.pitch + x - 1];
                pc[l] = d[l][(size_t)yp * L.pitch + x];
            }
        }
#pragma unroll 4
        for (int k = 0; k < 32; ++k) {
            const int y = ybeg + k;
            if (y >= L.h) break;
            float cl[3], cc[3];
#pragma unroll
            for (int l = 0; l < 3; ++l) {
                cl[l] = d[l][(size_t)y * L.pitch + x - 1];
                cc[l] = d[l][(size_t)y * L.pitch + x];
            }
            const float v = cc[0];
            bool gt = false, lt = false;
#pragma unroll
            for (int l = 0; l < 3; ++l) {
                gt |= (pl[l] > v) | (pc[l] > v) | (cl[l] > v) | (cc[l] > v);
                lt |= (pl[l] < v) | (pc[l] < v) | (cl[l] < v) | (cc[l] < v);
            }
            if ((!gt || !lt) && y >= 1 && y <= L.h - 2) word |= 1u << k;
#pragma unroll
            for (int l = 0; l < 3; ++l) { pl[l] = cl[l]; pc[l] = cc[l]; }
        }
    }
    mask[(size_t)b * mask_words_per_image + L.mask_off + (size_t)yw * L.w + x] = word;
}

__device__ __forceinline__ int find_layer(const ScanLayer* layers, int n_layers, uint32_t col) {
    int l = 0;
    while (l + 1 < n_layers && layers[l + 1].col_base <= col) ++l;
    return l;
}

// (B) one thread per global column.
__global__ void extrema_count_kernel(const ScanLayer* __restrict__ layers, int n_layers, int total_cols,
                                     const uint32_t* __restrict__ mask, uint32_t mask_words_per_image,
                                     uint32_t* __restrict__ col_count) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (g >= total_cols) return;
    const int li = find_layer(layers, n_layers, (uint32_t)g);
    const ScanLayer L = layers[li];
    const int x = g - (int)L.col_base;
    const uint32_t* m = mask + (size_t)b * mask_words_per_image + L.mask_off + x;
    uint32_t c = 0;
    for (int yw = 0; yw < L.n_yw; ++yw) c += __popc(m[(size_t)yw * L.w]);
    col_count[(size_t)b * total_cols + g] = c;
}

// (C) one CTA per image: exclusive scan of the column counts, total -> n_cand[b].
__global__ void __launch_bounds__(1024) column_scan_kernel(const uint32_t* __restrict__ col_count, int total_cols,
                                                           uint32_t* __restrict__ col_off, uint32_t* __restrict__ n_cand) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry;
    const int b = blockIdx.x;
    const uint32_t* in = col_count + (size_t)b * total_cols;
    uint32_t* out = col_off + (size_t)b * total_cols;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int base = 0; base < total_cols; base += 1024) {
        const int i = base + threadIdx.x;
        const uint32_t v = i < total_cols ? in[i] : 0u;
        uint32_t s = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += t;
        }
        if (lane == 31) warp_sums[wid] = s;
        __syncthreads();
        if (wid == 0) {
            uint32_t ws = warp_sums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, ws, o);
                if (lane >= o) ws += t;
            }
            warp_sums[lane] = ws;  // inclusive
        }
        __syncthreads();
        const uint32_t before = carry + (wid ? warp_sums[wid - 1] : 0u);
        if (i < total_cols) out[i] = before + s - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = before + s;
        __syncthreads();
    }
    if (threadIdx.x == 0) n_cand[b] = carry;
}

// (D) one warp per global column: lane l owns mask word l (32 rows each); a warp prefix sum of the popcounts gives
// every lane its slot, so the column's candidates come out top to bottom.
__global__ void __launch_bounds__(256) extrema_emit_kernel(const ScanLayer* __restrict__ layers, int n_layers, int total_cols,
                                                           const uint32_t* __restrict__ mask, uint32_t mask_words_per_image,
                                                           const uint32_t* __restrict__ col_off, Cand* __restrict__ cands,
                                                           size_t cand_stride) {
    const int lane = threadIdx.x & 31;
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int b = blockIdx.y;
    if (g >= total_cols) return;
    const int li = find_layer(layers, n_layers, (uint32_t)g);
    const ScanLayer L = layers[li];
    const int x = g - (int)L.col_base;
    const uint32_t* m = mask + (size_t)b * mask_words_per_image + L.mask_off + x;
    uint32_t base = col_off[(size_t)b * total_cols + g];
    Cand* out = cands + (size_t)b * cand_stride;
    for (int yw0 = 0; yw0 < L.n_yw; yw0 += 32) {
        const int yw = yw0 + lane;
        uint32_t word = yw < L.n_yw ? m[(size_t)yw * L.w] : 0u;
        const uint32_t cnt = (uint32_t)__popc(word);
        uint32_t incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        uint32_t slot = base + incl - cnt;
        while (word) {
            const int k = __ffs(word) - 1;
            word &= word - 1;
            Cand c;
            c.x = (uint16_t)x;
            c.y = (uint16_t)(yw * 32 + k);
            c.octave = L.octave;
            c.index = L.index;
            c.filtered = 0;
            c.pad = 0;
            out[slot++] = c;
        }
        base += __shfl_sync(0xffffffffu, incl, 31);
    }
}

int launch_extrema(const ScanLayer* layers_dev, const ScanLayer* layers_host, int n_layers, int total_cols,
                   uint32_t mask_words_per_image, uint32_t* mask, uint32_t* col_count, uint32_t* col_off,
                   Cand* cands, size_t cand_stride, uint32_t* n_cand, int batch, cudaStream_t s, uint64_t* launches) {
    for (int l = 0; l < n_layers; ++l) {
        const ScanLayer& L = layers_host[l];
        dim3 grid((L.w + kExThreads - 1) / kExThreads, L.n_yw, batch);
        extrema_mask_kernel<<<grid, kExThreads, 0, s>>>(L, mask, mask_words_per_image);
        if (launches) ++*launches;
    }
    dim3 gcol((total_cols + 127) / 128, batch);
    extrema_count_kernel<<<gcol, 128, 0, s>>>(layers_dev, n_layers, total_cols, mask, mask_words_per_image, col_count);
    column_scan_kernel<<<batch, 1024, 0, s>>>(col_count, total_cols, col_off, n_cand);
    dim3 gwarp((total_cols + 7) / 8, batch);  // 8 warps (columns) per CTA
    extrema_emit_kernel<<<gwarp, 256, 0, s>>>(layers_dev, n_layers, total_cols, mask, mask_words_per_image, col_off, cands,
                                              cand_stride);
    if (launches) *launches += 3;
    SIFT_CUDA_TRY(cudaGetLastError());
    return 0;
}

}  // namespace siftgpu
