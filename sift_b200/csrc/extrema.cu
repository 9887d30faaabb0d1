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

// (A) one launch over every scan layer of the pass; CTA = 128 columns x 32 rows of one layer, thread = column x.
// Rows are taken four at a time: all loads of the block are issued before any of them is used.  "Some neighbour is
// greater than v" is max(neighbours) > v, and the maximum over the 2 x 2 x 3 block is built from per-row maxima that
// are carried from row to row (fmaxf/fminf skip NaNs exactly like the chain of ordered compares they replace).
// PREFILTER: the two tests of _eliminateEdgeResponses that need no linear algebra (det < 0, edge ratio; sift.cpp:335-344,
// same expressions as eliminate.cu) only read the 3x3 neighbourhood of the middle layer, which streams through this
// kernel anyway (one row of look-ahead, one more column).  Candidates that fail them are marked in a second bit plane and
// come out of the emit kernel already filtered, so the elimination kernel gathers 19 DoG values only for the ~15 % that
// are left instead of re-reading most of the DoG pyramid sector by sector.
template <bool PREFILTER>
__global__ void __launch_bounds__(kExThreads) extrema_mask_kernel(const ScanLayer* __restrict__ layers, int n_layers, uint32_t* __restrict__ mask,
                                                                  uint32_t* __restrict__ pass_mask, uint32_t mask_words_per_image) {
    int li = 0;
    while (li + 1 < n_layers && layers[li + 1].tile_base <= blockIdx.x) ++li;
    const ScanLayer L = layers[li];
    const int tile = (int)(blockIdx.x - L.tile_base);
    const int yw = tile / (int)L.tiles_x;
    const int x = (tile - yw * (int)L.tiles_x) * kExThreads + threadIdx.x;
    const int b = blockIdx.y;
    if (x >= L.w) return;
    const size_t img = (size_t)b * L.stride;
    const float* d1 = L.d1 + img + x;  // middle layer
    const float* d0 = L.d0 + img + x;
    const float* d2 = L.d2 + img + x;
    const int pitch = L.pitch, h = L.h;
    uint32_t word = 0, pass = 0;
    if (x >= 1 && x <= L.w - 2) {
        const int ybeg = yw * 32;
        // carried from the previous row: max / min over the 6 values (x-1, x) x 3 layers, and the middle layer's 3 values
        float pmax, pmin, pl1, pc1, pr1 = 0.0f;
        {
            const size_t o = (size_t)(ybeg - 1 < 0 ? 0 : ybeg - 1) * pitch;
            const float a0 = d0[o - 1], a1 = d0[o], b0 = d1[o - 1], b1 = d1[o], c0 = d2[o - 1], c1 = d2[o];
            pmax = fmaxf(fmaxf(fmaxf(a0, a1), fmaxf(b0, b1)), fmaxf(c0, c1));
            pmin = fminf(fminf(fminf(a0, a1), fminf(b0, b1)), fminf(c0, c1));
            pl1 = b0; pc1 = b1;
            if (PREFILTER) pr1 = d1[o + 1];
        }
        constexpr int RB = 4;
        for (int k0 = 0; k0 < 32 && ybeg + k0 < h; k0 += RB) {
            float v0l[RB], v0c[RB], v1l[RB], v1c[RB], v1r[RB], v2l[RB], v2c[RB];
#pragma unroll
            for (int q = 0; q < RB; ++q) {
                const int y = ybeg + k0 + q;
                const size_t o = (size_t)(y < h ? y : h - 1) * pitch;
                v0l[q] = d0[o - 1]; v0c[q] = d0[o];
                v1l[q] = d1[o - 1]; v1c[q] = d1[o];
                v2l[q] = d2[o - 1]; v2c[q] = d2[o];
                if (PREFILTER) v1r[q] = d1[o + 1];
            }
            float nl = 0.0f, nc = 0.0f, nr = 0.0f;  // middle layer, first row after this block (look-ahead of the last row)
            if (PREFILTER) {
                const int y = ybeg + k0 + RB;
                const size_t o = (size_t)(y < h ? y : h - 1) * pitch;
                nl = d1[o - 1]; nc = d1[o]; nr = d1[o + 1];
            }
#pragma unroll
            for (int q = 0; q < RB; ++q) {
                const int y = ybeg + k0 + q;
                const float cmax = fmaxf(fmaxf(fmaxf(v0l[q], v0c[q]), fmaxf(v1l[q], v1c[q])), fmaxf(v2l[q], v2c[q]));
                const float cmin = fminf(fminf(fminf(v0l[q], v0c[q]), fminf(v1l[q], v1c[q])), fminf(v2l[q], v2c[q]));
                const float v = v1c[q];
                const bool gt = fmaxf(pmax, cmax) > v, lt = fminf(pmin, cmin) < v;
                if ((!gt || !lt) && y >= 1 && y <= h - 2) {
                    word |= 1u << (k0 + q);
                    if (PREFILTER) {
                        // algorithms.cpp:82-92 on the middle layer, then sift.cpp:335-344
                        const float dn_l = q + 1 < RB ? v1l[q + 1 < RB ? q + 1 : q] : nl;
                        const float dn_c = q + 1 < RB ? v1c[q + 1 < RB ? q + 1 : q] : nc;
                        const float dn_r = q + 1 < RB ? v1r[q + 1 < RB ? q + 1 : q] : nr;
                        const float c11 = v;
                        const float dxx = v1r[q] + v1l[q] - 2 * c11;
                        const float dyy = dn_c + pc1 - 2 * c11;
                        const float dxy = (dn_r - dn_l - pr1 + pl1) / 2;
                        const float t = (float)(121.0 / 10);
                        const float tr = dxx + dyy;
                        const float det = (float)((double)(dxx * dyy) - (double)dxy * (double)dxy);
                        bool filtered = false;
                        if (det < 0) filtered = true;
                        else {
                            // tr^2 and t*det are exact in double (24-bit factors), so the quotient's side of t is known without
                            // the division unless it lies within rounding distance of t (or det is 0 / NaN): divide only then
                            const double a = (double)tr * (double)tr, bq = (double)t * (double)det;
                            if (det > 0 && a > bq * (1.0 + 1e-12)) filtered = true;
                            else if (det > 0 && a <= bq) filtered = false;
                            else if (a / (double)det > (double)t) filtered = true;
                        }
                        if (!filtered) pass |= 1u << (k0 + q);
                    }
                }
                pmax = cmax; pmin = cmin;
                pl1 = v1l[q]; pc1 = v1c[q];
                if (PREFILTER) pr1 = v1r[q];
            }
        }
    }
    const size_t at = (size_t)b * mask_words_per_image + L.mask_off + (size_t)yw * L.w + x;
    mask[at] = word;
    if (PREFILTER) pass_mask[at] = pass;
}

void set_scan_tiles(ScanLayer* layers, int n_layers) {
    uint32_t base = 0;
    for (int l = 0; l < n_layers; ++l) {
        layers[l].tiles_x = (uint32_t)((layers[l].w + kExThreads - 1) / kExThreads);
        layers[l].tile_base = base;
        base += layers[l].tiles_x * (uint32_t)layers[l].n_yw;
    }
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
                                                           const uint32_t* __restrict__ mask, const uint32_t* __restrict__ pass_mask,
                                                           uint32_t mask_words_per_image, const uint32_t* __restrict__ col_off,
                                                           Cand* __restrict__ cands, size_t cand_stride) {
    const int lane = threadIdx.x & 31;
    const int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int b = blockIdx.y;
    if (g >= total_cols) return;
    const int li = find_layer(layers, n_layers, (uint32_t)g);
    const ScanLayer L = layers[li];
    const int x = g - (int)L.col_base;
    const uint32_t* m = mask + (size_t)b * mask_words_per_image + L.mask_off + x;
    const uint32_t* pm = pass_mask ? pass_mask + (size_t)b * mask_words_per_image + L.mask_off + x : nullptr;
    uint32_t base = col_off[(size_t)b * total_cols + g];
    Cand* out = cands + (size_t)b * cand_stride;
    for (int yw0 = 0; yw0 < L.n_yw; yw0 += 32) {
        const int yw = yw0 + lane;
        uint32_t word = yw < L.n_yw ? m[(size_t)yw * L.w] : 0u;
        const uint32_t pass = (pm && yw < L.n_yw) ? pm[(size_t)yw * L.w] : 0xffffffffu;  // failed a cheap elimination test already
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
            c.filtered = (pass >> k) & 1u ? 0 : 1;
            c.pad = 0;
            out[slot++] = c;
        }
        base += __shfl_sync(0xffffffffu, incl, 31);
    }
}

int launch_extrema(const ScanLayer* layers_dev, const ScanLayer* layers_host, int n_layers, int total_cols,
                   uint32_t mask_words_per_image, uint32_t* mask, uint32_t* pass_mask, uint32_t* col_count, uint32_t* col_off,
                   Cand* cands, size_t cand_stride, uint32_t* n_cand, int batch, cudaStream_t s, uint64_t* launches) {
    {
        const ScanLayer& last = layers_host[n_layers - 1];
        dim3 grid(last.tile_base + last.tiles_x * (uint32_t)last.n_yw, batch);
        if (pass_mask) extrema_mask_kernel<true><<<grid, kExThreads, 0, s>>>(layers_dev, n_layers, mask, pass_mask, mask_words_per_image);
        else extrema_mask_kernel<false><<<grid, kExThreads, 0, s>>>(layers_dev, n_layers, mask, nullptr, mask_words_per_image);
        if (launches) ++*launches;
    }
    dim3 gcol((total_cols + 127) / 128, batch);
    extrema_count_kernel<<<gcol, 128, 0, s>>>(layers_dev, n_layers, total_cols, mask, mask_words_per_image, col_count);
    column_scan_kernel<<<batch, 1024, 0, s>>>(col_count, total_cols, col_off, n_cand);
    dim3 gwarp((total_cols + 7) / 8, batch);  // 8 warps (columns) per CTA
    extrema_emit_kernel<<<gwarp, 256, 0, s>>>(layers_dev, n_layers, total_cols, mask, pass_mask, mask_words_per_image, col_off, cands,
                                              cand_stride);
    if (launches) *launches += 3;
    SIFT_CUDA_TRY(cudaGetLastError());
    return 0;
}

}  // namespace siftgpu
