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

// (A) one launch over every scan layer of the pass; CTA = 512 columns x 32 rows of one layer, thread = 4 adjacent columns
// (one 16-byte load per layer and row; the first version read 7 scalars per pixel and was bound by the load unit, not by
// HBM).  Rows are taken four at a time: all loads of the block are issued before any of them is used.  "Some neighbour is
// greater than v" is max(neighbours) > v; the maximum over the 2 x 2 x 3 block is built from per-column maxima over the
// three layers (the left neighbour's comes from the lane to the left by shuffle, a warp's first lane loads its halo column)
// that are carried from row to row (fmaxf/fminf skip NaNs exactly like the chain of ordered compares they replace).
// PREFILTER: the two tests of _eliminateEdgeResponses that need no linear algebra (det < 0, edge ratio; sift.cpp:335-344,
// same expressions as eliminate.cu) only read the 3x3 neighbourhood of the middle layer, which streams through this
// kernel anyway (one row of look-ahead, one more column).  Candidates that fail them are marked in a second bit plane and
// come out of the emit kernel already filtered, so the elimination kernel gathers 19 DoG values only for the ~15 % that
// are left instead of re-reading most of the DoG pyramid sector by sector.
constexpr int kMaskWarpCols = 124;                          // output columns per warp: its 32 lanes load 128 columns, the first 4 are the left halo
constexpr int kMaskCols = (kExThreads / 32) * kMaskWarpCols;   // columns per CTA of the mask kernel

__device__ __forceinline__ float ex_max3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }
__device__ __forceinline__ float ex_min3(float a, float b, float c) { return fminf(fminf(a, b), c); }

// sift.cpp:335-344 on the middle layer's 3x3 neighbourhood (u = row above, c = this row, d = row below; l / m / r columns)
__device__ __forceinline__ bool ex_cheap_tests_pass(float ul, float um, float ur, float cl, float cm, float cr, float dl, float dm, float dr) {
    // algorithms.cpp:82-92
    const float dxx = cr + cl - 2 * cm;
    const float dyy = dm + um - 2 * cm;
    const float dxy = (dr - dl - ur + ul) / 2;
    const float t = (float)(121.0 / 10);
    const float tr = dxx + dyy;
    const float det = (float)((double)(dxx * dyy) - (double)dxy * (double)dxy);
    if (det < 0) return false;
    // tr^2 and t*det are exact in double (24-bit factors), so the quotient's side of t is known without the division unless it
    // lies within rounding distance of t (or det is 0 / NaN): divide only then
    const double a = (double)tr * (double)tr, bq = (double)t * (double)det;
    if (det > 0 && a > bq * (1.0 + 1e-12)) return false;
    if (det > 0 && a <= bq) return true;
    return !(a / (double)det > (double)t);
}

template <bool PREFILTER>
__global__ void __launch_bounds__(kExThreads) extrema_mask_kernel(const ScanLayer* __restrict__ layers, int n_layers, uint32_t* __restrict__ mask,
                                                                  uint32_t* __restrict__ pass_mask, uint32_t mask_words_per_image) {
    int li = 0;
    while (li + 1 < n_layers && layers[li + 1].tile_base <= blockIdx.x) ++li;
    const ScanLayer L = layers[li];
    const int tile = (int)(blockIdx.x - L.tile_base);
    const int yw = tile / (int)L.tiles_x;
    const int lane = threadIdx.x & 31;
    // this thread's columns x0 .. x0 + 3; a warp's first lane holds the four columns to the left of the warp's outputs (only its
    // last one is needed, as the left neighbour of the second lane's first column): one aligned 16-byte load like the others
    // instead of a scalar halo load per layer and row
    const int xw = (tile - yw * (int)L.tiles_x) * kMaskCols + (int)(threadIdx.x >> 5) * kMaskWarpCols;   // first output column of the warp
    const int x0 = xw + 4 * lane - 4;
    const int b = blockIdx.y;
    if (xw >= L.w) return;                           // whole warps only: the shuffles below need every lane
    const bool in_row = x0 >= 0 && x0 < L.w && lane > 0;   // lanes that own output columns
    const size_t img = (size_t)b * L.stride;
    const int xl = (x0 >= 0 && x0 < L.w) ? x0 : 0;   // a safe column for the others
    const float* d0 = L.d0 + img + xl;
    const float* d1 = L.d1 + img + xl;               // middle layer
    const float* d2 = L.d2 + img + xl;
    const int pitch = L.pitch, h = L.h, w = L.w;
    uint32_t word[4] = {0, 0, 0, 0};
    const int ybeg = yw * 32;

    // one row of the three layers at this thread's columns -> per-column max / min over the layers (index 0 = column x0 - 1)
    float pmax[5], pmin[5];   // previous row
    auto column_extremes = [&](const float4& a, const float4& m, const float4& c, float (&mx)[5], float (&mn)[5]) {
        mx[1] = ex_max3(a.x, m.x, c.x); mx[2] = ex_max3(a.y, m.y, c.y); mx[3] = ex_max3(a.z, m.z, c.z); mx[4] = ex_max3(a.w, m.w, c.w);
        mn[1] = ex_min3(a.x, m.x, c.x); mn[2] = ex_min3(a.y, m.y, c.y); mn[3] = ex_min3(a.z, m.z, c.z); mn[4] = ex_min3(a.w, m.w, c.w);
        mx[0] = __shfl_up_sync(0xffffffffu, mx[4], 1);   // (lane 0 gets its own value back: it owns no output)
        mn[0] = __shfl_up_sync(0xffffffffu, mn[4], 1);
    };
    {
        const size_t o = (size_t)(ybeg - 1 < 0 ? 0 : ybeg - 1) * pitch;
        const float4 a = *reinterpret_cast<const float4*>(d0 + o), m = *reinterpret_cast<const float4*>(d1 + o), c = *reinterpret_cast<const float4*>(d2 + o);
        column_extremes(a, m, c, pmax, pmin);
    }
    constexpr int RB = 4;
    for (int k0 = 0; k0 < 32 && ybeg + k0 < h; k0 += RB) {
        float4 va[RB], vm[RB], vc[RB];
#pragma unroll
        for (int q = 0; q < RB; ++q) {
            const int y = ybeg + k0 + q;
            const size_t o = (size_t)(y < h ? y : h - 1) * pitch;
            va[q] = *reinterpret_cast<const float4*>(d0 + o);
            vm[q] = *reinterpret_cast<const float4*>(d1 + o);
            vc[q] = *reinterpret_cast<const float4*>(d2 + o);
        }
#pragma unroll
        for (int q = 0; q < RB; ++q) {
            const int y = ybeg + k0 + q;
            float cmax[5], cmin[5];
            column_extremes(va[q], vm[q], vc[q], cmax, cmin);
            const float v[4] = {vm[q].x, vm[q].y, vm[q].z, vm[q].w};
            const bool row_ok = y >= 1 && y <= h - 2;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float mx = fmaxf(fmaxf(pmax[j], pmax[j + 1]), fmaxf(cmax[j], cmax[j + 1]));
                const float mn = fminf(fminf(pmin[j], pmin[j + 1]), fminf(cmin[j], cmin[j + 1]));
                const bool gt = mx > v[j], lt = mn < v[j];
                const int x = x0 + j;
                if ((!gt || !lt) && row_ok && in_row && x >= 1 && x <= w - 2) word[j] |= 1u << (k0 + q);
            }
#pragma unroll
            for (int j = 0; j < 5; ++j) { pmax[j] = cmax[j]; pmin[j] = cmin[j]; }
        }
    }
    if (!in_row) return;
    const size_t at = (size_t)b * mask_words_per_image + L.mask_off + (size_t)yw * w + x0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (x0 + j >= w) break;
        mask[at + j] = word[j];
        if (PREFILTER) {
            // the candidates (a few per cent of the pixels) fetch their 3x3 neighbourhood of the middle layer again: it has
            // just streamed through this SM's L1, and keeping three rows of it in registers next to the scan above would
            // double the kernel's register count for the sake of one pixel in twenty
            const float* c1 = L.d1 + img + x0 + j;
            uint32_t todo = word[j], pass = 0;
            float nb[9];
            auto fetch = [&](int k) {   // 1 <= y <= h - 2 and 1 <= x <= w - 2 for every candidate
                const float* r = c1 + (size_t)(ybeg + k) * pitch;
                nb[0] = __ldg(r - pitch - 1); nb[1] = __ldg(r - pitch); nb[2] = __ldg(r - pitch + 1);
                nb[3] = __ldg(r - 1); nb[4] = __ldg(r); nb[5] = __ldg(r + 1);
                nb[6] = __ldg(r + pitch - 1); nb[7] = __ldg(r + pitch); nb[8] = __ldg(r + pitch + 1);
            };
            int k = todo ? __ffs(todo) - 1 : 0;
            if (todo) fetch(k);
            while (todo) {   // the next candidate's loads are issued before this one's arithmetic
                todo &= todo - 1;
                const float c0 = nb[0], c1v = nb[1], c2 = nb[2], c3 = nb[3], c4 = nb[4], c5 = nb[5], c6 = nb[6], c7 = nb[7], c8 = nb[8];
                const int kk = k;
                if (todo) { k = __ffs(todo) - 1; fetch(k); }
                if (ex_cheap_tests_pass(c0, c1v, c2, c3, c4, c5, c6, c7, c8)) pass |= 1u << kk;
            }
            pass_mask[at + j] = pass;
        }
    }
}

void set_scan_tiles(ScanLayer* layers, int n_layers) {
    uint32_t base = 0;
    for (int l = 0; l < n_layers; ++l) {
        layers[l].tiles_x = (uint32_t)((layers[l].w + kMaskCols - 1) / kMaskCols);
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
